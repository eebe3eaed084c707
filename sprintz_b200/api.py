"""Host-side mirror of the reference's operator interface for the codec path.

Function names, argument meaning and return conventions follow cpp/Compress/sprintz.h:16-32:
``sprintz_compress_{delta,xff}_{8b,16b}(src, ndims, write_size=True)`` returns the compressed stream and the
reference's return value (an ELEMENT count: floor(bytes/2) for 16-bit), ``sprintz_decompress_*`` returns the
decoded integers.  Both run the CUDA kernels through the C ABI on host (numpy) buffers.

``Frame`` + ``frame_encode`` / ``frame_decode`` are the throughput path on device (torch CUDA) tensors.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import DELTA, XFF, ChunkMeta, FrameHeader, SprintzError, check, lib

__all__ = [
    "DELTA", "XFF", "SprintzError",
    "compress", "decompress", "decompress_raw", "max_compressed_bytes",
    "sprintz_compress_delta_8b", "sprintz_decompress_delta_8b", "sprintz_compress_xff_8b",
    "sprintz_decompress_xff_8b", "sprintz_compress_delta_16b", "sprintz_decompress_delta_16b",
    "sprintz_compress_xff_16b", "sprintz_decompress_xff_16b",
    "Frame", "frame_encode", "frame_decode", "frame_compress_host", "frame_decompress_host",
    "compress_dev", "decompress_dev", "set_kernel_path",
    "RM_DELTA", "RM_PLAIN", "RM_XFF8", "compress_rowmajor", "decompress_rowmajor",
    "encode_forecast", "decode_forecast", "forecast_encode_dev", "forecast_decode_dev",
    "encode_xff_forecast", "decode_xff_forecast",
    "set_encode_overlap", "last_encode_overlapped", "Q_NOOP", "Q_MAX", "Q_SUM", "query", "frame_query", "reduce_rows_dev", "last_query_fused",
]

Q_NOOP, Q_MAX, Q_SUM = 0, 1, 2          # QueryTypes::Operation (query.hpp:22-24)

RM_DELTA, RM_PLAIN, RM_XFF8 = 1, 2, 3   # compress_rowmajor_delta_* / compress_rowmajor_* / compress8b_rowmajor_xff

_UDT = {1: np.uint8, 2: np.uint16}


def set_kernel_path(path: int) -> int:
    """0 = automatic, 1 = team-per-chunk kernels only, 2 = lane-per-octet kernels whenever the shape allows
    (include/sprintz_b200.h: spz_set_kernel_path).  Returns the previous setting."""
    return int(lib().spz_set_kernel_path(int(path)))


def max_compressed_bytes(esz: int, n: int, ndims: int) -> int:
    return int(lib().spz_max_compressed_bytes(esz, n, ndims))


# ------------------------------------------------------------------ reference-shaped host API

def compress(predictor: int, x: np.ndarray, ndims: int, write_size: bool = True):
    """-> (stream: uint8[true byte length], ret: the reference's return value in elements)."""
    x = np.ascontiguousarray(x)
    esz = x.dtype.itemsize
    if esz not in (1, 2):
        raise SprintzError("only 8- and 16-bit elements are supported (sprintz.h:16-32)")
    n = x.size
    if ndims == 0:
        return np.zeros(0, np.uint8), int(lib().spz_compress_host(predictor, esz, None, n, None, 0, 1, None))
    out = np.empty(max_compressed_bytes(esz, n, ndims), np.uint8)
    nb = C.c_uint64(0)
    ret = lib().spz_compress_host(predictor, esz, x.ctypes.data, n, out.ctypes.data, ndims, int(write_size),
                                  C.byref(nb))
    check(ret, "spz_compress_host")
    return out[: nb.value].copy(), int(ret)


def compress_rowmajor(kind: int, x: np.ndarray, ndims: int, write_size: bool = True):
    """Non-RLE sibling codecs (sprintz_delta.h:27-44): kind RM_DELTA = compress_rowmajor_delta_{8b,16b},
    RM_PLAIN = compress_rowmajor_{8b,16b}.  -> (stream, the reference's return value in elements)."""
    x = np.ascontiguousarray(x)
    esz = x.dtype.itemsize
    out = np.empty(max_compressed_bytes(esz, x.size, max(1, ndims)), np.uint8)
    nb = C.c_uint64(0)
    ret = lib().spz_compress_rowmajor_host(kind, esz, x.ctypes.data, x.size, out.ctypes.data, ndims, int(write_size),
                                           C.byref(nb))
    if ret == -1 and ndims == 0:
        return np.zeros(0, np.uint8), -1
    check(ret, "spz_compress_rowmajor_host")
    return out[: nb.value].copy(), int(ret)


def decompress_rowmajor(kind: int, esz: int, stream: np.ndarray):
    """-> (decoded elements, count); the length comes from the stream's 6-byte metadata."""
    stream = np.ascontiguousarray(stream, dtype=np.uint8)
    n = int(stream[:4].view(np.uint32)[0])
    out = np.empty(n, _UDT[esz])
    src = np.concatenate([stream, np.zeros(64, np.uint8)])
    ret = lib().spz_decompress_rowmajor_host(kind, esz, src.ctypes.data, out.ctypes.data)
    check(ret, "spz_decompress_rowmajor_host")
    return out, int(ret)


def encode_forecast(order: int, x: np.ndarray, ndims: int, write_size: bool = True):
    """Stand-alone forecaster transform (delta.h:17-67): order 1 = encode_delta_rowmajor_*, 2 =
    encode_doubledelta_rowmajor_*.  -> (output bytes incl. the optional 6-byte metadata, the reference's return)."""
    x = np.ascontiguousarray(x)
    esz = x.dtype.itemsize
    out = np.zeros(x.size * esz + 8, np.uint8)
    ret = lib().spz_encode_forecast_host(order, esz, x.ctypes.data, x.size, out.ctypes.data, ndims, int(write_size))
    check(ret, "spz_encode_forecast_host")
    return out[: int(ret) * esz].copy(), int(ret)


def decode_forecast(order: int, esz: int, y: np.ndarray, ndims: int):
    """y: the transformed elements without metadata. -> (decoded elements, count)"""
    y = np.ascontiguousarray(y)
    out = np.zeros(y.size, _UDT[esz])
    ret = lib().spz_decode_forecast_host(order, esz, y.ctypes.data, y.size, out.ctypes.data, ndims)
    check(ret, "spz_decode_forecast_host")
    return out, int(ret)


def encode_xff_forecast(x: np.ndarray, ndims: int, write_size: bool = True):
    """Stand-alone FIRE transform (predict.h:15-30): encode_xff_rowmajor_*.  -> (output bytes incl. the optional
    6-byte metadata, the reference's return)."""
    x = np.ascontiguousarray(x)
    esz = x.dtype.itemsize
    out = np.zeros(x.size * esz + 8, np.uint8)
    ret = lib().spz_xff_forecast_host(1, esz, x.ctypes.data, x.size, out.ctypes.data, ndims, int(write_size))
    check(ret, "spz_xff_forecast_host")
    return out[: int(ret) * esz].copy(), int(ret)


def decode_xff_forecast(esz: int, y: np.ndarray, ndims: int):
    """y: the transformed elements without metadata. -> (decoded elements, count)"""
    y = np.ascontiguousarray(y)
    out = np.zeros(y.size, _UDT[esz])
    ret = lib().spz_xff_forecast_host(0, esz, y.ctypes.data, y.size, out.ctypes.data, ndims, 0)
    check(ret, "spz_xff_forecast_host")
    return out, int(ret)


def forecast_encode_dev(order: int, x, ndims: int):
    """torch CUDA tensor in -> transformed tensor (same shape / dtype), on the current stream."""
    torch = _torch()
    n = x.numel()
    pad = torch.empty(n * x.element_size() + 32, dtype=torch.uint8, device=x.device)
    out = pad[: n * x.element_size()].view(x.dtype)
    check(lib().spz_forecast_encode_dev(order, x.element_size(), x.data_ptr(), n, ndims, out.data_ptr(), _stream_ptr()),
          "spz_forecast_encode_dev")
    return out


def forecast_decode_dev(order: int, y, ndims: int, work=None):
    torch = _torch()
    n = y.numel()
    wb = int(lib().spz_forecast_workspace_bytes(n, ndims))
    if work is None or work.numel() < wb:
        work = torch.empty(wb, dtype=torch.uint8, device=y.device)
    out = torch.empty_like(y)
    check(lib().spz_forecast_decode_dev(order, y.element_size(), y.data_ptr(), n, ndims, out.data_ptr(), work.data_ptr(),
                                        work.numel(), _stream_ptr()), "spz_forecast_decode_dev")
    return out


def measure(esz: int, stream: np.ndarray, has_meta=True, ndims=0, ngroups=0, remaining=0):
    sb, ne = C.c_uint64(0), C.c_uint64(0)
    check(lib().spz_measure_stream(esz, stream.ctypes.data, int(has_meta), ndims, ngroups, remaining,
                                   C.byref(sb), C.byref(ne)), "spz_measure_stream")
    return sb.value, ne.value


def decompress(predictor: int, esz: int, stream: np.ndarray):
    """-> (decoded uint array, ret: element count)."""
    stream = np.ascontiguousarray(stream).view(np.uint8).reshape(-1)
    padded = np.concatenate([stream, np.zeros(16, np.uint8)])
    _, ne = measure(esz, padded)
    out = np.empty(ne, _UDT[esz])
    ret = lib().spz_decompress_host(predictor, esz, padded.ctypes.data, out.ctypes.data)
    check(ret, "spz_decompress_host")
    return out[: max(ret, 0)], int(ret)


def decompress_raw(predictor: int, esz: int, stream: np.ndarray, ndims: int, ngroups: int, remaining: int):
    stream = np.ascontiguousarray(stream).view(np.uint8).reshape(-1)
    padded = np.concatenate([stream, np.zeros(16, np.uint8)])
    _, ne = measure(esz, padded, False, ndims, ngroups, remaining)
    out = np.empty(ne, _UDT[esz])
    ret = lib().spz_decompress_raw_host(predictor, esz, padded.ctypes.data, out.ctypes.data, ndims, ngroups,
                                        remaining)
    check(ret, "spz_decompress_raw_host")
    return out[: max(ret, 0)], int(ret)


def _mk(predictor, esz):
    def comp(src: np.ndarray, ndims: int, write_size: bool = True):
        if src.dtype.itemsize != esz:
            raise SprintzError(f"expected {8 * esz}-bit elements")
        return compress(predictor, src, ndims, write_size)

    def decomp(stream: np.ndarray):
        return decompress(predictor, esz, stream)

    return comp, decomp


sprintz_compress_delta_8b, sprintz_decompress_delta_8b = _mk(DELTA, 1)
sprintz_compress_xff_8b, sprintz_decompress_xff_8b = _mk(XFF, 1)
sprintz_compress_delta_16b, sprintz_decompress_delta_16b = _mk(DELTA, 2)
sprintz_compress_xff_16b, sprintz_decompress_xff_16b = _mk(XFF, 2)


# ------------------------------------------------------------------ device (torch) API

def _torch():
    import torch
    return torch


def _stream_ptr(stream=None):
    torch = _torch()
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


def compress_dev(predictor: int, x, ndims: int, write_size: bool = True):
    """Single reference stream from a CUDA tensor (uint8 / int16-viewable). -> (uint8 CUDA tensor, ChunkMeta)."""
    torch = _torch()
    esz = x.element_size()
    n = x.numel()
    out = torch.empty(max_compressed_bytes(esz, n, ndims), dtype=torch.uint8, device=x.device)
    meta = torch.zeros(4, dtype=torch.int32, device=x.device)
    check(lib().spz_compress_dev(predictor, esz, x.data_ptr(), n, ndims, int(write_size), out.data_ptr(),
                                 meta.data_ptr(), _stream_ptr()), "spz_compress_dev")
    m = meta.cpu().numpy().view(np.uint32)
    return out[: int(m[0])], ChunkMeta(int(m[0]), int(m[1]), int(m[2]), int(m[3]))


def decompress_dev(predictor: int, esz: int, stream, ndims: int, n_out: int):
    torch = _torch()
    src = torch.zeros(stream.numel() + 16, dtype=torch.uint8, device=stream.device)
    src[: stream.numel()] = stream
    out = torch.empty(n_out, dtype=torch.uint8 if esz == 1 else torch.int16, device=stream.device)
    meta = torch.zeros(4, dtype=torch.int32, device=stream.device)
    check(lib().spz_decompress_dev(predictor, esz, src.data_ptr(), src.numel(), 1, ndims, out.data_ptr(), n_out,
                                   meta.data_ptr(), _stream_ptr()), "spz_decompress_dev")
    m = meta.cpu().numpy().view(np.uint32)
    if int(m[3]) != 0:
        raise SprintzError(f"decode status {int(m[3])}")
    return out, int(m[2])


@dataclass
class Frame:
    """A compressed frame resident in HBM plus the call parameters needed to decode it."""
    buf: "object"            # uint8 CUDA tensor (capacity spz_frame_max_bytes)
    predictor: int
    elem_sz: int
    len: int
    ndims: int
    chunk_rows: int
    nbytes_dev: "object"     # 1-element int64 CUDA tensor: frame size in bytes

    @property
    def nchunks(self) -> int:
        return int(lib().spz_frame_nchunks(self.len, self.ndims, self.chunk_rows))

    def nbytes(self) -> int:
        return int(self.nbytes_dev.item())

    def header(self) -> FrameHeader:
        h = FrameHeader()
        check(lib().spz_frame_read_header(self.buf.data_ptr(), C.byref(h), _stream_ptr()), "spz_frame_read_header")
        return h

    def directory(self):
        """-> (chunk_off int64[nchunks+1], chunk_nbytes uint32[nchunks], payload_off) as numpy (host)."""
        h = self.header()
        n = int(h.nchunks)
        raw = self.buf[64: int(h.payload_off)].cpu().numpy()
        off = raw[: 8 * (n + 1)].view(np.uint64).astype(np.int64)
        nb = raw[8 * (n + 1): 8 * (n + 1) + 4 * n].view(np.uint32)
        return off, nb, int(h.payload_off)

    def chunk_stream(self, c: int) -> np.ndarray:
        off, nb, po = self.directory()
        a = po + int(off[c])
        return self.buf[a: a + int(nb[c])].cpu().numpy()


class FrameCodec:
    """Pre-allocated buffers for repeated frame encode/decode of one (codec, shape): the measured path."""

    def __init__(self, predictor: int, elem_sz: int, length: int, ndims: int, chunk_rows: int, device="cuda"):
        torch = _torch()
        L = lib()
        self.predictor, self.elem_sz, self.len, self.ndims, self.chunk_rows = predictor, elem_sz, length, ndims, chunk_rows
        self.frame_cap = int(L.spz_frame_max_bytes(elem_sz, length, ndims, chunk_rows))
        self.work_bytes = int(L.spz_frame_workspace_bytes(elem_sz, length, ndims, chunk_rows))
        self.nchunks = int(L.spz_frame_nchunks(length, ndims, chunk_rows))
        self.frame = torch.empty(self.frame_cap, dtype=torch.uint8, device=device)
        self.work = torch.empty(self.work_bytes, dtype=torch.uint8, device=device)
        self.nbytes_dev = torch.zeros(1, dtype=torch.int64, device=device)
        self.status = torch.zeros(4 * self.nchunks, dtype=torch.int32, device=device)

    def encode(self, x) -> Frame:
        assert x.numel() == self.len and x.element_size() == self.elem_sz and x.is_cuda
        check(lib().spz_frame_encode_dev(self.predictor, self.elem_sz, x.data_ptr(), self.len, self.ndims,
                                         self.chunk_rows, self.frame.data_ptr(), self.frame_cap,
                                         self.work.data_ptr(), self.work_bytes, self.nbytes_dev.data_ptr(),
                                         _stream_ptr()), "spz_frame_encode_dev")
        return Frame(self.frame, self.predictor, self.elem_sz, self.len, self.ndims, self.chunk_rows, self.nbytes_dev)

    def decode(self, frame: Frame, out):
        assert out.numel() >= self.len and out.element_size() == self.elem_sz and out.is_cuda
        check(lib().spz_frame_decode_dev(frame.predictor, frame.elem_sz, frame.buf.data_ptr(), frame.buf.numel(),
                                         frame.len, frame.ndims, frame.chunk_rows, out.data_ptr(), out.numel(),
                                         self.status.data_ptr(), _stream_ptr()), "spz_frame_decode_dev")
        return out

    def check_status(self):
        st = self.status.view(-1, 4)[:, 3]
        bad = int((st != 0).sum().item())
        if bad:
            raise SprintzError(f"{bad} chunk(s) failed to decode")


def frame_encode(predictor: int, x, ndims: int, chunk_rows: int) -> Frame:
    codec = FrameCodec(predictor, x.element_size(), x.numel(), ndims, chunk_rows, device=x.device)
    return codec.encode(x)


def frame_decode(frame: Frame):
    torch = _torch()
    out = torch.empty(frame.len, dtype=torch.uint8 if frame.elem_sz == 1 else torch.int16, device=frame.buf.device)
    codec = FrameCodec.__new__(FrameCodec)
    codec.len, codec.elem_sz = frame.len, frame.elem_sz
    codec.status = torch.zeros(4 * frame.nchunks, dtype=torch.int32, device=frame.buf.device)
    FrameCodec.decode(codec, frame, out)
    FrameCodec.check_status(codec)
    return out


def frame_query(frame: Frame, op: int, materialize: bool = False, out=None, work=None):
    """Query-on-compressed over a device frame (query.hpp:22-29). -> (result int64[ndims] CUDA tensor, decoded
    tensor or None).  materialize=False on frames of many 8/16-dim chunks runs fused (last_query_fused() == 1)."""
    torch = _torch()
    L = lib()
    dev = frame.buf.device
    res = torch.empty(frame.ndims, dtype=torch.int64, device=dev)
    status = torch.zeros(4 * frame.nchunks, dtype=torch.int32, device=dev)
    if materialize and out is None:
        out = torch.empty(frame.len, dtype=torch.uint8 if frame.elem_sz == 1 else torch.int16, device=dev)
    wb = int(L.spz_frame_query_workspace_bytes(frame.elem_sz, frame.len, int(materialize)))
    if wb and (work is None or work.numel() < wb):
        work = torch.empty(wb, dtype=torch.uint8, device=dev)
    check(L.spz_frame_query_dev(frame.predictor, frame.elem_sz, op, int(materialize), frame.buf.data_ptr(),
                                frame.buf.numel(), frame.len, frame.ndims, frame.chunk_rows,
                                out.data_ptr() if out is not None else None, out.numel() if out is not None else 0,
                                status.data_ptr(), res.data_ptr(), work.data_ptr() if work is not None else None,
                                work.numel() if work is not None else 0, _stream_ptr()), "spz_frame_query_dev")
    bad = int((status.view(-1, 4)[:, 3] != 0).sum().item())
    if bad:
        raise SprintzError(f"{bad} chunk(s) failed to decode")
    return res, (out if materialize else None)


def set_encode_overlap(mode: int) -> int:
    """0 (default): directory scan + compaction after the encoder; 1: EXPERIMENTAL overlapped scan + copy behind the lane
    encoders, applied to frames of up to 148 x 128 chunks only (the verified configuration).  -> previous mode"""
    return int(lib().spz_set_encode_overlap(mode))


def last_encode_overlapped() -> bool:
    return bool(lib().spz_last_encode_overlapped())


def last_query_fused() -> bool:
    return bool(lib().spz_last_query_fused())


def reduce_rows_dev(op: int, x, ndims: int):
    """per-dim max / sum of a row-major CUDA tensor -> int64[ndims] CUDA tensor"""
    torch = _torch()
    res = torch.empty(ndims, dtype=torch.int64, device=x.device)
    check(lib().spz_reduce_rows_dev(op, x.element_size(), x.data_ptr(), x.numel(), ndims, res.data_ptr(), _stream_ptr()),
          "spz_reduce_rows_dev")
    return res


def query(predictor: int, esz: int, stream: np.ndarray, op: int, materialize: bool = True):
    """query_rowmajor_{delta,xff}_rle_{8b,16b} on a host stream (sprintz_delta.h:95-98, sprintz_xff.h:90-93).
    -> (result uint64[ndims], decoded elements or None, the reference's return value)"""
    stream = np.ascontiguousarray(stream).view(np.uint8).reshape(-1)
    padded = np.concatenate([stream, np.zeros(16, np.uint8)])
    _, ne = measure(esz, padded)
    ndims = int(padded[6]) | (int(padded[7]) << 8)
    out = np.empty(ne, _UDT[esz])
    res = np.zeros(max(ndims, 1), np.uint64)
    ret = lib().spz_query_host(predictor, esz, padded.ctypes.data, out.ctypes.data if materialize else None, op,
                               int(materialize), res.ctypes.data)
    check(ret, "spz_query_host")
    return res, (out[: max(ret, 0)] if materialize else None), int(ret)


def frame_compress_host(predictor: int, x: np.ndarray, ndims: int, chunk_rows: int) -> np.ndarray:
    x = np.ascontiguousarray(x)
    esz = x.dtype.itemsize
    cap = int(lib().spz_frame_max_bytes(esz, x.size, ndims, chunk_rows))
    out = np.empty(cap, np.uint8)
    ret = lib().spz_frame_compress_host(predictor, esz, x.ctypes.data, x.size, ndims, chunk_rows, out.ctypes.data, cap)
    check(ret, "spz_frame_compress_host")
    return out[:ret].copy()


def frame_decompress_host(frame: np.ndarray) -> np.ndarray:
    frame = np.ascontiguousarray(frame).view(np.uint8).reshape(-1)
    h = FrameHeader.from_buffer_copy(frame[:64].tobytes())
    out = np.empty(int(h.len), _UDT[int(h.elem_sz)])
    ret = lib().spz_frame_decompress_host(frame.ctypes.data, frame.size, out.ctypes.data, out.size)
    check(ret, "spz_frame_decompress_host")
    return out
