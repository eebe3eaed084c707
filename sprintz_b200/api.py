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
    "set_encode_pipeline", "last_encode_slabs", "kernel_launches", "Communicator", "FrameGather", "shard_len",
    "HuffCodec", "huff_encode_dev", "huff_decode_dev",
    "quantize", "quantize_dev", "Q_NOOP", "Q_MAX", "Q_SUM", "query", "frame_query", "reduce_rows_dev", "last_query_fused",
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
    if n >= 1 << 32:
        raise SprintzError("a single reference stream holds fewer than 2**32 elements (uint32_t len, sprintz.h:16-32); "
                           "use the frame API")
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
    if x.size >= 1 << 32:
        raise SprintzError("a single reference stream holds fewer than 2**32 elements (uint32_t len)")
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


def compress_dev(predictor: int, x, ndims: int, write_size: bool = True, block_rows: int = 8):
    """Single reference stream from a CUDA tensor (uint8 / int16-viewable). -> (uint8 CUDA tensor, ChunkMeta).
    block_rows != 8: BASELINE config 5's block-size axis (spz_compress_bs_dev)."""
    torch = _torch()
    esz = x.element_size()
    n = x.numel()
    out = torch.empty(max_compressed_bytes(esz, n, ndims), dtype=torch.uint8, device=x.device)
    meta = torch.zeros(4, dtype=torch.int32, device=x.device)
    if block_rows != 8:
        check(lib().spz_compress_bs_dev(predictor, esz, block_rows, x.data_ptr(), n, ndims, int(write_size), out.data_ptr(),
                                        meta.data_ptr(), _stream_ptr()), "spz_compress_bs_dev")
    else:
        check(lib().spz_compress_dev(predictor, esz, x.data_ptr(), n, ndims, int(write_size), out.data_ptr(),
                                     meta.data_ptr(), _stream_ptr()), "spz_compress_dev")
    m = meta.cpu().numpy().view(np.uint32)
    return out[: int(m[0])], ChunkMeta(int(m[0]), int(m[1]), int(m[2]), int(m[3]))


def decompress_dev(predictor: int, esz: int, stream, ndims: int, n_out: int, block_rows: int = 8):
    torch = _torch()
    src = torch.zeros(stream.numel() + 16, dtype=torch.uint8, device=stream.device)
    src[: stream.numel()] = stream
    out = torch.empty(n_out, dtype=torch.uint8 if esz == 1 else torch.int16, device=stream.device)
    meta = torch.zeros(4, dtype=torch.int32, device=stream.device)
    if block_rows != 8:
        check(lib().spz_decompress_bs_dev(predictor, esz, block_rows, src.data_ptr(), src.numel(), 1, ndims, out.data_ptr(),
                                          n_out, meta.data_ptr(), _stream_ptr()), "spz_decompress_bs_dev")
    else:
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
    block_rows: int = 8      # rows per block (8 = the reference's format; 16 / 32 / 64 = BASELINE config 5's axis)

    @property
    def nchunks(self) -> int:
        return int(lib().spz_frame_nchunks(self.len, self.ndims, self.chunk_rows))

    def nbytes(self) -> int:
        if self.nbytes_dev is None:
            return int(self.header().frame_bytes)
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

    def __init__(self, predictor: int, elem_sz: int, length: int, ndims: int, chunk_rows: int, device="cuda", block_rows: int = 8):
        torch = _torch()
        L = lib()
        self.predictor, self.elem_sz, self.len, self.ndims, self.chunk_rows = predictor, elem_sz, length, ndims, chunk_rows
        self.block_rows = block_rows
        self.frame_cap = int(L.spz_frame_max_bytes(elem_sz, length, ndims, chunk_rows))
        self.work_bytes = int(L.spz_frame_workspace_bytes(elem_sz, length, ndims, chunk_rows))
        self.nchunks = int(L.spz_frame_nchunks(length, ndims, chunk_rows))
        self.frame = torch.empty(self.frame_cap, dtype=torch.uint8, device=device)
        self.work = torch.empty(self.work_bytes, dtype=torch.uint8, device=device)
        self.nbytes_dev = torch.zeros(1, dtype=torch.int64, device=device)
        self.status = torch.zeros(4 * self.nchunks, dtype=torch.int32, device=device)

    def encode(self, x) -> Frame:
        assert x.numel() == self.len and x.element_size() == self.elem_sz and x.is_cuda
        if self.block_rows != 8:
            check(lib().spz_frame_encode_bs_dev(self.predictor, self.elem_sz, self.block_rows, x.data_ptr(), self.len, self.ndims,
                                                self.chunk_rows, self.frame.data_ptr(), self.frame_cap,
                                                self.work.data_ptr(), self.work_bytes, self.nbytes_dev.data_ptr(),
                                                _stream_ptr()), "spz_frame_encode_bs_dev")
        else:
            check(lib().spz_frame_encode_dev(self.predictor, self.elem_sz, x.data_ptr(), self.len, self.ndims,
                                             self.chunk_rows, self.frame.data_ptr(), self.frame_cap,
                                             self.work.data_ptr(), self.work_bytes, self.nbytes_dev.data_ptr(),
                                             _stream_ptr()), "spz_frame_encode_dev")
        return Frame(self.frame, self.predictor, self.elem_sz, self.len, self.ndims, self.chunk_rows, self.nbytes_dev,
                     self.block_rows)

    def decode(self, frame: Frame, out):
        assert out.numel() >= self.len and out.element_size() == self.elem_sz and out.is_cuda
        if frame.block_rows != 8:
            check(lib().spz_frame_decode_bs_dev(frame.predictor, frame.elem_sz, frame.block_rows, frame.buf.data_ptr(),
                                                frame.buf.numel(), frame.len, frame.ndims, frame.chunk_rows, out.data_ptr(),
                                                out.numel(), self.status.data_ptr(), _stream_ptr()), "spz_frame_decode_bs_dev")
        else:
            check(lib().spz_frame_decode_dev(frame.predictor, frame.elem_sz, frame.buf.data_ptr(), frame.buf.numel(),
                                             frame.len, frame.ndims, frame.chunk_rows, out.data_ptr(), out.numel(),
                                             self.status.data_ptr(), _stream_ptr()), "spz_frame_decode_dev")
        return out

    def check_status(self):
        st = self.status.view(-1, 4)[:, 3]
        bad = int((st != 0).sum().item())
        if bad:
            raise SprintzError(f"{bad} chunk(s) failed to decode")


def frame_encode(predictor: int, x, ndims: int, chunk_rows: int, block_rows: int = 8) -> Frame:
    codec = FrameCodec(predictor, x.element_size(), x.numel(), ndims, chunk_rows, device=x.device, block_rows=block_rows)
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


def set_encode_pipeline(slabs: int) -> int:
    """Slabs of a frame encode (include/sprintz_b200.h: spz_set_encode_pipeline): 0 = automatic, 1 = codec kernel, directory
    scan and compaction back to back, n > 1 = n slabs whose scan + compaction overlap the next slab's encoding.
    -> previous setting"""
    return int(lib().spz_set_encode_pipeline(int(slabs)))


def last_encode_slabs() -> int:
    return int(lib().spz_last_encode_slabs())


def kernel_launches() -> int:
    return int(lib().spz_kernel_launches())


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


# ------------------------------------------------------------------ float -> integer quantiser

def quantize(x: np.ndarray, ndims: int, dtype=np.uint8) -> np.ndarray:
    """_quantize(mat, dtype, axis=0) of python/datasets/compress_bench.py:45-60 on a row-major (rows x ndims) float32 /
    float64 array -> flat uint8 / uint16 array, computed on the GPU."""
    x = np.ascontiguousarray(x)
    if x.dtype not in (np.float32, np.float64):
        raise SprintzError("quantize expects float32 or float64 input")
    esz = np.dtype(dtype).itemsize
    out = np.empty(x.size, _UDT[esz])
    check(lib().spz_quantize_host(esz, int(x.dtype == np.float64), x.ctypes.data, x.size, ndims, out.ctypes.data),
          "spz_quantize_host")
    return out


def quantize_dev(x, ndims: int, elem_sz: int):
    """torch CUDA float32 / float64 tensor -> uint8 / int16 (uint16 bit pattern) tensor, on the current stream."""
    torch = _torch()
    out = torch.empty(x.numel(), dtype=torch.uint8 if elem_sz == 1 else torch.int16, device=x.device)
    wb = int(lib().spz_quantize_workspace_bytes(ndims))
    work = torch.empty(wb, dtype=torch.uint8, device=x.device)
    check(lib().spz_quantize_dev(elem_sz, int(x.dtype == torch.float64), x.data_ptr(), x.numel(), ndims, out.data_ptr(),
                                 work.data_ptr(), wb, _stream_ptr()), "spz_quantize_dev")
    return out


# ------------------------------------------------------------------ entropy stage (include/sprintz_b200.h section 6)

class HuffCodec:
    """Pre-allocated buffers for the byte-wise Huffman stage over `nbytes` device bytes (parity unpinned: our container)."""

    def __init__(self, nbytes: int, device="cuda"):
        torch = _torch()
        L = lib()
        self.nbytes = nbytes
        self.cap = int(L.spz_huff_max_bytes(nbytes))
        self.work_bytes = int(L.spz_huff_workspace_bytes(nbytes))
        self.cont = torch.empty(self.cap, dtype=torch.uint8, device=device)
        self.work = torch.empty(self.work_bytes, dtype=torch.uint8, device=device)
        self.cont_bytes = torch.zeros(1, dtype=torch.int64, device=device)
        self.status = torch.zeros(1, dtype=torch.int32, device=device)

    def encode(self, x):
        """x: uint8 CUDA tensor of nbytes (16-byte aligned). -> the container tensor (self.cont[: size()])"""
        assert x.is_cuda and x.element_size() == 1 and x.numel() == self.nbytes
        check(lib().spz_huff_encode_dev(x.data_ptr(), self.nbytes, self.cont.data_ptr(), self.cap, self.work.data_ptr(),
                                        self.work_bytes, self.cont_bytes.data_ptr(), _stream_ptr()), "spz_huff_encode_dev")
        return self.cont

    def size(self) -> int:
        return int(self.cont_bytes.item())

    def decode(self, cont, out, cont_bytes=None):
        assert out.is_cuda and out.numel() >= self.nbytes
        check(lib().spz_huff_decode_dev(cont.data_ptr(), cont.numel() if cont_bytes is None else cont_bytes, self.nbytes,
                                        out.data_ptr(), out.numel(), self.work.data_ptr(), self.work_bytes,
                                        self.status.data_ptr(), _stream_ptr()), "spz_huff_decode_dev")
        return out

    def check_status(self):
        st = int(self.status.item())
        if st:
            raise SprintzError(f"huffman decode status {st}")


def huff_encode_dev(x):
    """-> uint8 CUDA tensor holding exactly the container"""
    h = HuffCodec(x.numel(), device=x.device)
    h.encode(x)
    return h.cont[: h.size()].clone()


def huff_decode_dev(cont, nbytes: int):
    torch = _torch()
    h = HuffCodec(nbytes, device=cont.device)
    out = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=cont.device)
    h.decode(cont, out)
    h.check_status()
    return out[:nbytes]


# ------------------------------------------------------------------ multi-GPU: one process per GPU (include/sprintz_b200.h section 4)

def shard_len(total_len: int, ndims: int, chunk_rows: int, stripe_chunks: int, nranks: int, rank: int) -> int:
    """Elements of a global stream that `rank` owns (stripes of stripe_chunks chunks dealt round-robin; 0 = one
    contiguous balanced range of whole chunks per rank)."""
    return int(lib().spz_shard_len(total_len, ndims, chunk_rows, stripe_chunks, nranks, rank))


class Communicator:
    """The library's NCCL communicators of this process (one process per GPU)."""

    def __init__(self, handle, rank, world):
        self.handle, self.rank, self.world = handle, rank, world

    @classmethod
    def from_torch(cls, rank: int, world: int):
        """Bootstraps through torch.distributed (already initialised by the harness): rank 0 draws the NCCL unique ids,
        everyone receives them, spz_comm_init joins."""
        import torch.distributed as dist
        torch = _torch()
        ident = C.create_string_buffer(256)
        if rank == 0:
            check(lib().spz_comm_unique_id(ident), "spz_comm_unique_id")
        backend = dist.get_backend()
        t = torch.frombuffer(bytearray(ident.raw), dtype=torch.uint8).clone()
        if backend == "nccl":
            t = t.cuda()
        dist.broadcast(t, src=0)
        blob = bytes(t.cpu().numpy().tobytes())
        h = C.c_void_p(None)
        check(lib().spz_comm_init(C.byref(h), world, rank, blob), "spz_comm_init")
        return cls(h, rank, world)

    def close(self):
        if self.handle:
            lib().spz_comm_destroy(self.handle)
            self.handle = None


class FrameGather:
    """Gathers every rank's frame of one codec shape into ONE decodable frame of the global stream on rank `dst`
    (spz_frame_gather_dev / spz_frame_encode_gather_dev).  All ranks hold shards of equal length here."""

    def __init__(self, comm: Communicator, codec: FrameCodec, dst: int = 0, stripe_chunks: int = 0, alloc: bool = True):
        torch = _torch()
        L = lib()
        self.comm, self.codec, self.dst, self.stripe_chunks = comm, codec, dst, stripe_chunks
        self.total_len = codec.len * comm.world
        for s in (0, stripe_chunks):
            if shard_len(self.total_len, codec.ndims, codec.chunk_rows, s, comm.world, comm.rank) != codec.len:
                raise SprintzError("FrameGather: shards must be whole stripes of equal length on every rank")
        self.cap = 0
        self.buf = None
        if comm.rank == dst:
            self.cap = int(L.spz_frame_gather_max_bytes(codec.elem_sz, self.total_len, codec.ndims, codec.chunk_rows))
            if alloc:                               # alloc=False: only the peer-memory gather is used (its buffer is the library's)
                self.buf = torch.empty(self.cap, dtype=torch.uint8, device=codec.frame.device)
        self.last_frame = None
        self.last_stripes = 0

    def _frame(self, stripes):
        if self.comm.rank != self.dst:
            return None
        self.last_stripes = stripes
        self.last_frame = Frame(self.buf, self.codec.predictor, self.codec.elem_sz, self.total_len, self.codec.ndims,
                                self.codec.chunk_rows, None)
        return self.last_frame

    def gather(self, frame: Frame):
        """contiguous shards (rank-major): the global stream is the concatenation of the ranks' shards"""
        check(lib().spz_frame_gather_dev(self.comm.handle, self.dst, self.total_len, 0, frame.buf.data_ptr(),
                                         self.buf.data_ptr() if self.buf is not None else None, self.cap, _stream_ptr()),
              "spz_frame_gather_dev")
        return self._frame(0)

    def encode_and_gather(self, x):
        """streamed: stripes of stripe_chunks chunks dealt round-robin; stripe k travels while stripe k+1 is encoded"""
        c = self.codec
        check(lib().spz_frame_encode_gather_dev(self.comm.handle, self.dst, c.predictor, c.elem_sz, x.data_ptr(), self.total_len,
                                                c.ndims, c.chunk_rows, self.stripe_chunks, c.frame.data_ptr(), c.frame_cap,
                                                c.work.data_ptr(), c.work_bytes,
                                                self.buf.data_ptr() if self.buf is not None else None, self.cap, _stream_ptr()),
              "spz_frame_encode_gather_dev")
        return self._frame(self.stripe_chunks)

    def enable_p2p(self):
        """Collective: allocates the shared (CUDA IPC) regions; the gathered frame of encode_and_gather_p2p lands in the
        destination's one."""
        torch = _torch()
        ptr = C.c_void_p(None)
        need = self.cap if self.comm.rank == self.dst else 0
        check(lib().spz_comm_shared_alloc(self.comm.handle, need, C.byref(ptr)), "spz_comm_shared_alloc")
        self.p2p_buf = None
        if need:
            class _DevMem:                      # CUDA array interface over the library-owned allocation
                pass
            m = _DevMem()
            m.__cuda_array_interface__ = {"shape": (need,), "typestr": "|u1", "data": (int(ptr.value), False), "version": 2}
            self.p2p_buf = torch.as_tensor(m, device=self.codec.frame.device)
        self.p2p = True

    def encode_and_gather_p2p(self, x):
        """streamed over peer memory: every stripe's compaction kernel stores straight into the destination's frame"""
        c = self.codec
        check(lib().spz_frame_encode_gather_p2p_dev(self.comm.handle, self.dst, c.predictor, c.elem_sz, x.data_ptr(),
                                                    self.total_len, c.ndims, c.chunk_rows, self.stripe_chunks, c.frame.data_ptr(),
                                                    c.frame_cap, c.work.data_ptr(), c.work_bytes, _stream_ptr()),
              "spz_frame_encode_gather_p2p_dev")
        if self.comm.rank != self.dst:
            return None
        self.last_stripes = self.stripe_chunks
        self.last_frame = Frame(self.p2p_buf, c.predictor, c.elem_sz, self.total_len, c.ndims, c.chunk_rows, None)
        return self.last_frame

    def shard_positions(self, rank: int, stripes: int):
        """[(global element lo, local element lo, count)] of `rank`'s pieces inside the global stream"""
        c = self.codec
        clen = c.chunk_rows * c.ndims
        if stripes == 0:
            return [(rank * c.len, 0, c.len)]
        S = stripes * clen
        out, k = [], 0
        while k * S < c.len:
            n = min(S, c.len - k * S)
            out.append(((k * self.comm.world + rank) * S, k * S, n))
            k += 1
        return out

    @staticmethod
    def _checksum(t):
        """position-weighted 64-bit checksum of a 1-D integer tensor (wraps; computed in slabs, no large temporaries)"""
        torch = _torch()
        acc = torch.zeros(2, dtype=torch.int64, device=t.device)
        step = 1 << 25
        for a in range(0, t.numel(), step):
            v = t[a: a + step].to(torch.int64)
            w = (torch.arange(a, a + v.numel(), device=t.device, dtype=torch.int64) % 65521) + 1
            acc[0] += v.sum()
            acc[1] += (v * w).sum()
        return acc

    def verify_against_shards(self, gframe, x, all_ranks: bool = False) -> bool:
        """Collective when all_ranks: dst decodes the gathered frame; its own pieces are compared element by element with
        x, every other rank's pieces through position-weighted checksums exchanged over torch.distributed.  Returns the
        verdict on dst (None elsewhere)."""
        torch = _torch()
        out = frame_decode(gframe) if self.comm.rank == self.dst else None
        ok = None
        if out is not None:
            ok = True
            for glo, llo, n in self.shard_positions(self.comm.rank, self.last_stripes):
                ok = ok and bool(torch.equal(out[glo: glo + n], x[llo: llo + n]))
        if all_ranks and self.comm.world > 1:
            import torch.distributed as dist
            stripes = self.last_stripes if self.comm.rank == self.dst else None
            box = [stripes]
            dist.broadcast_object_list(box, src=self.dst)
            stripes = box[0]
            mine = torch.stack([self._checksum(x[llo: llo + n]) for _, llo, n in self.shard_positions(self.comm.rank, stripes)])
            allsums = [torch.zeros_like(mine) for _ in range(self.comm.world)]
            dist.all_gather(allsums, mine)
            if out is not None:
                for r in range(self.comm.world):
                    if r == self.dst:
                        continue
                    for k, (glo, _, n) in enumerate(self.shard_positions(r, stripes)):
                        ok = ok and bool(torch.equal(self._checksum(out[glo: glo + n]), allsums[r][k]))
        del out
        return ok
