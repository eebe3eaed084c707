"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU checkers for the Sprintz hot path:

* ``Oracle``  : our plain-C restatement (oracle/sprintz_oracle.c -> oracle/liboracle.so)
* ``Ref``     : the UNMODIFIED reference compiled from /root/reference by oracle/Makefile
                (oracle/_ref/libsprintz_ref.so; prebuilt, travels to the GPU box)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may
import this package.  The product package (sprintz_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libsprintz_ref.so")

DELTA, XFF = 0, 1
RM_DELTA, RM_PLAIN, RM_XFF8 = 1, 2, 3   # non-RLE sibling codecs: compress_rowmajor_delta_*,
                                        # compress_rowmajor_*, compress8b_rowmajor_xff (8-bit only)
SLACK = 4096  # the reference reads/writes past buffer ends; give every buffer slack


def build(ref: bool = True) -> None:
    """(Re)build liboracle.so and, if /root/reference is present, oracle/_ref."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if ref:
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def _as_u8(a: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(a).view(np.uint8).reshape(-1)


class _Codec:
    """numpy-level wrapper shared by Oracle and Ref. Streams are returned as exact-length uint8 arrays."""

    name = "?"

    def max_bytes(self, esz: int, n: int, ndims: int) -> int:
        d = max(1, ndims)
        hb = 3 if esz == 1 else 4
        H = (2 * d * hb + 7) // 8
        return 8 + n * esz + (n // (16 * d) + 2) * H + 16

    def compress(self, predictor: int, x: np.ndarray, ndims: int, write_size: bool = True):
        raise NotImplementedError

    def decompress(self, predictor: int, esz: int, stream: np.ndarray, n_out: int):
        raise NotImplementedError


class Oracle(_Codec):
    name = "oracle"

    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        L = C.CDLL(ORACLE_SO)
        L.oracle_compress.restype = C.c_int64
        L.oracle_compress.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_uint32, C.c_void_p,
                                      C.c_uint16, C.c_int, C.POINTER(C.c_uint64)]
        L.oracle_decompress.restype = C.c_int64
        L.oracle_decompress.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.oracle_decompress_raw.restype = C.c_int64
        L.oracle_decompress_raw.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                            C.c_uint16, C.c_uint32, C.c_uint16]
        L.oracle_compress_rowmajor.restype = C.c_int64
        L.oracle_compress_rowmajor.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_uint32, C.c_void_p,
                                               C.c_uint16, C.c_int, C.POINTER(C.c_uint64)]
        L.oracle_decompress_rowmajor.restype = C.c_int64
        L.oracle_decompress_rowmajor.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.oracle_encode_forecast.restype = C.c_int64
        L.oracle_encode_forecast.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint16, C.c_int]
        L.oracle_decode_forecast.restype = C.c_int64
        L.oracle_decode_forecast.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint16]
        L.oracle_encode_xff_forecast.restype = C.c_int64
        L.oracle_encode_xff_forecast.argtypes = [C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint16, C.c_int]
        L.oracle_decode_xff_forecast.restype = C.c_int64
        L.oracle_decode_xff_forecast.argtypes = [C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint16]
        self.L = L

    def encode_xff_forecast(self, x, ndims, write_size=True):
        """Stand-alone FIRE transform (predict.h). -> (bytes incl. metadata, ret)"""
        esz = x.dtype.itemsize
        src = _as_u8(x)
        buf = np.zeros(x.size * esz + 6 + SLACK, np.uint8)
        ret = int(self.L.oracle_encode_xff_forecast(esz, src.ctypes.data, x.size, buf.ctypes.data, ndims, int(write_size)))
        return buf[: ret * esz].copy(), ret

    def decode_xff_forecast(self, esz, y, ndims):
        y = np.ascontiguousarray(y)
        n = y.size
        src = np.concatenate([_as_u8(y), np.zeros(SLACK, np.uint8)])
        out = np.full(n * esz + SLACK, 0xAA, np.uint8)
        ret = int(self.L.oracle_decode_xff_forecast(esz, src.ctypes.data, n, out.ctypes.data, ndims))
        return out[: n * esz].view(np.uint8 if esz == 1 else np.uint16).copy(), ret

    def encode_forecast(self, order, x, ndims, write_size=True):
        """Stand-alone forecaster transform (order 1 = delta, 2 = double delta). -> (bytes incl. metadata, ret)"""
        esz = x.dtype.itemsize
        src = _as_u8(x)
        buf = np.zeros(x.size * esz + 6 + SLACK, np.uint8)
        ret = int(self.L.oracle_encode_forecast(order, esz, src.ctypes.data, x.size, buf.ctypes.data, ndims, int(write_size)))
        return buf[: ret * esz].copy(), ret

    def decode_forecast(self, order, esz, y, ndims):
        """y: transformed elements without metadata. -> (decoded elements, ret)"""
        y = np.ascontiguousarray(y)
        n = y.size
        src = np.concatenate([_as_u8(y), np.zeros(SLACK, np.uint8)])
        out = np.full(n * esz + SLACK, 0xAA, np.uint8)
        ret = int(self.L.oracle_decode_forecast(order, esz, src.ctypes.data, n, out.ctypes.data, ndims))
        return out[: n * esz].view(np.uint8 if esz == 1 else np.uint16).copy(), ret

    def compress_rowmajor(self, kind, x, ndims, write_size=True):
        esz = x.dtype.itemsize
        src = _as_u8(x)
        buf = np.zeros(self.max_bytes(esz, x.size, ndims) + SLACK, np.uint8)
        nb = C.c_uint64(0)
        ret = self.L.oracle_compress_rowmajor(kind, esz, src.ctypes.data, x.size, buf.ctypes.data, ndims,
                                              int(write_size), C.byref(nb))
        return buf[: nb.value].copy(), int(ret)

    def decompress_rowmajor(self, kind, esz, stream, n_out):
        src = np.concatenate([_as_u8(stream), np.zeros(SLACK, np.uint8)])
        out = np.full((n_out + 64) * esz + SLACK, 0xAA, np.uint8)
        ret = self.L.oracle_decompress_rowmajor(kind, esz, src.ctypes.data, out.ctypes.data)
        dt = np.uint8 if esz == 1 else np.uint16
        return out[: n_out * esz].view(dt).copy(), int(ret)

    def compress(self, predictor, x, ndims, write_size=True):
        """-> (stream bytes [exact true length], returned element count)"""
        if predictor >= 2:                       # frame codec ids 2 / 3: the non-RLE sibling codecs
            return self.compress_rowmajor(predictor - 1, x, ndims, write_size)
        esz = x.dtype.itemsize
        src = _as_u8(x)
        n = x.size
        buf = np.zeros(self.max_bytes(esz, n, ndims) + SLACK, np.uint8)
        nb = C.c_uint64(0)
        ret = self.L.oracle_compress(predictor, esz, src.ctypes.data, n, buf.ctypes.data,
                                     ndims, int(write_size), C.byref(nb))
        return buf[: nb.value].copy(), int(ret)

    def decompress(self, predictor, esz, stream, n_out):
        src = np.concatenate([_as_u8(stream), np.zeros(SLACK, np.uint8)])
        out = np.full((n_out + 64) * esz + SLACK, 0xAA, np.uint8)
        ret = self.L.oracle_decompress(predictor, esz, src.ctypes.data, out.ctypes.data)
        dt = np.uint8 if esz == 1 else np.uint16
        return out[: n_out * esz].view(dt).copy(), int(ret)

    # ---- entropy stage (oracle/huff_oracle.c): OUR container, parity unpinned
    def huff_lengths(self, counts):
        cnt = np.ascontiguousarray(counts, dtype=np.uint64)
        out = np.zeros(256, np.uint8)
        self.L.oracle_huff_lengths.restype = None
        self.L.oracle_huff_lengths.argtypes = [C.c_void_p, C.c_void_p]
        self.L.oracle_huff_lengths(cnt.ctypes.data, out.ctypes.data)
        return out

    def huff_encode(self, x):
        src = np.ascontiguousarray(_as_u8(x))
        self.L.oracle_huff_max_bytes.restype = C.c_uint64
        self.L.oracle_huff_max_bytes.argtypes = [C.c_uint64]
        self.L.oracle_huff_encode.restype = C.c_int64
        self.L.oracle_huff_encode.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        buf = np.zeros(int(self.L.oracle_huff_max_bytes(src.size)), np.uint8)
        n = self.L.oracle_huff_encode(src.ctypes.data, src.size, buf.ctypes.data)
        return buf[: int(n)].copy()

    def huff_decode(self, cont, n):
        src = np.ascontiguousarray(_as_u8(cont))
        self.L.oracle_huff_decode.restype = C.c_int64
        self.L.oracle_huff_decode.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
        out = np.zeros(max(n, 1), np.uint8)
        r = self.L.oracle_huff_decode(src.ctypes.data, src.size, out.ctypes.data, n)
        return out[:n], int(r)

    def compress_bs(self, predictor, x, ndims, block_rows, write_size=True):
        """Block-size axis (BASELINE config 5): block_rows in {8, 16, 32, 64}; 8 is oracle_compress itself."""
        self.L.oracle_compress_bs.restype = C.c_int64
        self.L.oracle_compress_bs.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint16, C.c_int,
                                              C.c_void_p, C.c_int]
        esz = x.dtype.itemsize
        src = _as_u8(x)
        buf = np.zeros(self.max_bytes(esz, x.size, ndims) + SLACK, np.uint8)
        nb = C.c_uint64(0)
        ret = self.L.oracle_compress_bs(predictor, esz, src.ctypes.data, x.size, buf.ctypes.data, ndims, int(write_size),
                                        C.byref(nb), block_rows)
        return buf[: nb.value].copy(), int(ret)

    def decompress_bs(self, predictor, esz, stream, n_out, block_rows):
        self.L.oracle_decompress_bs.restype = C.c_int64
        self.L.oracle_decompress_bs.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        src = np.concatenate([_as_u8(stream), np.zeros(SLACK, np.uint8)])
        out = np.full((n_out + 64) * esz + SLACK, 0xAA, np.uint8)
        ret = self.L.oracle_decompress_bs(predictor, esz, src.ctypes.data, out.ctypes.data, block_rows)
        dt = np.uint8 if esz == 1 else np.uint16
        return out[: n_out * esz].view(dt).copy(), int(ret)

    def query(self, predictor, esz, stream, n_out, op):
        """-> (result uint64[ndims], decoded elements, returned count)"""
        self.L.oracle_query.restype = C.c_int64
        self.L.oracle_query.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        src = np.concatenate([_as_u8(stream), np.zeros(SLACK, np.uint8)])
        ndims = int(src[6]) | (int(src[7]) << 8)
        out = np.full((n_out + 64) * esz + SLACK, 0xAA, np.uint8)
        res = np.zeros(max(ndims, 1), np.uint64)
        ret = self.L.oracle_query(predictor, esz, src.ctypes.data, out.ctypes.data, op, res.ctypes.data)
        dt = np.uint8 if esz == 1 else np.uint16
        return res, out[: n_out * esz].view(dt).copy(), int(ret)

    def decompress_raw(self, predictor, esz, stream, n_out, ndims, ngroups, rem):
        src = np.concatenate([_as_u8(stream), np.zeros(SLACK, np.uint8)])
        out = np.full((n_out + 64) * esz + SLACK, 0xAA, np.uint8)
        ret = self.L.oracle_decompress_raw(predictor, esz, src.ctypes.data, out.ctypes.data,
                                           ndims, ngroups, rem)
        dt = np.uint8 if esz == 1 else np.uint16
        return out[: n_out * esz].view(dt).copy(), int(ret)


class Ref(_Codec):
    name = "reference"

    def __init__(self):
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(
                f"{REF_SO} missing: run `make -C oracle ref` where /root/reference exists")
        L = C.CDLL(REF_SO)
        L.ref_compress.restype = C.c_int64
        L.ref_compress.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_uint32, C.c_void_p,
                                   C.c_uint16, C.c_int]
        L.ref_decompress.restype = C.c_int64
        L.ref_decompress.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.ref_compress_chunks.restype = C.c_double
        L.ref_compress_chunks.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.c_uint64,
                                          C.c_uint16, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int]
        L.ref_decompress_chunks.restype = C.c_double
        L.ref_decompress_chunks.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.c_uint64,
                                            C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int]
        if hasattr(L, "ref_compress_rowmajor"):
            L.ref_compress_rowmajor.restype = C.c_int64
            L.ref_compress_rowmajor.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint16, C.c_int]
            L.ref_decompress_rowmajor.restype = C.c_int64
            L.ref_decompress_rowmajor.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        if hasattr(L, "ref_encode_forecast"):
            L.ref_encode_forecast.restype = C.c_int64
            L.ref_encode_forecast.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint16, C.c_int]
            L.ref_decode_forecast.restype = C.c_int64
            L.ref_decode_forecast.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint16]
        self.L = L

    def encode_xff_forecast(self, x, ndims, write_size=True):
        esz = x.dtype.itemsize
        self.L.ref_encode_xff_forecast.restype = C.c_int64
        self.L.ref_encode_xff_forecast.argtypes = [C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint16, C.c_int]
        # the reference's vector loads / stores run up to 32 bytes past a row and it reads ndims elements before the input
        pad = SLACK + 4 * max(1, ndims) * esz
        src = np.concatenate([np.zeros(pad, np.uint8), _as_u8(x), np.zeros(2 * SLACK, np.uint8)])
        buf = np.zeros(x.size * esz + 6 + 4 * SLACK, np.uint8)
        ret = int(self.L.ref_encode_xff_forecast(esz, src.ctypes.data + pad, x.size, buf.ctypes.data, ndims, int(write_size)))
        return buf[: ret * esz].copy(), ret

    def decode_xff_forecast(self, esz, y, ndims):
        self.L.ref_decode_xff_forecast.restype = C.c_int64
        self.L.ref_decode_xff_forecast.argtypes = [C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint16]
        y = np.ascontiguousarray(y)
        n = y.size
        src = np.concatenate([_as_u8(y), np.zeros(2 * SLACK, np.uint8)])
        pad = SLACK + 4 * max(1, ndims) * esz
        out = np.full(pad + n * esz + 4 * SLACK, 0, np.uint8)
        ret = int(self.L.ref_decode_xff_forecast(esz, src.ctypes.data, n, out.ctypes.data + pad, ndims))
        return out[pad: pad + n * esz].view(np.uint8 if esz == 1 else np.uint16).copy(), ret

    def encode_forecast(self, order, x, ndims, write_size=True):
        esz = x.dtype.itemsize
        # the reference reads up to ndims elements BEFORE the input (delta.cpp:112-115) and vector-width past it
        pad = SLACK + 4 * max(1, ndims) * esz
        src = np.concatenate([np.zeros(pad, np.uint8), _as_u8(x), np.zeros(SLACK, np.uint8)])
        buf = np.zeros(x.size * esz + 6 + 2 * SLACK, np.uint8)
        ret = int(self.L.ref_encode_forecast(order, esz, src.ctypes.data + pad, x.size, buf.ctypes.data, ndims, int(write_size)))
        return buf[: ret * esz].copy(), ret

    def decode_forecast(self, order, esz, y, ndims):
        y = np.ascontiguousarray(y)
        n = y.size
        src = np.concatenate([_as_u8(y), np.zeros(SLACK, np.uint8)])
        pad = SLACK + 4 * max(1, ndims) * esz
        out = np.full(pad + n * esz + 2 * SLACK, 0, np.uint8)
        ret = int(self.L.ref_decode_forecast(order, esz, src.ctypes.data, n, out.ctypes.data + pad, ndims))
        return out[pad: pad + n * esz].view(np.uint8 if esz == 1 else np.uint16).copy(), ret

    def compress_rowmajor(self, kind, x, ndims, write_size=True):
        esz = x.dtype.itemsize
        src = np.concatenate([_as_u8(x), np.zeros(SLACK, np.uint8)])
        buf = np.zeros(self.max_bytes(esz, x.size, ndims) + SLACK, np.uint8)
        ret = int(self.L.ref_compress_rowmajor(kind, esz, src.ctypes.data, x.size, buf.ctypes.data, ndims, int(write_size)))
        if ret < 0:
            return buf[:0].copy(), ret
        return buf[: ret * esz + (esz - 1)].copy(), ret

    def decompress_rowmajor(self, kind, esz, stream, n_out):
        src = np.concatenate([_as_u8(stream), np.zeros(SLACK, np.uint8)])
        out = np.full((n_out + 64) * esz + SLACK, 0xAA, np.uint8)
        ret = self.L.ref_decompress_rowmajor(kind, esz, src.ctypes.data, out.ctypes.data)
        dt = np.uint8 if esz == 1 else np.uint16
        return out[: n_out * esz].view(dt).copy(), int(ret)

    @staticmethod
    def available() -> bool:
        return os.path.exists(REF_SO)

    def compress(self, predictor, x, ndims, write_size=True):
        """-> (stream bytes [ret*esz bytes, + 1 possibly-dropped byte for 16b], returned element count).

        The reference reports floor(bytes/2) for 16-bit streams, so the buffer is
        pre-zeroed and we return ret*esz + (esz-1) bytes; callers compare the first ret*esz bytes
        and treat the extra byte separately."""
        esz = x.dtype.itemsize
        n = x.size
        src = np.concatenate([_as_u8(x), np.zeros(SLACK, np.uint8)])
        buf = np.zeros(self.max_bytes(esz, n, ndims) + SLACK, np.uint8)
        ret = int(self.L.ref_compress(predictor, esz, src.ctypes.data, n, buf.ctypes.data,
                                      ndims, int(write_size)))
        if ret < 0:
            return buf[:0].copy(), ret
        return buf[: ret * esz + (esz - 1)].copy(), ret

    def decompress(self, predictor, esz, stream, n_out):
        src = np.concatenate([_as_u8(stream), np.zeros(SLACK, np.uint8)])
        out = np.full((n_out + 64) * esz + SLACK, 0xAA, np.uint8)
        ret = self.L.ref_decompress(predictor, esz, src.ctypes.data, out.ctypes.data)
        dt = np.uint8 if esz == 1 else np.uint16
        return out[: n_out * esz].view(dt).copy(), int(ret)
