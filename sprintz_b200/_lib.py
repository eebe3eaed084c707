"""ctypes binding of libsprintz_b200.so (the C ABI declared in include/sprintz_b200.h).

The product path has NO fallback: if the CUDA library is missing or a call fails, this module raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsprintz_b200.so")
CSRC = os.path.join(HERE, "csrc")

DELTA, XFF = 0, 1

ERRORS = {-1: "SPZ_E_ARG", -2: "SPZ_E_CUDA", -3: "SPZ_E_CAPACITY", -4: "SPZ_E_STREAM", -5: "SPZ_E_UNSUPPORTED"}


class SprintzError(RuntimeError):
    pass


class ChunkMeta(C.Structure):
    _fields_ = [("nbytes", C.c_uint32), ("ngroups", C.c_uint32), ("remaining", C.c_uint32), ("status", C.c_uint32)]


class FrameHeader(C.Structure):
    _fields_ = [("magic", C.c_uint32), ("version", C.c_uint16), ("elem_sz", C.c_uint8), ("predictor", C.c_uint8),
                ("ndims", C.c_uint32), ("chunk_rows", C.c_uint32), ("len", C.c_uint64), ("nchunks", C.c_uint64),
                ("payload_off", C.c_uint64), ("payload_bytes", C.c_uint64), ("frame_bytes", C.c_uint64),
                ("reserved", C.c_uint64)]


# every symbol include/sprintz_b200.h declares: name -> (restype, argtypes)
_vp, _i, _u16, _u32, _u64, _i64 = C.c_void_p, C.c_int, C.c_uint16, C.c_uint32, C.c_uint64, C.c_int64
SYMBOLS = {
    "spz_version": (C.c_char_p, []),
    "spz_last_error": (C.c_char_p, []),
    "spz_compress_delta_8b": (_i64, [_vp, _u32, _vp, _u16, _i]),
    "spz_decompress_delta_8b": (_i64, [_vp, _vp]),
    "spz_compress_xff_8b": (_i64, [_vp, _u32, _vp, _u16, _i]),
    "spz_decompress_xff_8b": (_i64, [_vp, _vp]),
    "spz_compress_delta_16b": (_i64, [_vp, _u32, _vp, _u16, _i]),
    "spz_decompress_delta_16b": (_i64, [_vp, _vp]),
    "spz_compress_xff_16b": (_i64, [_vp, _u32, _vp, _u16, _i]),
    "spz_decompress_xff_16b": (_i64, [_vp, _vp]),
    "spz_compress_rowmajor_delta_8b": (_i64, [_vp, _u32, _vp, _u16, _i]),
    "spz_decompress_rowmajor_delta_8b": (_i64, [_vp, _vp]),
    "spz_compress_rowmajor_delta_16b": (_i64, [_vp, _u32, _vp, _u16, _i]),
    "spz_decompress_rowmajor_delta_16b": (_i64, [_vp, _vp]),
    "spz_compress_rowmajor_8b": (_i64, [_vp, _u32, _vp, _u16, _i]),
    "spz_decompress_rowmajor_8b": (_i64, [_vp, _vp]),
    "spz_compress_rowmajor_16b": (_i64, [_vp, _u32, _vp, _u16, _i]),
    "spz_decompress_rowmajor_16b": (_i64, [_vp, _vp]),
    "spz_compress8b_rowmajor_xff": (_i64, [_vp, _u64, _vp, _u16, _i]),
    "spz_decompress8b_rowmajor_xff": (_i64, [_vp, _vp]),
    "spz_compress_rowmajor_host": (_i64, [_i, _i, _vp, _u32, _vp, _u16, _i, C.POINTER(_u64)]),
    "spz_decompress_rowmajor_host": (_i64, [_i, _i, _vp, _vp]),
    "spz_encode_delta_rowmajor_8b": (_u32, [_vp, _u32, _vp, _u16, _i]),
    "spz_decode_delta_rowmajor_8b": (_u32, [_vp, _u32, _vp, _u16]),
    "spz_encode_delta_rowmajor_16b": (_u32, [_vp, _u32, _vp, _u16, _i]),
    "spz_decode_delta_rowmajor_16b": (_u32, [_vp, _u32, _vp, _u16]),
    "spz_encode_doubledelta_rowmajor_8b": (_u32, [_vp, _u32, _vp, _u16, _i]),
    "spz_decode_doubledelta_rowmajor_8b": (_u32, [_vp, _u32, _vp, _u16]),
    "spz_encode_doubledelta_rowmajor_16b": (_u32, [_vp, _u32, _vp, _u16, _i]),
    "spz_decode_doubledelta_rowmajor_16b": (_u32, [_vp, _u32, _vp, _u16]),
    "spz_encode_xff_rowmajor_8b": (_u32, [_vp, _u32, _vp, _u16, _i]),
    "spz_decode_xff_rowmajor_8b": (_u32, [_vp, _u32, _vp, _u16]),
    "spz_encode_xff_rowmajor_16b": (_u32, [_vp, _u32, _vp, _u16, _i]),
    "spz_decode_xff_rowmajor_16b": (_u32, [_vp, _u32, _vp, _u16]),
    "spz_xff_forecast_host": (_i64, [_i, _i, _vp, _u32, _vp, _u16, _i]),
    "spz_xff_forecast_dev": (_i, [_i, _i, _vp, _u64, _u16, _vp, _vp]),
    "spz_encode_forecast_host": (_i64, [_i, _i, _vp, _u32, _vp, _u16, _i]),
    "spz_decode_forecast_host": (_i64, [_i, _i, _vp, _u32, _vp, _u16]),
    "spz_forecast_workspace_bytes": (_u64, [_u64, _u16]),
    "spz_forecast_encode_dev": (_i, [_i, _i, _vp, _u64, _u16, _vp, _vp]),
    "spz_forecast_decode_dev": (_i, [_i, _i, _vp, _u64, _u16, _vp, _vp, _u64, _vp]),
    "spz_query_rowmajor_delta_rle_8b": (_i64, [_vp, _vp, _i, _i, _vp]),
    "spz_query_rowmajor_delta_rle_16b": (_i64, [_vp, _vp, _i, _i, _vp]),
    "spz_query_rowmajor_xff_rle_8b": (_i64, [_vp, _vp, _i, _i, _vp]),
    "spz_query_rowmajor_xff_rle_16b": (_i64, [_vp, _vp, _i, _i, _vp]),
    "spz_query_host": (_i64, [_i, _i, _vp, _vp, _i, _i, _vp]),
    "spz_frame_query_workspace_bytes": (_u64, [_i, _u64, _i]),
    "spz_frame_query_dev": (_i, [_i, _i, _i, _i, _vp, _u64, _u64, _u16, _u32, _vp, _u64, _vp, _vp, _vp, _u64, _vp]),
    "spz_last_query_fused": (_i, []),
    "spz_reduce_rows_dev": (_i, [_i, _i, _vp, _u64, _u16, _vp, _vp]),
    "spz_set_encode_pipeline": (_i, [_i]),
    "spz_last_encode_slabs": (_i, []),
    "spz_kernel_launches": (_u64, []),
    "spz_comm_unique_id": (_i, [_vp]),
    "spz_comm_init": (_i, [C.POINTER(_vp), _i, _i, _vp]),
    "spz_comm_destroy": (_i, [_vp]),
    "spz_shard_len": (_u64, [_u64, _u16, _u32, _u32, _i, _i]),
    "spz_frame_gather_max_bytes": (_u64, [_i, _u64, _u16, _u32]),
    "spz_frame_gather_dev": (_i, [_vp, _i, _u64, _u32, _vp, _vp, _u64, _vp]),
    "spz_frame_encode_gather_dev": (_i, [_vp, _i, _i, _i, _vp, _u64, _u16, _u32, _u32, _vp, _u64, _vp, _u64, _vp, _u64, _vp]),
    "spz_huff_max_bytes": (_u64, [_u64]),
    "spz_huff_workspace_bytes": (_u64, [_u64]),
    "spz_huff_encode_dev": (_i, [_vp, _u64, _vp, _u64, _vp, _u64, _vp, _vp]),
    "spz_huff_decode_dev": (_i, [_vp, _u64, _u64, _vp, _u64, _vp, _u64, _vp, _vp]),
    "spz_comm_shared_alloc": (_i, [_vp, _u64, C.POINTER(_vp)]),
    "spz_frame_encode_gather_p2p_dev": (_i, [_vp, _i, _i, _i, _vp, _u64, _u16, _u32, _u32, _vp, _u64, _vp, _u64, _vp]),
    "spz_quantize_workspace_bytes": (_u64, [_u16]),
    "spz_quantize_dev": (_i, [_i, _i, _vp, _u64, _u16, _vp, _vp, _u64, _vp]),
    "spz_quantize_host": (_i64, [_i, _i, _vp, _u64, _u16, _vp]),
    "spz_compress_host": (_i64, [_i, _i, _vp, _u32, _vp, _u16, _i, C.POINTER(_u64)]),
    "spz_decompress_host": (_i64, [_i, _i, _vp, _vp]),
    "spz_decompress_raw_host": (_i64, [_i, _i, _vp, _vp, _u16, _u32, _u16]),
    "spz_max_compressed_bytes": (_u64, [_i, _u64, _u16]),
    "spz_measure_stream": (_i, [_i, _vp, _i, _u16, _u32, _u16, C.POINTER(_u64), C.POINTER(_u64)]),
    "spz_compress_dev": (_i, [_i, _i, _vp, _u32, _u16, _i, _vp, _vp, _vp]),
    "spz_decompress_dev": (_i, [_i, _i, _vp, _u64, _i, _u16, _vp, _u64, _vp, _vp]),
    "spz_frame_nchunks": (_u64, [_u64, _u16, _u32]),
    "spz_frame_max_bytes": (_u64, [_i, _u64, _u16, _u32]),
    "spz_frame_workspace_bytes": (_u64, [_i, _u64, _u16, _u32]),
    "spz_frame_encode_dev": (_i, [_i, _i, _vp, _u64, _u16, _u32, _vp, _u64, _vp, _u64, _vp, _vp]),
    "spz_frame_decode_dev": (_i, [_i, _i, _vp, _u64, _u64, _u16, _u32, _vp, _u64, _vp, _vp]),
    "spz_compress_bs_dev": (_i, [_i, _i, _i, _vp, _u32, _u16, _i, _vp, _vp, _vp]),
    "spz_decompress_bs_dev": (_i, [_i, _i, _i, _vp, _u64, _i, _u16, _vp, _u64, _vp, _vp]),
    "spz_frame_encode_bs_dev": (_i, [_i, _i, _i, _vp, _u64, _u16, _u32, _vp, _u64, _vp, _u64, _vp, _vp]),
    "spz_frame_decode_bs_dev": (_i, [_i, _i, _i, _vp, _u64, _u64, _u16, _u32, _vp, _u64, _vp, _vp]),
    "spz_frame_read_header": (_i, [_vp, C.POINTER(FrameHeader), _vp]),
    "spz_frame_compress_host": (_i64, [_i, _i, _vp, _u64, _u16, _u32, _vp, _u64]),
    "spz_frame_decompress_host": (_i64, [_vp, _u64, _vp, _u64]),
    "spz_profile_enable": (_i, [_i]),
    "spz_profile_collect": (_i, [C.POINTER(C.c_float), C.POINTER(_i)]),
    "spz_set_kernel_path": (_i, [_i]),
    "spz_last_kernel_family": (_i, [_i]),
    "spz_release_workspaces": (None, []),
}

_lib = None


def build(verbose: bool = False) -> str:
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    out = subprocess.run(["make", "-s", "-j8", "-C", CSRC], capture_output=True, text=True)
    if out.returncode != 0:
        raise SprintzError("building libsprintz_b200.so failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)
    return LIB_PATH


def lib() -> C.CDLL:
    """The loaded library; raises (never falls back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SprintzError(f"{LIB_PATH} not built: run `make -C sprintz_b200/csrc` or __graft_entry__.build(); "
                               "there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)          # AttributeError if the header and the library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int, what: str) -> int:
    if rc < 0:
        msg = ERRORS.get(int(rc), str(rc))
        detail = lib().spz_last_error().decode()
        raise SprintzError(f"{what}: {msg} {detail}")
    return rc
