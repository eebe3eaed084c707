"""The scalar primitives of the CUDA path (spz_common.cuh: zigzag, sign extension, FIRE coefficient / prediction /
gradient) compiled for the HOST with nvcc and checked exhaustively, mirroring the reference's primitive tests
(test/test_bitpack.cpp:21-240: zigzag over all 256 / 65536 values; bitpack.h:302-303, sprintz_xff_rle.cpp:209-245,
sprintz_xff_lowdim.cpp:170-173, util.h:63-68)."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def prim(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    out = str(tmp_path_factory.mktemp("prim") / "libprim_host.so")
    subprocess.check_call([nvcc, "-std=c++17", "-O2", "-shared", "-Xcompiler", "-fPIC", "-o", out,
                           os.path.join(ROOT, "tests", "host", "primitives_host.cu")])
    L = C.CDLL(out)
    for name, args in (("h_zigzag_enc", [C.c_int, C.c_int32]), ("h_zigzag_dec", [C.c_int, C.c_uint32]),
                       ("h_fire_coef", [C.c_int, C.c_int, C.c_int32]), ("h_fire_predict", [C.c_int, C.c_int32, C.c_int32]),
                       ("h_fire_grad", [C.c_int, C.c_int32, C.c_int32]), ("h_sx", [C.c_int, C.c_int32])):
        getattr(L, name).argtypes = args
        getattr(L, name).restype = C.c_uint32 if name == "h_zigzag_enc" else C.c_int32
    L.h_swar.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
    L.h_swar.restype = None
    return L


def _sx(v, bits):
    v &= (1 << bits) - 1
    return v - (1 << bits) if v >> (bits - 1) else v


@pytest.mark.parametrize("w", [8, 16])
def test_zigzag_exhaustive(prim, w):
    lo, hi = -(1 << (w - 1)), 1 << (w - 1)
    seen = set()
    for e in range(lo, hi):
        u = prim.h_zigzag_enc(w, e)
        want = ((e << 1) ^ (e >> (w - 1))) & ((1 << w) - 1)          # bitpack.h:302
        assert u == want and 0 <= u < (1 << w)
        assert prim.h_zigzag_dec(w, u) == e                          # bitpack.h:303
        seen.add(u)
    assert len(seen) == 1 << w                                       # a bijection on w-bit values
    # small magnitudes get small codes: |e| < 2^k  =>  zigzag(e) < 2^(k+1)
    for k in range(w - 1):
        for e in (-(1 << k), (1 << k) - 1):
            assert prim.h_zigzag_enc(w, e) < (1 << (k + 1))


def test_sign_extension(prim):
    for bits in (8, 16):
        for v in list(range(-70000, 70000, 97)) + [0x7fffffff, -0x80000000, (1 << bits) - 1, 1 << (bits - 1)]:
            assert prim.h_sx(bits, v) == _sx(v, bits)
    assert prim.h_sx(32, -5) == -5


@pytest.mark.parametrize("w", [8, 16])
def test_fire_coefficient_prediction_gradient(prim, w):
    cb = 16 if w == 8 else 32
    rng = np.random.default_rng(w)
    counters = [0, 1, -1, 31, 32, -32, -33, (1 << (cb - 1)) - 1, -(1 << (cb - 1))] + [int(v) for v in rng.integers(-(1 << (cb - 1)), 1 << (cb - 1), 4000)]
    for c in counters:
        hi_coef = _sx((c >> (1 + w - 4)) << (w - 4), 16)             # sprintz_xff_rle.cpp:209-217: 4 MSBs kept, int16
        assert prim.h_fire_coef(w, 0, c) == hi_coef
        assert prim.h_fire_coef(w, 1, c) == c >> 1                   # sprintz_xff_lowdim.cpp:170-173
    deltas = [0, 1, -1, (1 << (w - 1)) - 1, -(1 << (w - 1))] + [int(v) for v in rng.integers(-(1 << (w - 1)), 1 << (w - 1), 300)]
    coefs = [0, 16, -16, 255, -256, 4080, -4096, 32767, -32768] + [int(v) for v in rng.integers(-(1 << 15), 1 << 15, 100)]
    for pd in deltas:
        for coef in coefs:
            want = _sx(_sx(pd * coef, 32) >> w, w)                   # sprintz_xff_rle.cpp:225
            assert prim.h_fire_predict(w, pd, coef) == want
        for err in (0, 1, -1, 77, -77):
            want = 0 if err == 0 else (-pd if err < 0 else pd)       # util.h:63-68
            assert prim.h_fire_grad(w, err, pd) == want


@pytest.mark.parametrize("w", [8, 16])
def test_packed_lane_arithmetic(prim, w):
    """spz_swar.cuh (the packed subtract of the delta forecaster kernels, and the per-lane zigzag planned for the packed
    delta encoder): every lane behaves like the scalar W-bit operation, with no borrow / carry / shift leaking between
    lanes -- all values of one lane against extreme neighbours, plus random words."""
    udt = np.uint8 if w == 8 else np.uint16
    nl = 32 // w
    rng = np.random.default_rng(w)
    words = [rng.integers(0, 1 << 32, 200000, dtype=np.uint64).astype(np.uint32) for _ in range(2)]
    vals = np.arange(1 << w, dtype=np.uint32)
    for lane in range(nl):
        for fill in (0, (1 << w) - 1, 1 << (w - 1)):
            base = np.uint32(sum(fill << (w * k) for k in range(nl) if k != lane))
            a = (vals << np.uint32(w * lane)) | base
            if w == 8:                                            # all 65536 pairs of one lane
                aa = np.repeat(a, 256); bb = np.tile(a, 256)
            else:
                aa = a; bb = np.roll(a, 12345)
            words.append(aa.astype(np.uint32)); words.append(bb.astype(np.uint32))
    a = np.ascontiguousarray(np.concatenate(words[0::2])); b = np.ascontiguousarray(np.concatenate(words[1::2]))
    n = min(a.size, b.size); a, b = a[:n].copy(), b[:n].copy()
    al, bl = a.view(udt).astype(np.int64), b.view(udt).astype(np.int64)
    out = np.zeros(n, np.uint32)
    mask = (1 << w) - 1
    want = {0: (al - bl) & mask, 1: (al + bl) & mask,
            2: ((al << 1) ^ np.where(al >> (w - 1), mask, 0)) & mask, 3: ((al >> 1) ^ np.where(al & 1, mask, 0)) & mask}
    for op in (0, 1, 2, 3):
        prim.h_swar(w, op, a.ctypes.data, b.ctypes.data, out.ctypes.data, n)
        assert np.array_equal(out.view(udt).astype(np.int64), want[op]), (w, op)
    # zigzag round trip per lane
    z = np.zeros(n, np.uint32)
    prim.h_swar(w, 2, a.ctypes.data, b.ctypes.data, z.ctypes.data, n)
    prim.h_swar(w, 3, z.ctypes.data, b.ctypes.data, out.ctypes.data, n)
    assert np.array_equal(out, a)
