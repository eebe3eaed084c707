"""The per-dim code of the stand-alone FIRE transform kernel (sprintz_b200/csrc/spz_xff_forecast.cuh) compiled for the
HOST with g++ and checked against the oracle restatement (which is pinned to the compiled reference in
test_forecast.py).  The CUDA kernel runs exactly this code with one thread per dim, so this pins its arithmetic on the
CPU; the GPU tests (test_zz_xff_forecast_gpu.py) then only have indexing and plumbing left to prove."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import inputs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DIMS = (1, 2, 3, 4, 5, 7, 8, 9, 16, 17, 31, 32, 33, 64, 65, 80)


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("xffp") / "libxffp_host.so")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-o", out,
                           os.path.join(ROOT, "tests", "host", "xffp_host.cpp")])
    L = C.CDLL(out)
    L.host_xffp.restype = C.c_int64
    L.host_xffp.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32]
    return L


@pytest.mark.parametrize("esz", [1, 2], ids=["xff8", "xff16"])
def test_kernel_code_on_host_matches_oracle(host_lib, oracle, esz):
    bad = []
    for D in DIMS:
        for n in inputs.REF_SIZES + [4096 + 3, 50000]:
            fams = inputs.reference_families(n, esz)
            for tag in ("i%64", "fuzz256", "fuzz16", "squares", "zeros", "sparse6"):
                x = np.ascontiguousarray(fams[tag])
                want, _ = oracle.encode_xff_forecast(x, D, write_size=False)
                y = np.zeros_like(x)
                host_lib.host_xffp(1, esz, x.ctypes.data, x.size, y.ctypes.data, D)
                if not np.array_equal(y.view(np.uint8), want):
                    bad.append((D, n, tag, "encode"))
                    continue
                z = np.zeros_like(x)
                host_lib.host_xffp(0, esz, y.ctypes.data, x.size, z.ctypes.data, D)
                if not np.array_equal(z, x):
                    bad.append((D, n, tag, "decode"))
    assert not bad, (len(bad), bad[:5])
