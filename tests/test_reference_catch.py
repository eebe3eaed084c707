"""The drop-in boundary, compiled: integration/sprintz_b200_shim.cpp gives the reference's sprintz.h entry points
(cpp/Compress/sprintz.h:16-32) bodies that forward to libsprintz_b200.so, and integration/Makefile links the
reference's UNMODIFIED Catch runner (cpp/Compress/main.cpp + test/test_sprintz_8b.cpp:24-64) against it.
CPU: the runner loads and lists the reference's two test cases.  GPU: it runs them -- the reference's own
round-trip suite (ndims 1..129 x its sizes x its input families, compress_testing.hpp:518-543) on the CUDA path."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUNNER = os.path.join(ROOT, "oracle", "_ref", "catch_sprintz_b200")


def _need_runner():
    if not os.path.exists(RUNNER):
        pytest.skip("oracle/_ref/catch_sprintz_b200 not built (make -C integration catch needs /root/reference)")


def test_catch_runner_links_against_the_library():
    _need_runner()
    r = subprocess.run([RUNNER, "--list-tests"], capture_output=True, text=True, timeout=120)
    assert "sprintz 8b delta" in r.stdout and "sprintz 8b xff" in r.stdout, r.stdout + r.stderr
    ldd = subprocess.run(["ldd", RUNNER], capture_output=True, text=True).stdout
    assert "libsprintz_b200.so" in ldd and "not found" not in ldd
    # the runner carries none of the reference's codec code: its sprintz_* symbols come from the shim
    nm = subprocess.run(["nm", "-C", RUNNER], capture_output=True, text=True).stdout
    assert "compress_rowmajor_delta_rle" not in nm and "spz_compress_delta_8b" in nm


@pytest.mark.gpu
def test_reference_catch_suite_passes_on_the_gpu_path():
    _need_runner()
    r = subprocess.run([RUNNER, "[sprintz]"], capture_output=True, text=True, timeout=1500)
    tail = r.stdout[-600:]
    assert r.returncode == 0 and "All tests passed" in tail, tail + r.stderr[-600:]
    assert "2 test cases" in tail
