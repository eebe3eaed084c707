"""CPU tests of the boundary: the C-ABI library loads, exports every symbol include/sprintz_b200.h declares,
and its host-only helpers (sizes, frame geometry, stream measuring) are right.  No compute calls."""
import ctypes as C
import os
import re

import numpy as np

import inputs
from oracle import DELTA, XFF

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "sprintz_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(spz_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import sprintz_b200 as spz
    L = C.CDLL(spz._lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    # the ctypes table and the header agree
    assert sorted(spz._lib.SYMBOLS) == names


def test_library_is_in_tree_and_sm100a():
    import sprintz_b200 as spz
    assert os.path.dirname(spz._lib.LIB_PATH) == os.path.join(ROOT, "sprintz_b200")
    assert b"sm_100a" in spz.lib().spz_version()


def test_no_cpu_fallback_when_library_missing(monkeypatch):
    import sprintz_b200 as spz
    monkeypatch.setattr(spz._lib, "_lib", None)
    monkeypatch.setattr(spz._lib, "LIB_PATH", "/nonexistent/libsprintz_b200.so")
    try:
        spz._lib.lib()
    except spz.SprintzError as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("missing library must raise")


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "sprintz_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("test infrastructure", ""), f"{f} mentions the oracle"


def test_size_helpers():
    import sprintz_b200 as spz
    L = spz.lib()
    for esz in (1, 2):
        for D in (1, 8, 80, 1024):
            for n in (0, 1, 127, 128, 10_000, 1 << 20):
                cap = L.spz_max_compressed_bytes(esz, n, D)
                assert cap % 16 == 0 and cap >= n * esz + 8
    assert L.spz_frame_nchunks(1 << 30, 8, 1024) == (1 << 30) // 8192
    assert L.spz_frame_nchunks(8193, 8, 1024) == 2
    assert L.spz_frame_nchunks(100, 0, 16) == 0
    fm = L.spz_frame_max_bytes(1, 1 << 20, 8, 256)
    assert fm > (1 << 20)
    assert L.spz_frame_workspace_bytes(1, 1 << 20, 8, 256) > (1 << 20)


def test_measure_stream_matches_oracle(oracle):
    """spz_measure_stream is the only host-side stream logic in the product (it sizes the H2D copy)."""
    import sprintz_b200 as spz
    from sprintz_b200.api import measure
    for pred in (DELTA, XFF):
        for esz in (1, 2):
            for D in (1, 2, 3, 5, 8, 33):
                for tag, x in inputs.adversarial_families(64, D, esz).items():
                    s, r = oracle.compress(pred, x, D)
                    padded = np.concatenate([s, np.zeros(16, np.uint8)])
                    nbytes, nelems = measure(esz, padded)
                    assert nbytes == len(s), (pred, esz, D, tag)
                    assert nelems == x.size, (pred, esz, D, tag)
                x = np.zeros(1000 * D, np.uint8 if esz == 1 else np.uint16)
                s, r = oracle.compress(pred, x, D)
                nbytes, nelems = measure(esz, np.concatenate([s, np.zeros(16, np.uint8)]))
                assert (nbytes, nelems) == (len(s), x.size)


def test_reference_named_wrappers_exist():
    import sprintz_b200 as spz
    for w in ("8b", "16b"):
        for p in ("delta", "xff"):
            assert callable(getattr(spz, f"sprintz_compress_{p}_{w}"))
            assert callable(getattr(spz, f"sprintz_decompress_{p}_{w}"))


def test_shard_plan_partitions_the_stream():
    """spz_shard_len (the partition the multi-GPU gathers use, spz_comm.cu make_plan) on the CPU: contiguous ranges and
    round-robin stripes cover the global stream exactly once, only the last global stripe is short, and the lengths
    agree with a Python model of the rule (tests/mgpu_worker.py uses the same model on the GPU box)."""
    import sprintz_b200 as spz
    L = spz.lib()

    def model(total, D, chunk_rows, S, world, rank):
        clen = chunk_rows * D
        nch = max(1, -(-total // clen))
        if S == 0:
            base, extra = divmod(nch, world)
            lo = sum(base + (1 if r < extra else 0) for r in range(rank))
            n = base + (1 if rank < extra else 0)
            return max(0, min(total, (lo + n) * clen) - min(total, lo * clen))
        nstripes = -(-nch // S)
        return sum(min(total, (j + 1) * S * clen) - j * S * clen for j in range(rank, nstripes, world))

    for total, D, chunk_rows in ((1 << 20, 8, 512), (1000003, 5, 16), (80 * 64 * 37 + 16, 80, 64), (17, 16, 512)):
        for world in (1, 2, 3, 8):
            for S in (0, 1, 7, 64):
                lens = [int(L.spz_shard_len(total, D, chunk_rows, S, world, r)) for r in range(world)]
                nch = max(1, -(-total // (chunk_rows * D)))
                if S and -(-(-(-nch // S)) // world) > 62:              # more than 62 stripes per rank: not a supported plan
                    assert lens == [0] * world
                    continue
                assert sum(lens) == total, (total, D, chunk_rows, S, world, lens)
                assert lens == [model(total, D, chunk_rows, S, world, r) for r in range(world)]
    assert L.spz_shard_len(100, 0, 16, 0, 2, 0) == 0 and L.spz_shard_len(100, 4, 16, 0, 2, 5) == 0
