"""Non-RLE sibling codecs: compress_rowmajor_delta_{8b,16b} / compress_rowmajor_{8b,16b}
(sprintz_delta.h:27-44; reference tests: test/test_sprintz_delta.cpp:70-165).

CPU: the oracle's restatement against the compiled reference (bytes, return values, both decoders).
GPU: the CUDA path (generic kernels, variant flag) against the oracle through the C ABI."""
import numpy as np
import pytest

import inputs
import parity
from oracle import RM_DELTA, RM_PLAIN, RM_XFF8

KINDS = [(RM_DELTA, 1), (RM_PLAIN, 1), (RM_DELTA, 2), (RM_PLAIN, 2), (RM_XFF8, 1)]
KIDS = ["delta8", "plain8", "delta16", "plain16", "xff8"]
DIMS = (1, 2, 3, 4, 5, 7, 8, 9, 16, 17, 33, 80)


@pytest.mark.parametrize("kind,esz", KINDS, ids=KIDS)
def test_oracle_rowmajor_vs_reference(oracle, ref, kind, esz):
    bad = []
    for D in DIMS:
        for n in inputs.REF_SIZES:
            for tag, x in inputs.reference_families(n, esz).items():
                ws, wr = oracle.compress_rowmajor(kind, x, D)
                rs, rr = ref.compress_rowmajor(kind, x, D)
                msg = parity.compare_stream(ws, wr, rs, rr, esz)
                if msg is None:
                    d, dn = oracle.decompress_rowmajor(kind, esz, ws, x.size)
                    msg = parity.compare_decoded(d, dn, x)
                    rd, rn = ref.decompress_rowmajor(kind, esz, ws, x.size)
                    if msg is None and not (rn == x.size and np.array_equal(rd, x)):
                        msg = "the reference's decoder disagrees on the oracle's stream"
                if msg:
                    bad.append((D, n, tag, msg))
    assert not bad, (len(bad), bad[:5])


@pytest.mark.gpu
@pytest.mark.parametrize("kind,esz", KINDS, ids=KIDS)
def test_gpu_rowmajor_matches_oracle(oracle, kind, esz):
    import sprintz_b200 as spz
    spz.lib()
    bad = []
    for D in DIMS + (129,):
        cases = []
        for n in inputs.REF_SIZES:
            cases += list(inputs.reference_families(n, esz).items())
        for nrows in (24, 1000):
            cases += [(f"{t}/r{nrows}", x) for t, x in inputs.adversarial_families(nrows, D, esz).items()]
        cases.append(("zeros", np.zeros(40 * 16 * D + 5, np.uint8 if esz == 1 else np.uint16)))
        for tag, x in cases:
            ws, wr = oracle.compress_rowmajor(kind, x, D)
            gs, gr = spz.compress_rowmajor(kind, x, D)
            msg = parity.compare_stream(gs, gr, ws, wr, esz)
            if msg is None and len(gs) != len(ws):
                msg = f"true byte length {len(gs)} != {len(ws)}"
            if msg is None:
                dec, dr = spz.decompress_rowmajor(kind, esz, ws)
                msg = parity.compare_decoded(dec, dr, x)
                if msg:
                    msg = "decode: " + msg
            if msg:
                bad.append((D, x.size, tag, msg))
    assert not bad, (len(bad), bad[:5])


@pytest.mark.gpu
def test_gpu_rowmajor_named_entry_points_vs_live_reference(ref):
    import ctypes as C
    import sprintz_b200 as spz
    L = spz.lib()
    for name, kind, esz, D in (("rowmajor_delta_8b", RM_DELTA, 1, 8), ("rowmajor_delta_16b", RM_DELTA, 2, 16),
                               ("rowmajor_8b", RM_PLAIN, 1, 5), ("rowmajor_16b", RM_PLAIN, 2, 3),
                               ("8b_rowmajor_xff", RM_XFF8, 1, 8)):
        x = inputs.random_walk_with_runs(900, D, esz, seed=3)
        rs, rr = ref.compress_rowmajor(kind, x, D)
        out = np.zeros(spz.max_compressed_bytes(esz, x.size, D), np.uint8)
        cname, dname = ("spz_compress" + name, "spz_decompress" + name) if kind == RM_XFF8 else \
                       ("spz_compress_" + name, "spz_decompress_" + name)
        ret = getattr(L, cname)(x.ctypes.data, x.size, out.ctypes.data, D, 1)
        assert ret == rr and np.array_equal(out[: rr * esz], rs[: rr * esz])
        dec = np.zeros(x.size, x.dtype)
        n = getattr(L, dname)(out.ctypes.data, dec.ctypes.data)
        assert n == x.size and np.array_equal(dec, x)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", [RM_DELTA, RM_PLAIN, RM_XFF8], ids=["delta", "plain", "xff8"])
@pytest.mark.parametrize("esz,D", [(1, 8), (2, 16), (1, 3), (1, 80)])
def test_gpu_rowmajor_frames(oracle, kind, esz, D):
    if kind == RM_XFF8 and esz != 1:
        pytest.skip("compress8b_rowmajor_xff is 8-bit only")
    """Frames carry the non-RLE codecs as codec ids 2 / 3: every chunk is that codec's reference stream."""
    import sprintz_b200 as spz
    from test_gpu_frame import _frame_vs_oracle
    spz.lib()
    for chunk_rows, nrows, ragged in ((16, 400, 0), (256, 5000, 3), (256, 100, 0)):
        x = inputs.ar1(nrows, D, esz, seed=nrows)
        if ragged:
            x = x[:-ragged]
        _frame_vs_oracle(spz, oracle, kind + 1, esz, D, x, chunk_rows)
    _frame_vs_oracle(spz, oracle, kind + 1, esz, D, np.zeros(4096 * D, x.dtype), 1024)
    y = inputs.ar1(3000, D, esz, seed=1)
    fr = spz.frame_compress_host(kind + 1, y, D, 256)
    assert np.array_equal(spz.frame_decompress_host(fr), y)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,esz", KINDS, ids=KIDS)
def test_gpu_rowmajor_lane_kernels(oracle, kind, esz):
    """The lane kernels also carry the non-RLE variants (ndims 8 / 16); forced here on small inputs."""
    import sprintz_b200 as spz
    from test_gpu_frame import _frame_vs_oracle
    spz.lib()
    old = spz.set_kernel_path(2)
    try:
        bad = []
        for D in (8, 16):
            cases = []
            for n in inputs.REF_SIZES:
                cases += list(inputs.reference_families(n, esz).items())
            for nrows in (24, 1000):
                cases += [(f"{t}/r{nrows}", x) for t, x in inputs.adversarial_families(nrows, D, esz).items()]
            cases.append(("zeros", np.zeros(40 * 16 * D + 5, np.uint8 if esz == 1 else np.uint16)))
            for tag, x in cases:
                ws, wr = oracle.compress_rowmajor(kind, x, D)
                gs, gr = spz.compress_rowmajor(kind, x, D)
                msg = parity.compare_stream(gs, gr, ws, wr, esz)
                if msg is None and len(gs) != len(ws):
                    msg = f"true byte length {len(gs)} != {len(ws)}"
                if msg is None:
                    dec, dr = spz.decompress_rowmajor(kind, esz, ws)
                    msg = parity.compare_decoded(dec, dr, x)
                if msg:
                    bad.append((D, x.size, tag, msg))
            for chunk_rows, nrows, ragged in ((16, 400, 0), (256, 5000, 3), (64, 64 * 300, 0)):
                x = inputs.ar1(nrows, D, esz, seed=nrows)
                _frame_vs_oracle(spz, oracle, kind + 1, esz, D, x[:-ragged] if ragged else x, chunk_rows)
        assert not bad, (len(bad), bad[:5])
    finally:
        spz.set_kernel_path(old)


def _golden_rowmajor():
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_rowmajor_v1.npz"))
    for k, line in enumerate(z["index"]):
        kind, esz, D, ret, tag = str(line).split("|", 4)
        yield int(kind), int(esz), int(D), int(ret), tag, z[f"x{k}"], z[f"s{k}"]


def test_oracle_rowmajor_golden_vectors(oracle):
    """Committed fixtures produced by the compiled reference (tests/golden/make_golden_rowmajor.py)."""
    bad = []
    for kind, esz, D, ret, tag, x, stream in _golden_rowmajor():
        ws, wr = oracle.compress_rowmajor(kind, x, D)
        msg = parity.compare_stream(ws, wr, stream, ret, esz)
        if msg is None:
            d, dn = oracle.decompress_rowmajor(kind, esz, ws, x.size)
            msg = parity.compare_decoded(d, dn, x)
        if msg:
            bad.append((kind, esz, D, tag, msg))
    assert not bad, (len(bad), bad[:5])


@pytest.mark.gpu
def test_gpu_rowmajor_golden_vectors():
    import sprintz_b200 as spz
    spz.lib()
    bad = []
    for kind, esz, D, ret, tag, x, stream in _golden_rowmajor():
        gs, gr = spz.compress_rowmajor(kind, x, D)
        msg = parity.compare_stream(gs, gr, stream, ret, esz)
        if msg is None:
            d, dn = spz.decompress_rowmajor(kind, esz, gs)
            msg = parity.compare_decoded(d, dn, x)
        if msg:
            bad.append((kind, esz, D, tag, msg))
    assert not bad, (len(bad), bad[:5])
