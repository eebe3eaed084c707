"""GPU parity tests of the lane-per-octet kernels (spz_decode_lane.cu / spz_encode_lane.cu).

In production they only run for frames with many chunks; spz_set_kernel_path(2) forces them for every
eligible shape (ndims 8 / 16) so the same matrices that pin the team kernels pin them too, including
single streams, ragged tails, run slots and the reference's golden vectors."""
import numpy as np
import pytest

import inputs
import parity
from oracle import DELTA, XFF
from test_gpu_frame import _frame_vs_oracle
from test_gpu_parity import CODECS, IDS, _check_case

pytestmark = pytest.mark.gpu

LANE_DIMS = (8, 16)


@pytest.fixture(scope="module")
def spz():
    import sprintz_b200
    sprintz_b200.lib()
    old = sprintz_b200.set_kernel_path(2)
    yield sprintz_b200
    sprintz_b200.set_kernel_path(old)


@pytest.mark.parametrize("pred,esz", CODECS, ids=IDS)
def test_lane_reference_matrix(spz, oracle, pred, esz):
    bad = []
    for D in LANE_DIMS:
        for n in inputs.REF_SIZES:
            for tag, x in inputs.reference_families(n, esz).items():
                _check_case(spz, oracle, pred, esz, D, x, bad, tag)
    assert not bad, (len(bad), bad[:5])


@pytest.mark.parametrize("pred,esz", CODECS, ids=IDS)
def test_lane_adversarial_and_runs(spz, oracle, pred, esz):
    udt = np.uint8 if esz == 1 else np.uint16
    bad = []
    for D in LANE_DIMS:
        for nrows in (24, 100, 1000):
            for tag, x in inputs.adversarial_families(nrows, D, esz).items():
                _check_case(spz, oracle, pred, esz, D, x, bad, f"{tag}/r{nrows}")
            x = inputs.random_walk_with_runs(nrows, D, esz, seed=9)[:-3]
            _check_case(spz, oracle, pred, esz, D, x, bad, "ragged")
        B = 8 * D
        for k in (2, 3, 18):                                   # RLE termination edges
            for j in range(0, 2 * B + 1, 3):
                L = k * B + j
                _check_case(spz, oracle, pred, esz, D, np.zeros(L, udt), bad, "zeros")
                _check_case(spz, oracle, pred, esz, D, inputs.const_then_data(L, esz, (k - 1) * B), bad, "const+data")
        for nblocks in (0x7f, 0x80, 0x7fff, 0x8000):           # run-length byte edges
            x = np.full((nblocks + 3) * B, 9, udt)
            x[-2 * B:] = (np.arange(2 * B) % 251).astype(udt)
            _check_case(spz, oracle, pred, esz, D, x, bad, f"const{nblocks}+data")
    fams = {}
    n = 256 * 1024 + 7
    fams.update(inputs.fuzz(n, esz)); fams.update(inputs.sparse(n, esz))
    for D in LANE_DIMS:
        for tag in ("fuzz256", "fuzz16", "fuzz2", "sparse6", "sparse2"):
            _check_case(spz, oracle, pred, esz, D, fams[tag], bad, tag)
    assert not bad, (len(bad), bad[:5])


@pytest.mark.parametrize("pred", [DELTA, XFF], ids=["delta", "xff"])
@pytest.mark.parametrize("esz,D", [(1, 8), (1, 16), (2, 8), (2, 16)])
def test_lane_frames_match_oracle(spz, oracle, pred, esz, D):
    for chunk_rows, nrows, ragged in ((16, 400, 0), (256, 5000, 3), (4096, 20000, 0), (256, 256, 0), (256, 100, 0),
                                      (64, 64 * 300, 0)):
        x = inputs.ar1(nrows, D, esz, seed=nrows)
        if ragged:
            x = x[:-ragged]
        _frame_vs_oracle(spz, oracle, pred, esz, D, x, chunk_rows)
    x = inputs.random_walk_with_runs(3000, D, esz, seed=5)
    _frame_vs_oracle(spz, oracle, pred, esz, D, x, 512)
    _frame_vs_oracle(spz, oracle, pred, esz, D, np.zeros(4096 * D, x.dtype), 1024)


def test_lane_golden_vectors(spz, golden):
    bad = []
    for c in golden.cases:
        if c["D"] not in LANE_DIMS:
            continue
        gs, gr = spz.compress(c["pred"], c["x"], c["D"])
        msg = parity.compare_stream(gs, gr, c["stream"], c["ret"], c["esz"])
        if msg is None:
            dec, dr = spz.decompress(c["pred"], c["esz"], gs)
            msg = parity.compare_decoded(dec, dr, c["x"])
        if msg:
            bad.append((c["pred"], c["esz"], c["D"], c["tag"], msg))
    assert not bad, (len(bad), bad[:5])


# ------------------------------------------------------------------ low-dim lane decoder (spz_decode_lowlane.cu)

LOW_DIMS = {1: (1, 2, 3, 4), 2: (1, 2)}


@pytest.mark.parametrize("pred,esz", CODECS, ids=IDS)
def test_lowlane_reference_matrix_and_runs(spz, oracle, pred, esz):
    """ndims <= 4 (8-bit) / <= 2 (16-bit): column-major payload, flat header, FIRE coef = counter >> 1.  Single
    streams through the forced lane path: the reference's size x family matrix, RLE termination and run-length byte
    edges, adversarial families."""
    udt = np.uint8 if esz == 1 else np.uint16
    bad = []
    for D in LOW_DIMS[esz]:
        for n in inputs.REF_SIZES:
            for tag, x in inputs.reference_families(n, esz).items():
                _check_case(spz, oracle, pred, esz, D, x, bad, tag)
        assert spz.lib().spz_last_kernel_family(1) == 2, "low-dim lane decoder did not run"
        assert spz.lib().spz_last_kernel_family(0) == 2, "low-dim lane encoder did not run"
        for nrows in (24, 100, 1000):
            for tag, x in inputs.adversarial_families(nrows, D, esz).items():
                _check_case(spz, oracle, pred, esz, D, x, bad, f"{tag}/r{nrows}")
            x = inputs.random_walk_with_runs(nrows, D, esz, seed=9)
            _check_case(spz, oracle, pred, esz, D, x[:-1] if D > 1 else x[:-3], bad, "ragged")
        B = 8 * D
        for k in (2, 17, 18):                                  # RLE termination edges
            for j in range(0, 2 * B + 1):
                L = k * B + j
                _check_case(spz, oracle, pred, esz, D, np.zeros(L, udt), bad, "zeros")
                _check_case(spz, oracle, pred, esz, D, inputs.const_then_data(L, esz, (k - 1) * B), bad, "const+data")
        for nblocks in (0x7f, 0x80, 0x7fff, 0x8000):           # run-length byte edges
            x = np.full((nblocks + 3) * B, 9, udt)
            x[-2 * B:] = (np.arange(2 * B) % 251).astype(udt)
            _check_case(spz, oracle, pred, esz, D, x, bad, f"const{nblocks}+data")
        n = 256 * 1024 + 7
        fams = {}
        fams.update(inputs.fuzz(n, esz)); fams.update(inputs.sparse(n, esz))
        for tag in ("fuzz256", "fuzz16", "fuzz2", "sparse6", "sparse2"):
            _check_case(spz, oracle, pred, esz, D, fams[tag], bad, tag)
        # a long steep ramp: the low-dim FIRE counter is 32 bits wide for 16-bit data and its product wraps
        _check_case(spz, oracle, pred, esz, D, inputs.linear_ramp(60000, D, esz, slope=997), bad, "ramp")
        _check_case(spz, oracle, pred, esz, D, inputs.ramps_then_const(3000, D, esz), bad, "ramps_then_const")
    assert not bad, (len(bad), bad[:5])


@pytest.mark.parametrize("pred", [DELTA, XFF], ids=["delta", "xff"])
@pytest.mark.parametrize("esz,D", [(1, 1), (1, 2), (1, 3), (1, 4), (2, 1), (2, 2)])
def test_lowlane_frames_match_oracle(spz, oracle, pred, esz, D):
    for chunk_rows, nrows, ragged in ((16, 400, 0), (256, 5000, 3), (4096, 20000, 0), (256, 100, 0), (64, 64 * 300, 0)):
        x = inputs.ar1(nrows, D, esz, seed=nrows)
        if ragged:
            x = x[:-ragged]
        _frame_vs_oracle(spz, oracle, pred, esz, D, x, chunk_rows)
    assert spz.lib().spz_last_kernel_family(1) == 2, "low-dim lane decoder did not run"
    assert spz.lib().spz_last_kernel_family(0) == 2, "low-dim lane encoder did not run"
    _frame_vs_oracle(spz, oracle, pred, esz, D, inputs.random_walk_with_runs(3000, D, esz, seed=5), 512)
    _frame_vs_oracle(spz, oracle, pred, esz, D, np.zeros(4096 * D, x.dtype), 1024)
    _frame_vs_oracle(spz, oracle, pred, esz, D, inputs.linear_ramp(70000, D, esz, slope=997), 35008)


def test_lowlane_golden_vectors(spz, golden):
    bad = []
    for c in golden.cases:
        if c["D"] not in LOW_DIMS[c["esz"]]:
            continue
        gs, gr = spz.compress(c["pred"], c["x"], c["D"])
        msg = parity.compare_stream(gs, gr, c["stream"], c["ret"], c["esz"])
        if msg is None:
            dec, dr = spz.decompress(c["pred"], c["esz"], c["stream"])
            msg = parity.compare_decoded(dec, dr, c["x"])
        if msg:
            bad.append((c["pred"], c["esz"], c["D"], c["tag"], msg))
    assert not bad, (len(bad), bad[:5])
