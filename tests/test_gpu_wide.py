"""GPU parity tests of the wide-row SWAR kernels (spz_encode_wide.cu / spz_decode_wide.cu): 8-bit delta and FIRE streams
whose ndims is a multiple of 8 in 24..256 (BASELINE config 4 is 80 dims).  Every call goes through the C ABI; the checker is
the oracle (pinned to the compiled reference).  The tests assert that the wide family really ran."""
import numpy as np
import pytest

import inputs
from oracle import DELTA, XFF
from test_gpu_frame import _frame_vs_oracle
from test_gpu_parity import _check_case

pytestmark = pytest.mark.gpu

WIDE_DIMS = (24, 32, 40, 56, 64, 80, 96, 128, 200, 256)
FAMILY_WIDE = 3


@pytest.fixture(scope="module")
def spz():
    import sprintz_b200
    sprintz_b200.lib()
    return sprintz_b200


def _ran_wide(spz):
    L = spz.lib()
    assert L.spz_last_kernel_family(0) == FAMILY_WIDE, "wide-row encoder did not run"
    assert L.spz_last_kernel_family(1) == FAMILY_WIDE, "wide-row decoder did not run"


PREDICTORS = pytest.mark.parametrize("pred", (DELTA, XFF), ids=("delta", "fire"))


@PREDICTORS
def test_wide_reference_matrix(spz, oracle, pred):
    bad = []
    for D in WIDE_DIMS:
        for n in inputs.REF_SIZES:
            for tag, x in inputs.reference_families(n, 1).items():
                _check_case(spz, oracle, pred, 1, D, x, bad, tag)
        _ran_wide(spz)
    assert not bad, (len(bad), bad[:5])


@PREDICTORS
def test_wide_adversarial_and_runs(spz, oracle, pred):
    bad = []
    for D in WIDE_DIMS:
        for nrows in (24, 100, 1000):
            for tag, x in inputs.adversarial_families(nrows, D, 1).items():
                _check_case(spz, oracle, pred, 1, D, x, bad, f"{tag}/r{nrows}")
            x = inputs.random_walk_with_runs(nrows, D, 1, seed=9)[:-3]
            _check_case(spz, oracle, pred, 1, D, x, bad, "ragged")
        # every field width, including single dims that differ from their neighbours (pack constants, header bytes)
        rng = np.random.default_rng(D)
        nb = rng.integers(0, 9, size=D)
        amp = np.where(nb > 0, (1 << np.maximum(nb - 1, 0)) - 1, 0)
        steps = rng.integers(-1, 2, size=(600, D)) * amp[None, :]
        steps[::7, :] = -amp[None, :] - (nb > 0)
        x = (128 + np.cumsum(steps, axis=0)) & 0xff
        _check_case(spz, oracle, pred, 1, D, x.astype(np.uint8).reshape(-1), bad, "mixed widths")
    for D in (24, 80, 256):
        B = 8 * D
        for k in (2, 3, 18):                                   # RLE termination edges
            for j in list(range(0, 2 * B + 1, 37)) + [B - 1, B, B + 1, 2 * B - 1, 2 * B]:
                L = k * B + j
                _check_case(spz, oracle, pred, 1, D, np.zeros(L, np.uint8), bad, "zeros")
                _check_case(spz, oracle, pred, 1, D, inputs.const_then_data(L, 1, (k - 1) * B), bad, "const+data")
    for D in (24, 80):
        B = 8 * D
        for nblocks in (0x7f, 0x80, 0x7fff, 0x8000):           # run-length byte edges
            x = np.full((nblocks + 3) * B, 9, np.uint8)
            x[-2 * B:] = (np.arange(2 * B) % 251).astype(np.uint8)
            _check_case(spz, oracle, pred, 1, D, x, bad, f"const{nblocks}+data")
    fams = {}
    n = 256 * 1024 + 7
    fams.update(inputs.fuzz(n, 1)); fams.update(inputs.sparse(n, 1))
    for D in (24, 80, 256):
        for tag in ("fuzz256", "fuzz16", "fuzz2", "sparse6", "sparse2"):
            _check_case(spz, oracle, pred, 1, D, fams[tag], bad, tag)
    assert not bad, (len(bad), bad[:5])


@PREDICTORS
@pytest.mark.parametrize("D", WIDE_DIMS)
def test_wide_frames_match_oracle(spz, oracle, D, pred):
    """Frames: many teams per CTA, idle lanes (ndims/8 does not divide 32), ragged last chunks, chunks in runs."""
    for chunk_rows, nrows, ragged in ((16, 400, 0), (256, 5000, 3), (512, 6000, 0), (256, 256, 0), (256, 100, 0),
                                      (64, 64 * 150, 0)):
        x = inputs.ar1(nrows, D, 1, seed=nrows)
        if ragged:
            x = x[:-ragged]
        _frame_vs_oracle(spz, oracle, pred, 1, D, x, chunk_rows)
        if x.size % D == 0:
            _ran_wide(spz)
    x = inputs.random_walk_with_runs(3000, D, 1, seed=5)
    _frame_vs_oracle(spz, oracle, pred, 1, D, x, 512)
    _frame_vs_oracle(spz, oracle, pred, 1, D, np.zeros(4096 * D, np.uint8), 1024)
    # chunks that differ wildly in length of work: constant chunks next to noisy ones
    rng = np.random.default_rng(D)
    x = inputs.ar1(4096, D, 1, seed=3).reshape(4096, D).copy()
    for c in range(0, 4096 // 128, 3):
        x[c * 128:(c + 1) * 128] = rng.integers(0, 256, size=D)
    _frame_vs_oracle(spz, oracle, pred, 1, D, x.reshape(-1), 128)
