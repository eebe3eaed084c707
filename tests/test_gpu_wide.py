"""GPU parity tests of the wide-row kernels (spz_encode_wide.cu / spz_decode_wide.cu): 8- and 16-bit delta and FIRE
streams whose ndims is a multiple of 8 in 24..256 (BASELINE config 4 is 80 dims).  Every call goes through the C ABI; the checker is
the oracle (pinned to the compiled reference).  The tests assert that the wide family really ran."""
import numpy as np
import pytest

import inputs
from oracle import DELTA, XFF
from test_gpu_frame import _frame_vs_oracle
from test_gpu_parity import _check_case

pytestmark = pytest.mark.gpu

# partial octets / unaligned rows (the PART form of the encoder): any ndims >= 5
PART_DIMS = (5, 6, 7, 9, 12, 15, 17, 20, 30, 33, 100, 255)
WIDE_DIMS = PART_DIMS + (16, 24, 32, 40, 56, 64, 80, 96, 128, 200, 256)
FAMILY_WIDE = 3


@pytest.fixture(scope="module")
def spz():
    import sprintz_b200
    sprintz_b200.lib()
    return sprintz_b200


def _ran_wide(spz, D=80, esz=1):
    L = spz.lib()
    if D % 8:                                                 # partial octets: the encoder only (team decoders measured faster)
        assert L.spz_last_kernel_family(0) == FAMILY_WIDE, "wide-row encoder did not run"
        return
    if D == 16:                                               # decoded by the lane kernels; 16-bit frames are encoded by them too
        assert esz == 2 or L.spz_last_kernel_family(0) == FAMILY_WIDE, "wide-row encoder did not run"
        return
    assert L.spz_last_kernel_family(0) == FAMILY_WIDE, "wide-row encoder did not run"
    if D >= 32 and (D, esz) != (32, 2):                       # 24 dims / 16-bit x 32 dims decode faster on the team kernel
        assert L.spz_last_kernel_family(1) == FAMILY_WIDE, "wide-row decoder did not run"


def _dt(esz):
    return np.uint8 if esz == 1 else np.uint16


def _both(f):
    f = pytest.mark.parametrize("pred", (DELTA, XFF), ids=("delta", "fire"))(f)
    return pytest.mark.parametrize("esz", (1, 2), ids=("u8", "u16"))(f)


PREDICTORS = _both


@PREDICTORS
def test_wide_reference_matrix(spz, oracle, pred, esz):
    bad = []
    for D in WIDE_DIMS:
        for n in inputs.REF_SIZES:
            for tag, x in inputs.reference_families(n, esz).items():
                _check_case(spz, oracle, pred, esz, D, x, bad, tag)
        _ran_wide(spz, D, esz)
    assert not bad, (len(bad), bad[:5])


@PREDICTORS
def test_wide_adversarial_and_runs(spz, oracle, pred, esz):
    bad = []
    for D in WIDE_DIMS:
        for nrows in (24, 100, 1000):
            for tag, x in inputs.adversarial_families(nrows, D, esz).items():
                _check_case(spz, oracle, pred, esz, D, x, bad, f"{tag}/r{nrows}")
            x = inputs.random_walk_with_runs(nrows, D, esz, seed=9)[:-3]
            _check_case(spz, oracle, pred, esz, D, x, bad, "ragged")
        # every field width, including single dims that differ from their neighbours (pack constants, header bytes)
        rng = np.random.default_rng(D)
        nb = rng.integers(0, 8 * esz + 1, size=D)
        amp = np.where(nb > 0, (1 << np.maximum(nb - 1, 0)) - 1, 0)
        steps = rng.integers(-1, 2, size=(600, D)) * amp[None, :]
        steps[::7, :] = -amp[None, :] - (nb > 0)
        x = ((1 << (8 * esz - 1)) + np.cumsum(steps, axis=0)) & ((1 << (8 * esz)) - 1)
        _check_case(spz, oracle, pred, esz, D, x.astype(_dt(esz)).reshape(-1), bad, "mixed widths")
    for D in (7, 24, 33, 80, 256):
        B = 8 * D
        for k in (2, 3, 18):                                   # RLE termination edges
            for j in list(range(0, 2 * B + 1, 37)) + [B - 1, B, B + 1, 2 * B - 1, 2 * B]:
                L = k * B + j
                _check_case(spz, oracle, pred, esz, D, np.zeros(L, _dt(esz)), bad, "zeros")
                _check_case(spz, oracle, pred, esz, D, inputs.const_then_data(L, esz, (k - 1) * B), bad, "const+data")
    for D in (6, 24, 80, 100):
        B = 8 * D
        for nblocks in (0x7f, 0x80, 0x7fff, 0x8000):           # run-length byte edges
            x = np.full((nblocks + 3) * B, 9, _dt(esz))
            x[-2 * B:] = ((np.arange(2 * B) * 2654435761) % (251 if esz == 1 else 65521)).astype(_dt(esz))
            _check_case(spz, oracle, pred, esz, D, x, bad, f"const{nblocks}+data")
    fams = {}
    n = 256 * 1024 + 7
    fams.update(inputs.fuzz(n, esz)); fams.update(inputs.sparse(n, esz))
    for D in (5, 24, 30, 80, 256):
        for tag in ("fuzz256", "fuzz16", "fuzz2", "sparse6", "sparse2"):
            _check_case(spz, oracle, pred, esz, D, fams[tag], bad, tag)
    assert not bad, (len(bad), bad[:5])


@PREDICTORS
@pytest.mark.parametrize("D", WIDE_DIMS)
def test_wide_frames_match_oracle(spz, oracle, D, pred, esz):
    """Frames: many teams per CTA, idle lanes (ndims/8 does not divide 32), ragged last chunks, chunks in runs."""
    for chunk_rows, nrows, ragged in ((16, 400, 0), (256, 5000, 3), (512, 6000, 0), (256, 256, 0), (256, 100, 0),
                                      (64, 64 * 150, 0)):
        x = inputs.ar1(nrows, D, esz, seed=nrows)
        if ragged:
            x = x[:-ragged]
        _frame_vs_oracle(spz, oracle, pred, esz, D, x, chunk_rows)
        if x.size % D == 0:
            _ran_wide(spz, D, esz)
    x = inputs.random_walk_with_runs(3000, D, esz, seed=5)
    _frame_vs_oracle(spz, oracle, pred, esz, D, x, 512)
    _frame_vs_oracle(spz, oracle, pred, esz, D, np.zeros(4096 * D, _dt(esz)), 1024)
    # chunks that differ wildly in length of work: constant chunks next to noisy ones
    rng = np.random.default_rng(D)
    x = inputs.ar1(4096, D, esz, seed=3).reshape(4096, D).copy()
    for c in range(0, 4096 // 128, 3):
        x[c * 128:(c + 1) * 128] = rng.integers(0, 1 << (8 * esz), size=D)
    _frame_vs_oracle(spz, oracle, pred, esz, D, x.reshape(-1), 128)
