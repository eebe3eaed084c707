"""EXPERIMENTAL path: overlapped directory scan + compaction (spz_chase_kernel, opt-in via spz_set_encode_overlap).
Kept in its own late-sorting file so that the experimental switch cannot stop the main parity suites under -x."""
import numpy as np
import pytest

import inputs
from oracle import DELTA, XFF
from test_gpu_frame import _frame_vs_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def spz():
    import sprintz_b200
    sprintz_b200.lib()
    old = sprintz_b200.set_kernel_path(2)
    yield sprintz_b200
    sprintz_b200.set_kernel_path(old)


# ------------------------------------------------------------------ overlapped directory + compaction (spz_chase_kernel)

@pytest.mark.parametrize("pred", [DELTA, XFF], ids=["delta", "xff"])
@pytest.mark.parametrize("esz,D", [(1, 8), (2, 8), (2, 16), (1, 1), (1, 3), (2, 2)])
def test_overlapped_frame_encode_matches_oracle(spz, oracle, pred, esz, D):
    """The chase kernel (scan + copy on a second stream while the lane encoders run) must produce the same frames:
    every chunk against the oracle, exact directory, round trip -- forced on small frames, several tiles of 128
    chunks, ragged last tile, single-chunk frames, run slots."""
    import torch
    old = spz.set_encode_overlap(1)
    try:
        # the first launch of an encoder instantiation takes the classic path (lazy kernel load): warm it
        w = inputs.ar1(64, D, esz, seed=1)
        spz.frame_encode(pred, torch.from_numpy(w.view(np.uint8 if esz == 1 else np.int16)).cuda(), D, 16)
        torch.cuda.synchronize()
        for chunk_rows, nrows, ragged in ((16, 16 * 700 + 5, 0), (64, 64 * 300, 3), (256, 5000, 0), (4096, 20000, 0), (256, 100, 0)):
            x = inputs.ar1(nrows, D, esz, seed=nrows)
            if ragged:
                x = x[:-ragged]
            _frame_vs_oracle(spz, oracle, pred, esz, D, x, chunk_rows)
            assert spz.last_encode_overlapped(), "the overlapped path did not run"
        _frame_vs_oracle(spz, oracle, pred, esz, D, inputs.random_walk_with_runs(3000, D, esz, seed=5), 32)
        _frame_vs_oracle(spz, oracle, pred, esz, D, np.zeros(4096 * D, np.uint8 if esz == 1 else np.uint16), 16)
    finally:
        spz.set_encode_overlap(old)
