"""Slab-pipelined frame encode (spz_set_encode_pipeline): slab k's directory scan + compaction run on a second stream
while slab k+1 is being encoded.  The frame bytes must not depend on the slab count."""
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
    yield sprintz_b200
    sprintz_b200.set_encode_pipeline(0)
    sprintz_b200.set_kernel_path(0)


@pytest.mark.parametrize("force", [0, 2], ids=["auto", "lane"])
@pytest.mark.parametrize("pred", [DELTA, XFF], ids=["delta", "xff"])
@pytest.mark.parametrize("esz,D", [(1, 8), (2, 16), (1, 3), (1, 80), (2, 5)])
def test_pipelined_frame_encode_matches_oracle(spz, oracle, pred, esz, D, force):
    """Every chunk against the oracle, exact directory, round trip -- with 2, 3, 7 and 64 slabs forced on small frames
    (ragged last slab, ragged last chunk, single-chunk frames, run slots), on the team and on the lane kernels."""
    spz.set_kernel_path(force)
    for slabs in (2, 3, 7, 64):
        spz.set_encode_pipeline(slabs)
        for chunk_rows, nrows, ragged in ((16, 16 * 700 + 5, 0), (64, 64 * 300, 3), (256, 5000, 0), (256, 100, 0)):
            x = inputs.ar1(nrows, D, esz, seed=nrows)
            if ragged:
                x = x[:-ragged]
            _frame_vs_oracle(spz, oracle, pred, esz, D, x, chunk_rows)
            nch = (x.size + chunk_rows * D - 1) // (chunk_rows * D)
            per = -(-nch // slabs)
            assert spz.last_encode_slabs() == -(-nch // per), (slabs, nch, spz.last_encode_slabs())
        _frame_vs_oracle(spz, oracle, pred, esz, D, inputs.random_walk_with_runs(3000, D, esz, seed=5), 32)
        _frame_vs_oracle(spz, oracle, pred, esz, D, np.zeros(4096 * D, np.uint8 if esz == 1 else np.uint16), 16)


def test_slab_count_does_not_change_the_frame(spz):
    import torch
    from sprintz_b200.synth import ar1_stream
    x = ar1_stream(1 << 18, 16, 2, seed=9, device="cuda")
    frames = []
    for slabs in (1, 2, 5, 16, 0):
        spz.set_encode_pipeline(slabs)
        fr = spz.frame_encode(XFF, x, 16, 256)
        torch.cuda.synchronize()
        frames.append(fr.buf[: fr.nbytes()].clone())
    for f in frames[1:]:
        assert torch.equal(f, frames[0])


def test_pipelined_long_frame_round_trip(spz):
    """c3-shaped frame of 256 MiB at 256-row chunks in 7 slabs on the production (lane) kernels: same round trip."""
    import torch
    from sprintz_b200.synth import ar1_stream
    spz.set_encode_pipeline(7)
    x = ar1_stream(1 << 23, 16, 2, seed=3, device="cuda")
    codec = spz.FrameCodec(XFF, 2, x.numel(), 16, 256)
    fr = codec.encode(x)
    assert spz.last_encode_slabs() == 7
    out = torch.empty_like(x)
    codec.decode(fr, out)
    torch.cuda.synchronize()
    codec.check_status()
    assert torch.equal(out, x)
    spz.set_encode_pipeline(0)
