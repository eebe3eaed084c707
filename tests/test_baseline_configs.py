"""BASELINE.json's configs as parity cases.  configs[0] -- "int8, 1-dim, delta+bitpack, 8-sample blocks, 64 KB stream on
CPU (cpp/Compress reference path)" -- is the reference's own CPU-runnable case: the oracle is pinned to the compiled
reference on exactly that stream (CPU), and the CUDA path must reproduce the reference's bytes through the
reference-named entry point and as a frame (GPU).  configs[1..3] at their full sizes live in
test_gpu_frame.py::test_full_size_properties; the ndims axis of configs[4] in test_frame_many_chunks_every_team_shape."""
import numpy as np
import pytest

import inputs
import parity
from oracle import DELTA, XFF


def _c1_stream():
    return inputs.ar1(64 << 10, 1, 1, seed=12345)          # 64 KB, int8, 1 dim


@pytest.mark.parametrize("pred", [DELTA, XFF], ids=["delta", "xff"])
def test_config1_oracle_equals_compiled_reference(oracle, ref, pred):
    x = _c1_stream()
    ws, wr = oracle.compress(pred, x, 1)
    rs, rr = ref.compress(pred, x, 1)
    assert parity.compare_stream(ws, wr, rs, rr, 1) is None
    assert len(ws) < x.size                                  # it does compress (AR(1), sigma 2: ~2x)
    dec, n = oracle.decompress(pred, 1, rs, x.size)
    rdec, rn = ref.decompress(pred, 1, ws, x.size)
    assert n == rn == x.size and np.array_equal(dec, x) and np.array_equal(rdec, x)


@pytest.mark.gpu
@pytest.mark.parametrize("pred,name", [(DELTA, "delta_8b"), (XFF, "xff_8b")], ids=["delta", "xff"])
def test_config1_gpu_drop_in_equals_compiled_reference(ref, pred, name):
    import sprintz_b200 as spz
    L = spz.lib()
    x = _c1_stream()
    rs, rr = ref.compress(pred, x, 1)
    out = np.zeros(int(L.spz_max_compressed_bytes(1, x.size, 1)), np.uint8)
    ret = getattr(L, "spz_compress_" + name)(x.ctypes.data, x.size, out.ctypes.data, 1, 1)
    assert ret == rr and np.array_equal(out[:rr], rs[:rr])
    dec = np.zeros(x.size, np.uint8)
    n = getattr(L, "spz_decompress_" + name)(out.ctypes.data, dec.ctypes.data)
    assert n == x.size and np.array_equal(dec, x)
    # the reference decodes our bytes
    rdec, rn = ref.decompress(pred, 1, out[:rr], x.size)
    assert rn == x.size and np.array_equal(rdec, x)


@pytest.mark.gpu
@pytest.mark.parametrize("pred", [DELTA, XFF], ids=["delta", "xff"])
@pytest.mark.parametrize("path", [0, 2], ids=["auto", "lane"])
def test_config1_gpu_frame_chunks_equal_compiled_reference(ref, pred, path):
    import torch
    import sprintz_b200 as spz
    spz.lib()
    x = _c1_stream()
    old = spz.set_kernel_path(path)
    try:
        fr = spz.frame_encode(pred, torch.from_numpy(x).cuda(), 1, 2048)
        out = spz.frame_decode(fr)
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), x)
        off, nb, po = fr.directory()
        buf = fr.buf[: fr.nbytes()].cpu().numpy()
        assert fr.nchunks == 32
        for c in range(fr.nchunks):
            rs, rr = ref.compress(pred, x[c * 2048: (c + 1) * 2048], 1)
            gs = buf[po + off[c]: po + off[c] + nb[c]]
            assert len(gs) == rr and np.array_equal(gs, rs[:rr]), f"chunk {c}"
    finally:
        spz.set_kernel_path(old)
