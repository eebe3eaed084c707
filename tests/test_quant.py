"""float -> integer quantiser: python/datasets/compress_bench.py:45-60 (_quantize, axis = 0).

Fixtures: tests/golden/golden_quant_v1.npz, produced by executing the reference's own function text
(tests/golden/make_golden_quant.py).  CPU: a numpy restatement of the arithmetic the kernel performs equals the fixtures
(so the kernel's recipe is right before a GPU sees it).  GPU: spz_quantize_host / spz_quantize_dev equal the fixtures."""
import os

import numpy as np
import pytest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_quant_v1.npz")


def _cases():
    z = np.load(G)
    return [(z[f"x{k}"], z[f"y{k}"]) for k in range(int(z["n"]))]


def _restated(x, odt):
    """what spz_quant.cu computes, in numpy: shift in the input precision, float32 from there on, truncation"""
    mn = x.min(axis=0, keepdims=True)
    y = (x - mn).astype(np.float32)
    top = y.max(axis=0, keepdims=True)
    den = np.where(top > 1, top, np.float32(1)).astype(np.float32)
    scale = np.float32(255 if odt == np.uint8 else 65535)
    return ((y / den).astype(np.float32) * scale).astype(odt)


def test_restatement_matches_reference_fixtures():
    bad = 0
    for x, want in _cases():
        got = _restated(x, want.dtype.type)
        bad += int(not np.array_equal(got, want))
    assert bad == 0


@pytest.mark.gpu
def test_gpu_quantiser_matches_reference_fixtures():
    import torch
    import sprintz_b200 as spz
    spz.lib()
    bad = []
    for k, (x, want) in enumerate(_cases()):
        rows, D = x.shape
        got = spz.quantize(x.reshape(-1), D, want.dtype.type)
        if not np.array_equal(got.reshape(rows, D), want):
            bad.append((k, "host", x.dtype, want.dtype, x.shape))
        xt = torch.from_numpy(x.reshape(-1).copy()).cuda()
        gd = spz.quantize_dev(xt, D, want.dtype.itemsize).cpu().numpy().view(want.dtype)
        if not np.array_equal(gd.reshape(rows, D), want):
            bad.append((k, "dev", x.dtype, want.dtype, x.shape))
    assert not bad, (len(bad), bad[:5])


@pytest.mark.gpu
def test_gpu_quantiser_feeds_the_codec():
    """the quantised stream is what the codecs eat: quantise a smooth float matrix, compress, decompress, compare"""
    import sprintz_b200 as spz
    rng = np.random.default_rng(2)
    t = np.arange(20000)[:, None]
    x = (np.sin(t / 300.0 + np.arange(8)[None, :]) * 50 + rng.normal(0, 0.5, size=(20000, 8))).astype(np.float32)
    q = spz.quantize(x.reshape(-1), 8, np.uint8)
    assert np.array_equal(q.reshape(20000, 8), _restated(x, np.uint8))
    fr = spz.frame_compress_host(spz.XFF, q, 8, 512)
    assert len(fr) < 0.8 * q.size
    assert np.array_equal(spz.frame_decompress_host(fr), q)
