"""Hardening: streams, frames and directories are input from outside.  Truncated, bit-flipped and mislabelled inputs
must come back as an error status / SPZ_E_STREAM -- never a wild read (the CUDA context must survive: every test ends
with a successful round trip) and never a byte written past the destination capacity.
The reference has no bounds checks at all; these are properties of this implementation."""
import ctypes as C

import numpy as np
import pytest

import inputs
from oracle import DELTA, XFF

pytestmark = pytest.mark.gpu

SHAPES = [(1, 8), (2, 16), (1, 80), (1, 2), (2, 5), (1, 300)]


@pytest.fixture(scope="module")
def spz():
    import sprintz_b200
    sprintz_b200.lib()
    yield sprintz_b200
    sprintz_b200.set_kernel_path(0)


def _alive(spz):
    import torch
    torch.cuda.synchronize()
    x = inputs.ar1(64, 8, 1, seed=1)
    s, r = spz.compress(XFF, x, 8)
    d, n = spz.decompress(XFF, 1, s)
    assert np.array_equal(d, x)


def _decode_frame_raw(spz, fr, buf, dest_elems=None, guard=4096):
    """decode a (possibly corrupted) frame image `buf` (CUDA uint8 tensor) -> (rc, status array, out, guard intact)"""
    import torch
    n = fr.len if dest_elems is None else dest_elems
    dt = torch.uint8 if fr.elem_sz == 1 else torch.int16
    out = torch.full((n + guard,), 0x55 if fr.elem_sz == 1 else 0x5555, dtype=dt, device=buf.device)
    status = torch.zeros(4 * fr.nchunks, dtype=torch.int32, device=buf.device)
    rc = spz.lib().spz_frame_decode_dev(fr.predictor, fr.elem_sz, buf.data_ptr(), buf.numel(), fr.len, fr.ndims, fr.chunk_rows,
                                       out.data_ptr(), n, status.data_ptr(), None)
    torch.cuda.synchronize()
    st = status.view(-1, 4)[:, 3].cpu().numpy()
    intact = bool((out[n:] == (0x55 if fr.elem_sz == 1 else 0x5555)).all().item())
    return rc, st, out, intact


@pytest.mark.parametrize("force", [0, 2], ids=["auto", "lane"])
@pytest.mark.parametrize("pred", [DELTA, XFF], ids=["delta", "xff"])
@pytest.mark.parametrize("esz,D", SHAPES)
def test_corrupt_frames_fail_cleanly(spz, pred, esz, D, force):
    import torch
    spz.set_kernel_path(force)
    rng = np.random.default_rng(D * 7 + esz)
    x = inputs.ar1(64 * 40, D, esz, seed=D)
    xt = torch.from_numpy(x.view(np.uint8 if esz == 1 else np.int16)).cuda()
    fr = spz.frame_encode(pred, xt, D, 64)
    nbytes = fr.nbytes()
    good = fr.buf[:nbytes].clone()
    off, nb, po = fr.directory()
    nch = fr.nchunks

    def image():
        return good.clone()

    # 1. directory entries pointing past the frame / far past it / backwards
    for victim, value in ((3, nbytes * 2), (5, 1 << 62), (nch - 1, (1 << 64) - 16), (7, 16)):
        b = image()
        b[64 + 8 * victim: 64 + 8 * victim + 8] = torch.from_numpy(np.array([value], np.uint64).view(np.uint8)).cuda()
        rc, st, out, intact = _decode_frame_raw(spz, fr, b)
        assert intact
        assert rc < 0 or st[victim] != 0 or value == 16, (victim, value, rc, st[victim])
        good_chunks = [c for c in range(nch) if c != victim]
        o = out.cpu().numpy().view(x.dtype)
        for c in good_chunks[:6]:
            assert np.array_equal(o[c * 64 * D: (c + 1) * 64 * D], x[c * 64 * D: (c + 1) * 64 * D])
    # 2. truncated frame: readable bytes end in the middle of the payload
    for cut in (po + int(off[nch // 2]) + 5, po + 16, po + int(off[nch - 1]) + 3):
        b = image()[:cut].clone()
        rc, st, out, intact = _decode_frame_raw(spz, fr, b)
        assert intact
        assert rc < 0 or (st != 0).any()
    # 3. huge / absurd group counts and tail lengths in chunk metadata
    for c, field in ((2, 0xFFFFFFFF), (4, 0x7FFFFFFF), (6, 123456)):
        b = image()
        a = po + int(off[c])
        b[a: a + 4] = torch.from_numpy(np.array([field], np.uint32).view(np.uint8)).cuda()
        rc, st, out, intact = _decode_frame_raw(spz, fr, b)
        assert intact and (rc < 0 or st[c] != 0), (c, field, rc, st[c])
    # 4. wrong ndims in one chunk's metadata
    b = image()
    a = po + int(off[1])
    b[a + 6: a + 8] = torch.from_numpy(np.array([D + 1], np.uint16).view(np.uint8)).cuda()
    rc, st, out, intact = _decode_frame_raw(spz, fr, b)
    assert intact and st[1] != 0
    # 5. random bit flips in the payload: any outcome but a crash or an overrun
    for trial in range(8):
        b = image()
        pos = rng.integers(po, nbytes - 16, size=24)
        flip = torch.from_numpy(rng.integers(1, 256, size=24).astype(np.uint8)).cuda()
        b[torch.from_numpy(pos).cuda()] ^= flip
        rc, st, out, intact = _decode_frame_raw(spz, fr, b)
        assert intact
    # 6. destination too small
    rc, st, out, intact = _decode_frame_raw(spz, fr, image(), dest_elems=fr.len - 64 * D - 3)
    assert rc == -3 and intact
    _alive(spz)


@pytest.mark.parametrize("pred", [DELTA, XFF], ids=["delta", "xff"])
@pytest.mark.parametrize("esz,D", SHAPES)
def test_truncated_single_streams(spz, oracle, pred, esz, D):
    """spz_decompress_dev with src_bytes smaller than the stream: status 3 (truncated), nothing past dest_cap."""
    import torch
    x = inputs.ar1(200, D, esz, seed=3)
    s, _ = oracle.compress(pred, x, D)
    L = spz.lib()
    for keep in (len(s) // 2, 9, len(s) - 1):
        src = torch.zeros(len(s) + 64, dtype=torch.uint8, device="cuda")
        src[: len(s)] = torch.from_numpy(s).cuda()
        out = torch.full((x.size + 1024,), 0x55, dtype=torch.uint8 if esz == 1 else torch.int16, device="cuda")
        meta = torch.zeros(4, dtype=torch.int32, device="cuda")
        rc = L.spz_decompress_dev(pred, esz, src.data_ptr(), keep, 1, D, out.data_ptr(), x.size, meta.data_ptr(), None)
        torch.cuda.synchronize()
        assert rc == 0
        assert int(meta[3].item()) != 0, (keep, len(s))
        assert bool((out[x.size:] == 0x55).all().item())
    _alive(spz)


def test_host_frame_rejects_bad_directories(spz):
    x = inputs.ar1(64 * 30, 8, 1, seed=2)
    fr = spz.frame_compress_host(XFF, x, 8, 64)
    assert np.array_equal(spz.frame_decompress_host(fr), x)
    L = spz.lib()
    out = np.zeros(x.size + 64, np.uint8)

    def rc_of(img):
        return L.spz_frame_decompress_host(img.ctypes.data, img.size, out.ctypes.data, x.size)

    for victim, value in ((3, 1 << 40), (4, 8), (5, 17)):
        img = fr.copy()
        img[64 + 8 * victim: 64 + 8 * victim + 8] = np.array([value], np.uint64).view(np.uint8)
        assert rc_of(img) == -4
    img = fr.copy(); img[0] ^= 1
    assert rc_of(img) == -4                       # magic
    img = fr.copy(); img[16:24] = np.array([1 << 50], np.uint64).view(np.uint8)
    assert rc_of(img) in (-3, -4)                 # absurd length
    assert rc_of(fr[: len(fr) // 2].copy()) == -4  # truncated image
    _alive(spz)
