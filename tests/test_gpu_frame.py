"""GPU tests of the chunked frame: every chunk byte-identical to the reference's stream for that slice,
decode == input, directory consistent; plus size-independent properties at BASELINE.json's full sizes."""
import numpy as np
import pytest

import inputs
import parity
from oracle import DELTA, XFF

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def spz():
    import sprintz_b200
    sprintz_b200.lib()
    return sprintz_b200


def _frame_vs_oracle(spz, oracle, pred, esz, D, x, chunk_rows):
    import torch
    xt = torch.from_numpy(x.view(np.uint8 if esz == 1 else np.int16)).cuda()
    fr = spz.frame_encode(pred, xt, D, chunk_rows)
    torch.cuda.synchronize()
    h = fr.header()
    off, nb, po = fr.directory()
    buf = fr.buf[: fr.nbytes()].cpu().numpy()
    assert h.magic == 0x315A5053 and h.len == x.size and h.ndims == D and h.chunk_rows == chunk_rows
    clen = chunk_rows * D
    nch = (x.size + clen - 1) // clen if x.size else 1
    assert h.nchunks == nch and len(nb) == nch
    # directory: 16-byte aligned, contiguous, exclusive scan of rounded sizes
    assert off[0] == 0 and np.all(off % 16 == 0)
    assert np.array_equal(np.diff(off), (nb.astype(np.int64) + 15) // 16 * 16)
    assert h.payload_bytes == off[-1] and h.frame_bytes == po + off[-1] + 16
    for c in range(nch):
        ws, wr = oracle.compress(pred, x[c * clen: (c + 1) * clen], D)
        gs = buf[po + off[c]: po + off[c] + nb[c]]
        assert len(gs) == len(ws), f"chunk {c}: {len(gs)} bytes, oracle {len(ws)}"
        assert np.array_equal(gs, ws), f"chunk {c}: " + str(parity.compare_stream(gs, wr, ws, wr, esz))
        pad = buf[po + off[c] + nb[c]: po + off[c + 1]]
        assert not pad.any(), "alignment padding must be zero"
    out = spz.frame_decode(fr)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy().view(x.dtype), x)


@pytest.mark.parametrize("pred", [DELTA, XFF], ids=["delta", "xff"])
@pytest.mark.parametrize("esz,D", [(1, 8), (2, 16), (1, 80), (1, 1), (1, 4), (1, 5), (2, 2), (2, 3), (2, 33), (1, 200)])
def test_frame_chunks_match_oracle(spz, oracle, pred, esz, D):
    for chunk_rows, nrows, ragged in ((16, 400, 0), (256, 5000, 3), (4096, 20000, 0), (256, 256, 0), (256, 100, 0)):
        x = inputs.ar1(nrows, D, esz, seed=nrows)
        if ragged:
            x = x[:-ragged]
        _frame_vs_oracle(spz, oracle, pred, esz, D, x, chunk_rows)
    x = inputs.random_walk_with_runs(3000, D, esz, seed=5)       # run slots inside chunks
    _frame_vs_oracle(spz, oracle, pred, esz, D, x, 512)
    _frame_vs_oracle(spz, oracle, pred, esz, D, np.zeros(4096 * D, x.dtype), 1024)


def test_frame_host_api_roundtrip(spz, oracle):
    for pred, esz, D in ((XFF, 1, 8), (XFF, 2, 16), (DELTA, 1, 80)):
        x = inputs.ar1(10000, D, esz)
        fr = spz.frame_compress_host(pred, x, D, 256)
        y = spz.frame_decompress_host(fr)
        assert np.array_equal(y, x)
        # the host frame is the same bytes the device API produces
        import torch
        xt = torch.from_numpy(x.view(np.uint8 if esz == 1 else np.int16)).cuda()
        f2 = spz.frame_encode(pred, xt, D, 256)
        assert np.array_equal(f2.buf[: f2.nbytes()].cpu().numpy()[:len(fr) - 16], fr[:-16])


@pytest.mark.parametrize("cfg", ["c2", "c3", "c4"])
def test_full_size_properties(spz, ref, cfg):
    """BASELINE.json's full sizes (1 GiB int8/8-dim FIRE, 4 GiB int16/16-dim FIRE, 4 GiB int8/80-dim delta):
    encode -> decode round trip is the identity, the directory is an exact scan, and a sample of chunks spread
    over the stream is byte-identical to the compiled reference."""
    import torch
    import bench
    from sprintz_b200.synth import ar1_stream
    c = bench.CONFIGS[cfg]
    esz, D, pred, chunk_rows = c["esz"], c["D"], c["pred"], c["chunk_rows"]
    rows = c["nbytes"] // (esz * D); rows -= rows % 16
    x = ar1_stream(rows, D, esz, seed=4242, device="cuda")
    codec = spz.FrameCodec(pred, esz, x.numel(), D, chunk_rows)
    fr = codec.encode(x)
    out = torch.empty_like(x)
    codec.decode(fr, out)
    torch.cuda.synchronize()
    codec.check_status()
    assert torch.equal(out, x)
    del out
    off, nb, po = fr.directory()
    assert np.array_equal(np.diff(off), (nb.astype(np.int64) + 15) // 16 * 16)
    nch = codec.nchunks
    clen = chunk_rows * D
    rng = np.random.default_rng(1)
    sample = sorted(set([0, 1, nch - 1, nch // 2] + list(rng.integers(0, nch, size=60))))
    udt = np.uint8 if esz == 1 else np.uint16
    for ci in sample:
        xs = x[ci * clen: (ci + 1) * clen].cpu().numpy().view(udt)
        rs, rr = ref.compress(pred, xs, D)
        gs = fr.buf[po + int(off[ci]): po + int(off[ci]) + int(nb[ci])].cpu().numpy()
        msg = parity.compare_stream(gs, len(gs) // esz, rs, rr, esz)
        assert msg is None, f"{cfg} chunk {ci}: {msg}"
    # idempotence: encoding the same input again yields the same frame bytes
    payload_bytes = int(fr.header().payload_bytes)
    sums1 = fr.buf[po: po + payload_bytes].view(torch.int64)[: payload_bytes // 8].sum().item()
    fr2 = codec.encode(x)
    torch.cuda.synchronize()
    sums2 = fr2.buf[po: po + payload_bytes].view(torch.int64)[: payload_bytes // 8].sum().item()
    assert sums1 == sums2


SWEEP_DIMS = (1, 2, 3, 4, 5, 6, 7, 8, 9, 12, 16, 17, 24, 31, 32, 33, 40, 48, 64, 65, 80, 96, 128, 129, 200, 256)


@pytest.mark.parametrize("pred", [DELTA, XFF], ids=["delta", "xff"])
@pytest.mark.parametrize("esz", [1, 2])
def test_frame_many_chunks_every_team_shape(spz, oracle, pred, esz):
    """BASELINE config 5's dims axis (1..256) at the reference's block size: frames with hundreds of chunks, so every
    team shape fills whole CTAs (the last team of a CTA sits at the end of its shared-memory window).  Round trip
    over the whole frame + a sample of chunks byte-for-byte against the oracle."""
    import torch
    rng = np.random.default_rng(7)
    for D in SWEEP_DIMS:
        chunk_rows = 64
        nrows = chunk_rows * 300 + 16
        x = inputs.ar1(nrows, D, esz, seed=D)
        xt = torch.from_numpy(x.view(np.uint8 if esz == 1 else np.int16)).cuda()
        fr = spz.frame_encode(pred, xt, D, chunk_rows)
        out = spz.frame_decode(fr)
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy().view(x.dtype), x), f"round trip, ndims {D}"
        off, nb, po = fr.directory()
        buf = fr.buf[: fr.nbytes()].cpu().numpy()
        clen = chunk_rows * D
        for c in sorted(set([0, 1, 31, 32, 127, 128, 299, 300] + list(rng.integers(0, 301, size=6)))):
            ws, wr = oracle.compress(pred, x[c * clen: (c + 1) * clen], D)
            gs = buf[po + off[c]: po + off[c] + nb[c]]
            assert len(gs) == len(ws) and np.array_equal(gs, ws), f"ndims {D} chunk {c}"
