"""GPU parity of the entropy stage (spz_huff_encode_dev / spz_huff_decode_dev) against its CPU specification
(oracle/huff_oracle.c), through the C ABI: container bytes identical, decode == input, malformed containers flagged.
PARITY UNPINNED beyond that specification: Huff0 is not in the reference tree."""
import numpy as np
import pytest

import inputs
from oracle import XFF
from test_huff_oracle import _families

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def spz():
    import sprintz_b200
    sprintz_b200.lib()
    return sprintz_b200


def _check(spz, oracle, x, name):
    import torch
    xt = torch.from_numpy(x.copy()).cuda() if x.size else torch.zeros(0, dtype=torch.uint8, device="cuda")
    got = spz.huff_encode_dev(xt).cpu().numpy()
    want = oracle.huff_encode(x)
    assert len(got) == len(want) and np.array_equal(got, want), (name, len(got), len(want),
                                                                 int(np.nonzero(got[:min(len(got), len(want))] != want[:min(len(got), len(want))])[0][:1].sum()))
    back = spz.huff_decode_dev(torch.from_numpy(want.copy()).cuda(), x.size).cpu().numpy()
    assert np.array_equal(back, x), name


def test_containers_match_the_specification(spz, oracle):
    for name, x in _families().items():
        _check(spz, oracle, x, name)


def test_codec_payloads(spz, oracle):
    """what the stage is for: the bit-packed bytes of a frame (headers + payload)"""
    import torch
    for esz, D in ((1, 8), (2, 16), (1, 80)):
        x = inputs.ar1(40000, D, esz, seed=D)
        xt = torch.from_numpy(x.view(np.uint8 if esz == 1 else np.int16)).cuda()
        fr = spz.frame_encode(XFF if D != 80 else 0, xt, D, 512)
        torch.cuda.synchronize()
        h = fr.header()
        pay = fr.buf[int(h.payload_off): int(h.payload_off) + int(h.payload_bytes)].clone()
        _check(spz, oracle, pay.cpu().numpy(), f"payload {esz}x{D}")


def test_large_buffer_round_trip(spz):
    import torch
    g = torch.Generator(device="cuda").manual_seed(5)
    n = (192 << 20) + 12345
    x = (torch.randn(n, device="cuda", generator=g) * 9).to(torch.int32).clamp(-128, 127).to(torch.int8).view(torch.uint8)
    h = spz.HuffCodec(n)
    cont = h.encode(x)
    size = h.size()
    assert size < 0.8 * n                                   # a peaked byte distribution must shrink
    out = torch.empty_like(x)
    h.decode(cont, out)
    h.check_status()
    assert torch.equal(out, x)


def test_malformed_containers_are_flagged(spz, oracle):
    import torch
    x = _families()["geometric"]
    c = oracle.huff_encode(x)
    h = spz.HuffCodec(x.size)
    out = torch.empty(x.size, dtype=torch.uint8, device="cuda")

    def status(buf, nbytes=None):
        t = torch.from_numpy(buf.copy()).cuda()
        h.decode(t, out, cont_bytes=nbytes)
        torch.cuda.synchronize()
        return int(h.status.item())

    assert status(c) == 0
    bad = c.copy(); bad[0] ^= 1
    assert status(bad) & 2
    bad = c.copy(); bad[64:192] = 0x11
    assert status(bad) & 1
    bad = c.copy(); bad[192 + 8] = 0xff; bad[192 + 14] = 0x7f
    assert status(bad) & 4
    bad = c.copy()
    po = int.from_bytes(bad[32:40].tobytes(), "little")
    bad[po + 20: po + 60] ^= 0x5a                          # garbage inside a segment's stream
    assert status(bad) != 0 or not np.array_equal(out.cpu().numpy(), x)
