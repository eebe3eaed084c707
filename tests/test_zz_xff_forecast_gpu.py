"""GPU side of the stand-alone FIRE transform (encode/decode_xff_rowmajor_{8b,16b}, predict.h:15-30): the kernel runs the
per-dim code that test_xff_forecast_host.py pins on the CPU, one thread per dim; here the plumbing (indexing, C ABI,
metadata, returns) is checked against the oracle and the live reference.  (Written after this round's GPU budget was
spent: first executed by the round-end GPU run -- which is why the file sorts last.)"""
import numpy as np
import pytest

import inputs

pytestmark = pytest.mark.gpu

DIMS = (1, 2, 3, 4, 5, 7, 8, 9, 16, 17, 31, 32, 33, 64, 65, 80, 200, 1000)


def _cases(esz):
    for n in inputs.REF_SIZES + [4096 + 3, 50000]:
        fams = inputs.reference_families(n, esz)
        for tag in ("i%64", "fuzz256", "squares", "zeros", "sparse6"):
            yield f"{tag}/n{n}", fams[tag]


@pytest.mark.parametrize("esz", [1, 2], ids=["xff8", "xff16"])
def test_gpu_xff_forecast_matches_oracle(oracle, esz):
    import sprintz_b200 as spz
    spz.lib()
    bad = []
    for D in DIMS:
        for tag, x in _cases(esz):
            ws, wr = oracle.encode_xff_forecast(x, D)
            gs, gr = spz.encode_xff_forecast(x, D)
            msg = None
            if gr != wr or not np.array_equal(gs, ws):
                msg = "encode differs"
            else:
                nometa, nr = spz.encode_xff_forecast(x, D, write_size=False)
                if nr != x.size or not np.array_equal(nometa, ws[6:]):
                    msg = "write_size=false differs"
            if msg is None:
                d, dn = spz.decode_xff_forecast(esz, ws[6:].view(x.dtype), D)
                if dn != x.size or not np.array_equal(d, x):
                    msg = "decode differs"
            if msg:
                bad.append((D, tag, msg))
    assert not bad, (len(bad), bad[:5])


def test_gpu_xff_forecast_named_entry_points_vs_live_reference(ref):
    import sprintz_b200 as spz
    L = spz.lib()
    for name, esz, D in (("xff_rowmajor_8b", 1, 8), ("xff_rowmajor_8b", 1, 33), ("xff_rowmajor_16b", 2, 5), ("xff_rowmajor_16b", 2, 16)):
        x = inputs.random_walk_with_runs(2000, D, esz, seed=4)[:-3]
        rs, rr = ref.encode_xff_forecast(x, D)
        out = np.zeros(x.size * esz + 16, np.uint8)
        ret = getattr(L, "spz_encode_" + name)(x.ctypes.data, x.size, out.ctypes.data, D, 1)
        assert ret == rr and np.array_equal(out[: rr * esz], rs)
        dec = np.zeros(x.size, x.dtype)
        n = getattr(L, "spz_decode_" + name)(out.ctypes.data + 6, x.size, dec.ctypes.data, D)
        assert n == x.size and np.array_equal(dec, x)
        rdec, rn = ref.decode_xff_forecast(esz, out[6: 6 + x.size * esz].view(x.dtype), D)
        assert rn == x.size and np.array_equal(rdec, x)


def test_gpu_xff_forecast_device_api_roundtrip():
    import torch
    import sprintz_b200 as spz
    from sprintz_b200.synth import ar1_stream
    L = spz.lib()
    for esz, D in ((1, 8), (2, 16), (1, 80)):
        x = ar1_stream(4096 + 3, D, esz, seed=5, device="cuda")[: (4096 + 3) * D - 1]
        y = torch.empty_like(x)
        z = torch.empty_like(x)
        sp = spz.api._stream_ptr()
        assert L.spz_xff_forecast_dev(1, esz, x.data_ptr(), x.numel(), D, y.data_ptr(), sp) == 0
        assert L.spz_xff_forecast_dev(0, esz, y.data_ptr(), x.numel(), D, z.data_ptr(), sp) == 0
        torch.cuda.synchronize()
        assert torch.equal(z, x) and not torch.equal(y, x)


def test_gpu_xff_forecast_matches_golden_vectors():
    """against the reference's own output bytes (tests/golden/golden_forecast_v1.npz, transform 3)"""
    import sprintz_b200 as spz
    from test_forecast import _golden_forecast
    spz.lib()
    n = 0
    for tr, esz, D, ret, tag, x, y in _golden_forecast():
        if tr != 3:
            continue
        got, r = spz.encode_xff_forecast(x, D)
        assert r == ret and np.array_equal(got, y), (esz, D, tag)
        dec, _ = spz.decode_xff_forecast(esz, y[6:].view(x.dtype), D)
        assert np.array_equal(dec, x), (esz, D, tag)
        n += 1
    assert n > 300
