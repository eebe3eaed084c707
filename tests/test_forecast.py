"""Stand-alone forecaster transforms: encode/decode_{delta,doubledelta}_rowmajor_{8b,16b}
(delta.h:17-67; reference tests: test/test_delta.cpp:27-74).

CPU: the oracle's restatement against the compiled reference (bytes incl. metadata, return values, round trips).
GPU: the CUDA kernels (packed element-wise encode, chained-scan decode) against the oracle through the C ABI."""
import numpy as np
import pytest

import inputs

ORDERS = [(1, 1), (2, 1), (1, 2), (2, 2)]
OIDS = ["delta8", "ddelta8", "delta16", "ddelta16"]
DIMS = (1, 2, 3, 4, 5, 7, 8, 9, 16, 17, 31, 32, 33, 64, 65, 80)


def _cases(esz, big=True):
    for n in inputs.REF_SIZES + ([4096 + 3, 100000] if big else []):
        fams = inputs.reference_families(n, esz)
        for tag in ("i%64", "fuzz256", "squares", "zeros", "sparse6"):
            if tag in fams:
                yield f"{tag}/n{n}", fams[tag]


@pytest.mark.parametrize("order,esz", ORDERS, ids=OIDS)
def test_oracle_forecast_vs_reference(oracle, ref, order, esz):
    bad = []
    for D in DIMS:
        for tag, x in _cases(esz):
            ws, wr = oracle.encode_forecast(order, x, D)
            rs, rr = ref.encode_forecast(order, x, D)
            msg = None
            if wr != rr:
                msg = f"return {wr} != {rr}"
            elif not np.array_equal(ws, rs):
                msg = "transformed bytes differ"
            else:
                y = ws[6:].view(x.dtype)
                d, dn = oracle.decode_forecast(order, esz, y, D)
                rd, rn = ref.decode_forecast(order, esz, y, D)
                if dn != rn or not np.array_equal(d, x) or not np.array_equal(rd, x):
                    msg = "decode differs"
            if msg:
                bad.append((D, tag, msg))
    assert not bad, (len(bad), bad[:5])


@pytest.mark.parametrize("esz", [1, 2], ids=["xff8", "xff16"])
def test_oracle_xff_forecast_vs_reference(oracle, ref, esz):
    """encode/decode_xff_rowmajor_{8b,16b} (predict.h:15-30): the restatement of the prototype FIRE transform -- block
    count rule, delta-coded remainder, unsigned even-dim multiply (8-bit), wrapping gradient sums -- against the
    compiled reference: bytes incl. metadata, return values, and both decoders invert both encoders' output.
    Groundwork: there is no product kernel for this transform yet."""
    bad = []
    for D in DIMS:
        for tag, x in _cases(esz):
            ws, wr = oracle.encode_xff_forecast(x, D)
            rs, rr = ref.encode_xff_forecast(x, D)
            msg = None
            if wr != rr:
                msg = f"return {wr} != {rr}"
            elif not np.array_equal(ws, rs):
                msg = "transformed bytes differ"
            else:
                y = ws[6:].view(x.dtype)
                d, dn = oracle.decode_xff_forecast(esz, y, D)
                rd, rn = ref.decode_xff_forecast(esz, y, D)
                if dn != rn or not np.array_equal(d, x) or not np.array_equal(rd, x):
                    msg = "decode differs"
            if msg:
                bad.append((D, tag, msg))
    assert not bad, (len(bad), bad[:5])


def _golden_forecast():
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_forecast_v1.npz"))
    for k, line in enumerate(z["index"]):
        tr, esz, D, ret, tag = str(line).split("|", 4)
        yield int(tr), int(esz), int(D), int(ret), tag, z[f"x{k}"], z[f"y{k}"]


def test_oracle_matches_golden_forecast_vectors(oracle):
    """committed fixtures generated from the compiled reference (tests/golden/make_golden_forecast.py): they travel to
    hosts without the reference tree"""
    n = 0
    for tr, esz, D, ret, tag, x, y in _golden_forecast():
        got, r = oracle.encode_xff_forecast(x, D) if tr == 3 else oracle.encode_forecast(tr, x, D)
        assert r == ret and np.array_equal(got, y), (tr, esz, D, tag)
        dec, _ = (oracle.decode_xff_forecast(esz, y[6:].view(x.dtype), D) if tr == 3
                  else oracle.decode_forecast(tr, esz, y[6:].view(x.dtype), D))
        assert np.array_equal(dec, x), (tr, esz, D, tag)
        n += 1
    assert n > 1000


@pytest.mark.gpu
def test_gpu_forecast_matches_golden_vectors():
    """the delta / double-delta kernels against the reference's own output bytes"""
    import sprintz_b200 as spz
    spz.lib()
    for tr, esz, D, ret, tag, x, y in _golden_forecast():
        if tr == 3:
            continue                                   # the FIRE transform: tests/test_zz_xff_forecast_gpu.py
        got, r = spz.encode_forecast(tr, x, D)
        assert r == ret and np.array_equal(got, y), (tr, esz, D, tag)
        dec, _ = spz.decode_forecast(tr, esz, y[6:].view(x.dtype), D)
        assert np.array_equal(dec, x), (tr, esz, D, tag)


@pytest.mark.gpu
@pytest.mark.parametrize("order,esz", ORDERS, ids=OIDS)
def test_gpu_forecast_matches_oracle(oracle, order, esz):
    import sprintz_b200 as spz
    spz.lib()
    bad = []
    for D in DIMS + (200, 1000):
        for tag, x in _cases(esz):
            ws, wr = oracle.encode_forecast(order, x, D)
            gs, gr = spz.encode_forecast(order, x, D)
            msg = None
            if gr != wr or not np.array_equal(gs, ws):
                msg = "encode differs"
            else:
                nometa, nr = spz.encode_forecast(order, x, D, write_size=False)
                if nr != x.size or not np.array_equal(nometa, ws[6:]):
                    msg = "write_size=false differs"
            if msg is None:
                d, dn = spz.decode_forecast(order, esz, ws[6:].view(x.dtype), D)
                if dn != x.size or not np.array_equal(d, x):
                    msg = "decode differs"
            if msg:
                bad.append((D, tag, msg))
    assert not bad, (len(bad), bad[:5])


@pytest.mark.gpu
@pytest.mark.parametrize("order,esz,D", [(1, 1, 8), (2, 1, 8), (1, 2, 16), (2, 2, 16), (1, 1, 80), (2, 2, 3)])
def test_gpu_forecast_large_device_roundtrip(oracle, order, esz, D):
    """64 MiB on the device API: many scan tiles per dim; checked against the oracle on a prefix and the tail,
    and decode(encode(x)) == x over the whole array."""
    import torch
    import sprintz_b200 as spz
    from sprintz_b200.synth import ar1_stream
    rows = (64 << 20) // (esz * D) + 5
    x = ar1_stream(rows, D, esz, seed=3, device="cuda")[: rows * D - 2]       # ragged last row
    y = spz.forecast_encode_dev(order, x, D)
    z = spz.forecast_decode_dev(order, y, D)
    torch.cuda.synchronize()
    assert torch.equal(z, x)
    udt = np.uint8 if esz == 1 else np.uint16
    xh = x.cpu().numpy().view(udt)
    yh = y.cpu().numpy().view(udt)
    want, _ = oracle.encode_forecast(order, xh[: 1 << 20], D, write_size=False)
    assert np.array_equal(yh[: 1 << 20].view(np.uint8), want)
    # the oracle's transform is local (lags D and 2D): check the tail against a window with enough history
    k = (1 << 18) // D * D
    w2, _ = oracle.encode_forecast(order, xh[xh.size - k - 2 * D - (xh.size - k - 2 * D) % D * 0:], D, write_size=False)
    assert np.array_equal(yh[-(1 << 16):].view(np.uint8), w2[-(1 << 16) * esz:])


@pytest.mark.gpu
def test_gpu_forecast_named_entry_points_vs_live_reference(ref):
    import sprintz_b200 as spz
    L = spz.lib()
    for name, order, esz, D in (("delta_rowmajor_8b", 1, 1, 8), ("delta_rowmajor_16b", 1, 2, 5),
                                ("doubledelta_rowmajor_8b", 2, 1, 33), ("doubledelta_rowmajor_16b", 2, 2, 16)):
        x = inputs.random_walk_with_runs(2000, D, esz, seed=4)[:-3]
        rs, rr = ref.encode_forecast(order, x, D)
        out = np.zeros(x.size * esz + 16, np.uint8)
        ret = getattr(L, "spz_encode_" + name)(x.ctypes.data, x.size, out.ctypes.data, D, 1)
        assert ret == rr and np.array_equal(out[: rr * esz], rs)
        dec = np.zeros(x.size, x.dtype)
        n = getattr(L, "spz_decode_" + name)(out.ctypes.data + 6, x.size, dec.ctypes.data, D)
        assert n == x.size and np.array_equal(dec, x)


@pytest.mark.gpu
@pytest.mark.parametrize("esz,D", [(1, 8), (2, 4), (1, 16), (2, 8), (2, 16), (1, 80), (1, 48), (2, 40)])
def test_gpu_delta_decode_vector_path(oracle, esz, D):
    """Rows of 8 / 16 / k*16 bytes with a whole number of 16-byte units take the vectorised chained scan
    (spz_delta_decode_vec_kernel): one tile, tile edges, many tiles per column, many CTAs -- against the oracle on the small
    sizes and against an exact cumulative sum everywhere."""
    import torch
    import sprintz_b200 as spz
    udt = np.uint8 if esz == 1 else np.uint16
    row_bytes = D * esz
    unit_rows = 2 if row_bytes == 8 else 1
    rng = np.random.default_rng(D + esz)
    for nrows in (unit_rows, 2 * unit_rows, 62, 64, 66, 128, 130, 64 * 7 + 2, 4096, 300000):
        nrows -= nrows % unit_rows
        y = rng.integers(0, 1 << (8 * esz), size=(nrows, D), dtype=np.uint32).astype(udt)
        want = np.cumsum(y.astype(np.uint64), axis=0).astype(udt)            # mod 2^w
        yt = torch.from_numpy(y.reshape(-1).view(np.uint8 if esz == 1 else np.int16)).cuda()
        got = spz.forecast_decode_dev(1, yt, D).cpu().numpy().view(udt).reshape(nrows, D)
        assert np.array_equal(got, want), (esz, D, nrows)
        if nrows <= 4096:
            od, n = oracle.decode_forecast(1, esz, y.reshape(-1), D)
            assert n == y.size and np.array_equal(od.reshape(nrows, D), want)
