"""Query-on-compressed: query_rowmajor_{delta,xff}_rle_{8b,16b} with QueryParams{op, materialize}
(query.hpp:22-29, sprintz_delta.h:92-98, sprintz_xff.h:86-93; reference tests: test/test_query.cpp).

The reference's query functors never return their state (query.hpp:231), so -- as SURVEY 8f says -- the oracle is the
reduction of the decoded oracle output (oracle_query); the decoded elements stay pinned to the reference as before.
CPU: oracle_query against numpy on the reference's test matrix.  GPU: host streams, frames on the fused lane path
(compressed bytes only), frames on the decode + reduce path, and a full-size property check."""
import numpy as np
import pytest

import inputs
from oracle import DELTA, XFF

CODECS = [(DELTA, 1), (XFF, 1), (DELTA, 2), (XFF, 2)]
IDS = ["delta8", "xff8", "delta16", "xff16"]
Q_MAX, Q_SUM = 1, 2


def _np_reduce(x, D, op):
    r = np.zeros(D, np.uint64)
    for d in range(D):
        col = x[d::D].astype(np.uint64)
        if col.size:
            r[d] = col.max() if op == Q_MAX else col.sum()
    return r


@pytest.mark.parametrize("pred,esz", CODECS, ids=IDS)
def test_oracle_query_is_reduction_of_decoded_output(oracle, pred, esz):
    for D in (1, 3, 8, 16, 33, 80):
        for n in (1, 127, 128, 137, 4096 + 17):
            for tag, x in inputs.reference_families(n, esz).items():
                if tag not in ("i%64", "fuzz256", "zeros", "sparse6"):
                    continue
                s, _ = oracle.compress(pred, x, D)
                for op in (Q_MAX, Q_SUM):
                    res, dec, ret = oracle.query(pred, esz, s, x.size, op)
                    assert ret == x.size and np.array_equal(dec, x)
                    assert np.array_equal(res, _np_reduce(x, D, op)), (D, n, tag, op)


@pytest.mark.gpu
@pytest.mark.parametrize("pred,esz", CODECS, ids=IDS)
def test_gpu_query_host_streams_match_oracle(oracle, pred, esz):
    import sprintz_b200 as spz
    spz.lib()
    bad = []
    for D in (1, 2, 3, 5, 8, 16, 17, 80):
        cases = [(f"{tag}/n{n}", x) for n in (1, 127, 128, 137, 4096 + 17)
                 for tag, x in inputs.reference_families(n, esz).items() if tag in ("i%64", "fuzz256", "zeros", "sparse6")]
        cases.append(("walk", inputs.random_walk_with_runs(700, D, esz, seed=D)[:-1]))
        for tag, x in cases:
            s, _ = oracle.compress(pred, x, D)
            for op in (Q_MAX, Q_SUM):
                want, dec, wret = oracle.query(pred, esz, s, x.size, op)
                got, out, ret = spz.query(pred, esz, s, op, materialize=True)
                got2, out2, ret2 = spz.query(pred, esz, s, op, materialize=False)
                if ret != wret or ret2 != wret or not np.array_equal(out, dec) or out2 is not None:
                    bad.append((D, tag, op, "decoded output"))
                elif not np.array_equal(got[:D], want[:D]) or not np.array_equal(got2[:D], want[:D]):
                    bad.append((D, tag, op, "result"))
    assert not bad, (len(bad), bad[:5])


@pytest.mark.gpu
def test_gpu_query_named_entry_points(oracle):
    import sprintz_b200 as spz
    L = spz.lib()
    for name, pred, esz, D in (("delta_rle_8b", DELTA, 1, 8), ("delta_rle_16b", DELTA, 2, 5),
                               ("xff_rle_8b", XFF, 1, 33), ("xff_rle_16b", XFF, 2, 16)):
        x = inputs.random_walk_with_runs(1000, D, esz, seed=2)
        s, _ = oracle.compress(pred, x, D)
        src = np.concatenate([s, np.zeros(16, np.uint8)])
        for op in (Q_MAX, Q_SUM):
            out = np.zeros(x.size, x.dtype)
            res = np.zeros(D, np.uint64)
            n = getattr(L, "spz_query_rowmajor_" + name)(src.ctypes.data, out.ctypes.data, op, 1, res.ctypes.data)
            assert n == x.size and np.array_equal(out, x) and np.array_equal(res, _np_reduce(x, D, op))
        # NOOP + materialize is exactly decompress (test_query.cpp:61-75)
        out = np.zeros(x.size, x.dtype)
        n = getattr(L, "spz_query_rowmajor_" + name)(src.ctypes.data, out.ctypes.data, 0, 1, None)
        assert n == x.size and np.array_equal(out, x)


def _frame_query_case(spz, pred, esz, D, x, chunk_rows, expect_fused):
    import torch
    xt = torch.from_numpy(x.view(np.uint8 if esz == 1 else np.int16)).cuda()
    fr = spz.frame_encode(pred, xt, D, chunk_rows)
    for op in (Q_MAX, Q_SUM):
        want = _np_reduce(x, D, op)
        res, none = spz.frame_query(fr, op, materialize=False)
        torch.cuda.synchronize()
        assert none is None
        assert spz.last_query_fused() == expect_fused
        assert np.array_equal(res.cpu().numpy().view(np.uint64), want), (pred, esz, D, chunk_rows, op, "no-materialize")
        res, out = spz.frame_query(fr, op, materialize=True)
        torch.cuda.synchronize()
        assert not spz.last_query_fused()
        assert np.array_equal(out.cpu().numpy().view(x.dtype), x)
        assert np.array_equal(res.cpu().numpy().view(np.uint64), want), (pred, esz, D, chunk_rows, op, "materialize")


@pytest.mark.gpu
@pytest.mark.parametrize("pred", [DELTA, XFF, 2, 3], ids=["delta", "xff", "rm_delta", "rm_plain"])
@pytest.mark.parametrize("esz,D", [(1, 8), (1, 16), (2, 8), (2, 16)])
def test_gpu_frame_query_fused_lane_path(pred, esz, D):
    """materialize=False on 8/16-dim frames: decoder lanes fold rows in registers (forced here on small frames)."""
    import sprintz_b200 as spz
    spz.lib()
    old = spz.set_kernel_path(2)
    try:
        udt = np.uint8 if esz == 1 else np.uint16
        for chunk_rows, nrows, ragged in ((16, 400, 0), (256, 5000, 3), (4096, 20000, 0), (256, 100, 0), (64, 64 * 300, 0)):
            x = inputs.ar1(nrows, D, esz, seed=nrows)
            if ragged:
                x = x[:-ragged]
            _frame_query_case(spz, pred, esz, D, x, chunk_rows, True)
        _frame_query_case(spz, pred, esz, D, inputs.random_walk_with_runs(3000, D, esz, seed=5), 512, True)
        _frame_query_case(spz, pred, esz, D, np.zeros(4096 * D, udt), 1024, True)
        # long runs of a large constant (delta run folding, 64-bit totals), then noise
        x = np.full(70000 * D, (1 << (8 * esz)) - 2, udt)
        x[-500 * D:] = inputs.fuzz(500 * D, esz)["fuzz256"]
        _frame_query_case(spz, pred, esz, D, x, 35008, True)
        # a chunk longer than 4096 blocks of near-maximal values: the 32-bit partial sums are drained
        x = ((1 << (8 * esz)) - 1 - inputs.fuzz(40000 * D, esz)["fuzz16"]).astype(udt)
        _frame_query_case(spz, pred, esz, D, x, 40000, True)
        # a linear ramp: FIRE learns it and emits run slots with a non-zero previous delta
        _frame_query_case(spz, pred, esz, D, inputs.linear_ramp(6000, D, esz), 2048, True)
    finally:
        spz.set_kernel_path(old)


@pytest.mark.gpu
@pytest.mark.parametrize("pred", [DELTA, XFF], ids=["delta", "xff"])
@pytest.mark.parametrize("esz,D", [(1, 1), (1, 3), (1, 80), (2, 2), (2, 33), (1, 200), (2, 1000)])
def test_gpu_frame_query_decode_then_reduce(pred, esz, D):
    """shapes without a fused kernel: decode into the workspace, reduce the decoded rows"""
    import sprintz_b200 as spz
    spz.lib()
    for chunk_rows, nrows, ragged in ((16, 400, 0), (256, 5000, 3), (4096, 20000, 7)):
        x = inputs.ar1(nrows, D, esz, seed=nrows)
        if ragged:
            x = x[:-ragged]
        _frame_query_case(spz, pred, esz, D, x, chunk_rows, False)


@pytest.mark.gpu
@pytest.mark.parametrize("esz,D", [(1, 1), (1, 3), (1, 7), (2, 5), (1, 48), (2, 100), (1, 1024), (2, 1024)])
def test_gpu_reduce_rows_kernel(esz, D):
    import torch
    import sprintz_b200 as spz
    spz.lib()
    for n in (1, 15, 16, 17, 1000, 123457, 3_000_001):
        x = inputs.fuzz(n, esz, seed=n)["fuzz256"]
        xt = torch.from_numpy(x.view(np.uint8 if esz == 1 else np.int16)).cuda()
        for op in (Q_MAX, Q_SUM):
            got = spz.reduce_rows_dev(op, xt, D).cpu().numpy().view(np.uint64)
            assert np.array_equal(got, _np_reduce(x, D, op)), (esz, D, n, op)


@pytest.mark.gpu
@pytest.mark.parametrize("pred,esz,D,rows", [(XFF, 1, 8, 1 << 27), (XFF, 2, 16, 1 << 26), (DELTA, 1, 80, 1 << 22)],
                         ids=["c2-1GiB", "c3-2GiB", "c4-320MiB"])
def test_gpu_frame_query_full_size(pred, esz, D, rows):
    """BASELINE-sized frames: the fused query (c2, c3) / decode+reduce (80 dims) equals a reduction of the input."""
    import torch
    import sprintz_b200 as spz
    from sprintz_b200.synth import ar1_stream
    x = ar1_stream(rows, D, esz, seed=12345, device="cuda")
    fr = spz.frame_encode(pred, x, D, 1024)
    xv = x.view(-1, D)
    xu = xv.to(torch.int64) & ((1 << (8 * esz)) - 1)
    for op in (Q_MAX, Q_SUM):
        res, _ = spz.frame_query(fr, op, materialize=False)
        torch.cuda.synchronize()
        assert spz.last_query_fused() == (D in (8, 16))
        want = xu.max(0).values if op == Q_MAX else xu.sum(0)
        assert torch.equal(res, want)


@pytest.mark.gpu
@pytest.mark.parametrize("pred", [DELTA, XFF], ids=["delta", "xff"])
@pytest.mark.parametrize("esz,D", [(1, 1), (1, 2), (1, 3), (1, 4), (2, 1), (2, 2)])
def test_gpu_frame_query_fused_lowdim_lane_path(pred, esz, D):
    """the low-dim lane decoder (univariate / few-variable frames) folds rows the same way"""
    import sprintz_b200 as spz
    spz.lib()
    old = spz.set_kernel_path(2)
    try:
        udt = np.uint8 if esz == 1 else np.uint16
        for chunk_rows, nrows, ragged in ((16, 400, 0), (256, 5000, 3 if D > 1 else 1), (4096, 20000, 0), (256, 100, 0),
                                          (64, 64 * 300, 0)):
            x = inputs.ar1(nrows, D, esz, seed=nrows)
            if ragged:
                x = x[:-ragged]
            _frame_query_case(spz, pred, esz, D, x, chunk_rows, True)
        _frame_query_case(spz, pred, esz, D, inputs.random_walk_with_runs(3000, D, esz, seed=5), 512, True)
        _frame_query_case(spz, pred, esz, D, np.zeros(4096 * D, udt), 1024, True)
        x = np.full(70000 * D, (1 << (8 * esz)) - 2, udt)
        x[-500 * D:] = inputs.fuzz(500 * D, esz)["fuzz256"]
        _frame_query_case(spz, pred, esz, D, x, 35008, True)
        x = ((1 << (8 * esz)) - 1 - inputs.fuzz(40000 * D, esz)["fuzz16"]).astype(udt)
        _frame_query_case(spz, pred, esz, D, x, 40000, True)
        _frame_query_case(spz, pred, esz, D, inputs.linear_ramp(6000, D, esz), 2048, True)
    finally:
        spz.set_kernel_path(old)
