"""CPU test (gloo, world_size 2) of the multi-GPU host logic: contiguous chunk sharding, the size exchange,
the exclusive scan that places each rank's payload, and the variable-size gather.  The payloads here are
produced by the oracle (test infrastructure) standing in for each rank's GPU encoder."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import inputs


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, esz, D, pred, nrows, chunk_rows, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import Oracle
        from sprintz_b200 import dist as sdist
        O = Oracle()
        x = inputs.ar1(nrows, D, esz, seed=3)
        plan = sdist.shard_plan(x.size, D, chunk_rows, world)
        lo, hi = plan.elem_range(rank)
        # "encode" my shard chunk by chunk
        clen = chunk_rows * D
        streams = []
        for c in range(plan.chunk_lo[rank], plan.chunk_hi[rank]):
            s, _ = O.compress(pred, x[c * clen: min(x.size, (c + 1) * clen)], D)
            streams.append(s)
        sizes = np.array([len(s) for s in streams], np.int64)
        payload = np.concatenate([np.pad(s, (0, (-len(s)) % 16)) for s in streams]) if streams else np.zeros(0, np.uint8)
        gathered = sdist.gather_frames(torch.from_numpy(payload), torch.from_numpy(sizes), plan, dst=0)
        if rank == 0:
            pay, chunk_off, chunk_nbytes = gathered
            q.put((pay.numpy(), chunk_off.numpy(), chunk_nbytes.numpy(), lo, hi))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("esz,D,pred", [(1, 8, 1), (2, 16, 1), (1, 80, 0)])
def test_two_rank_gather_matches_single_stream(oracle, esz, D, pred):
    nrows, chunk_rows, world = 1000, 64, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, esz, D, pred, nrows, chunk_rows, q)) for r in range(world)]
    for p in procs:
        p.start()
    pay, chunk_off, chunk_nbytes, lo, hi = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x = inputs.ar1(nrows, D, esz, seed=3)
    clen = chunk_rows * D
    nch = (x.size + clen - 1) // clen
    assert len(chunk_nbytes) == nch and len(chunk_off) == nch + 1
    assert np.array_equal(np.diff(chunk_off), (chunk_nbytes + 15) // 16 * 16)
    for c in range(nch):
        want, _ = oracle.compress(pred, x[c * clen: (c + 1) * clen], D)
        got = pay[chunk_off[c]: chunk_off[c] + chunk_nbytes[c]]
        assert np.array_equal(got, want), c
        d, n = oracle.decompress(pred, esz, got, len(x[c * clen: (c + 1) * clen]))
        assert np.array_equal(d, x[c * clen: (c + 1) * clen])


def test_shard_plan_covers_everything():
    from sprintz_b200 import dist as sdist
    for n, D, cr, world in ((1000 * 8, 8, 64, 2), (12345, 5, 16, 8), (100, 80, 256, 4), (1 << 20, 16, 1024, 8)):
        plan = sdist.shard_plan(n, D, cr, world)
        assert plan.chunk_lo[0] == 0 and plan.chunk_hi[-1] == plan.nchunks
        covered = 0
        for r in range(world):
            assert plan.chunk_lo[r] <= plan.chunk_hi[r]
            if r:
                assert plan.chunk_lo[r] == plan.chunk_hi[r - 1]
            lo, hi = plan.elem_range(r)
            covered += hi - lo
        assert covered == n
