"""Multi-GPU host logic: one process per GPU, chunks sharded contiguously, no data-path collective.

The codec path shards at chunk granularity (each chunk is an independent reference stream, SURVEY.md
section 8e), so encode/decode need no communication at all.  The only exchange step is optional: gathering
the compressed payloads and their directory on one rank (NCCL over NVLink on GPUs; the same code runs on
gloo for the CPU tests).  It is an all-gather of sizes, an exclusive scan, and variable-size send/recv.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


@dataclass
class ShardPlan:
    n: int               # total elements
    ndims: int
    chunk_rows: int
    world: int
    nchunks: int
    chunk_lo: List[int]  # first chunk of each rank
    chunk_hi: List[int]  # one past the last chunk of each rank

    @property
    def chunk_len(self) -> int:
        return self.chunk_rows * self.ndims

    def elem_range(self, rank: int) -> Tuple[int, int]:
        return min(self.n, self.chunk_lo[rank] * self.chunk_len), min(self.n, self.chunk_hi[rank] * self.chunk_len)

    def nchunks_of(self, rank: int) -> int:
        return self.chunk_hi[rank] - self.chunk_lo[rank]


def shard_plan(n: int, ndims: int, chunk_rows: int, world: int) -> ShardPlan:
    """Contiguous, balanced ranges of whole chunks per rank (rank k gets chunks [lo_k, hi_k))."""
    clen = chunk_rows * ndims
    nchunks = max(1, (n + clen - 1) // clen)
    base, extra = divmod(nchunks, world)
    lo, hi, cur = [], [], 0
    for r in range(world):
        cnt = base + (1 if r < extra else 0)
        lo.append(cur); hi.append(cur + cnt); cur += cnt
    return ShardPlan(n, ndims, chunk_rows, world, nchunks, lo, hi)


def gather_frames(payload: torch.Tensor, sizes: torch.Tensor, plan: ShardPlan, dst: int = 0,
                  group: Optional[dist.ProcessGroup] = None):
    """Gathers every rank's compacted payload (uint8, chunk streams 16-byte aligned back to back) and true
    chunk sizes (int64) on `dst`.  Returns (payload_all, chunk_off[nchunks+1], chunk_nbytes[nchunks]) on dst,
    None elsewhere.  chunk_off is the exclusive scan of the 16-byte-rounded sizes over ALL chunks."""
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    dev = payload.device
    # 1. exchange sizes (tiny): every rank's chunk sizes, padded to the longest shard
    maxc = max(plan.nchunks_of(r) for r in range(world))
    mine = torch.zeros(maxc, dtype=torch.int64, device=dev)
    mine[: sizes.numel()] = sizes.to(dev)
    allsz = torch.empty(world * maxc, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(allsz, mine, group=group)
    allsz = allsz.view(world, maxc)
    per_rank = [allsz[r, : plan.nchunks_of(r)] for r in range(world)]
    chunk_nbytes = torch.cat(per_rank) if per_rank else torch.zeros(0, dtype=torch.int64, device=dev)
    padded = (chunk_nbytes + 15) // 16 * 16
    # 2. exclusive scan -> where each chunk (and each rank's payload) lands
    chunk_off = torch.zeros(plan.nchunks + 1, dtype=torch.int64, device=dev)
    chunk_off[1:] = torch.cumsum(padded, 0)
    rank_bytes = [int(padded[plan.chunk_lo[r]: plan.chunk_hi[r]].sum().item()) for r in range(world)]
    assert rank_bytes[rank] == payload.numel(), (rank_bytes[rank], payload.numel())
    # 3. variable-size gather of the payloads
    if rank == dst:
        total = int(chunk_off[-1].item())
        out = torch.empty(total, dtype=torch.uint8, device=dev)
        reqs = []
        for r in range(world):
            a = int(chunk_off[plan.chunk_lo[r]].item())
            if r == rank:
                out[a: a + rank_bytes[r]] = payload
            elif rank_bytes[r]:
                reqs.append(dist.irecv(out[a: a + rank_bytes[r]], src=r, group=group))
        for q in reqs:
            q.wait()
        return out, chunk_off, chunk_nbytes
    if rank_bytes[rank]:
        dist.send(payload.contiguous(), dst=dst, group=group)
    return None
