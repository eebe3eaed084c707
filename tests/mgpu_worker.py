"""Multi-GPU parity worker (run under torchrun, one process per GPU; see tests/test_multigpu_gpu.py).

Every rank builds the same global stream, keeps its own shard (contiguous ranges, or stripes dealt round-robin),
encodes it on its GPU and the frames are gathered on rank 0 by the library's NCCL path
(spz_frame_gather_dev / spz_frame_encode_gather_dev).  Rank 0 then checks the gathered frame as a whole:
  * header and directory (exact exclusive scan of the 16-byte-rounded sizes),
  * EVERY chunk byte-identical to the unmodified reference on the same slice of the global stream (oracle/_ref),
  * decode(gathered frame) == the global stream.
Test infrastructure: uses oracle/ as the checker."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import inputs  # noqa: E402
import sprintz_b200 as spz  # noqa: E402
from oracle import Oracle, Ref  # noqa: E402


def shard_pieces(total_len, D, chunk_rows, S, world, rank):
    """[(global element lo, count)] of the pieces `rank` owns, in local order (mirrors spz_shard_len's rule)."""
    clen = chunk_rows * D
    nch = max(1, -(-total_len // clen))
    out = []
    if S == 0:
        base, extra = divmod(nch, world)
        lo = sum(base + (1 if r < extra else 0) for r in range(rank))
        n = base + (1 if rank < extra else 0)
        if n:
            out.append((lo * clen, min(total_len, (lo + n) * clen) - lo * clen))
    else:
        nstripes = -(-nch // S)
        for j in range(rank, nstripes, world):
            a = j * S * clen
            out.append((a, min(total_len, (j + 1) * S * clen) - a))
    return out


def run_case(comm, rank, world, checker, pred, esz, D, chunk_rows, nrows, ragged, S, streamed, force):
    L = spz.lib()
    dev = torch.device("cuda", torch.cuda.current_device())
    xg = inputs.ar1(nrows, D, esz, seed=nrows + D)
    if ragged:
        xg = xg[:-ragged]
    total = xg.size
    pieces = shard_pieces(total, D, chunk_rows, S, world, rank)
    xl = np.concatenate([xg[a: a + n] for a, n in pieces]) if pieces else np.zeros(0, xg.dtype)
    assert xl.size == spz.shard_len(total, D, chunk_rows, S, world, rank), (xl.size, "shard_len disagrees")
    xt = torch.from_numpy(xl.view(np.uint8 if esz == 1 else np.int16)).to(dev)
    spz.set_kernel_path(force)
    codec = spz.FrameCodec(pred, esz, xl.size, D, chunk_rows, device=dev)
    cap = int(L.spz_frame_gather_max_bytes(esz, total, D, chunk_rows)) if rank == 0 else 0
    gbuf = torch.zeros(max(cap, 16), dtype=torch.uint8, device=dev)
    st = spz.api._stream_ptr()
    if streamed == "p2p":
        # peer-memory gather: the gathered frame lands in rank 0's shared (CUDA IPC) area
        ptr = C.c_void_p(None)
        spz._lib.check(L.spz_comm_shared_alloc(comm.handle, cap, C.byref(ptr)), "spz_comm_shared_alloc")
        for rep in range(2):                    # twice: the epoch tags must make the flag words reusable
            spz._lib.check(L.spz_frame_encode_gather_p2p_dev(comm.handle, 0, pred, esz, xt.data_ptr(), total, D, chunk_rows, S,
                                                             codec.frame.data_ptr(), codec.frame_cap, codec.work.data_ptr(),
                                                             codec.work_bytes, st), "spz_frame_encode_gather_p2p_dev")
            torch.cuda.synchronize()
        if rank == 0:
            class _DevMem:
                pass
            m = _DevMem()
            m.__cuda_array_interface__ = {"shape": (cap,), "typestr": "|u1", "data": (int(ptr.value), False), "version": 2}
            gbuf = torch.as_tensor(m, device=dev)
    elif streamed:
        spz._lib.check(L.spz_frame_encode_gather_dev(comm.handle, 0, pred, esz, xt.data_ptr(), total, D, chunk_rows, S,
                                                     codec.frame.data_ptr(), codec.frame_cap, codec.work.data_ptr(),
                                                     codec.work_bytes, gbuf.data_ptr() if rank == 0 else None, cap, st),
                       "spz_frame_encode_gather_dev")
        torch.cuda.synchronize()
        # the local frame is complete and valid too
        lf = spz.Frame(codec.frame, pred, esz, xl.size, D, chunk_rows, None)
        back = torch.empty_like(xt)
        codec.decode(lf, back)
        torch.cuda.synchronize()
        codec.check_status()
        assert torch.equal(back, xt), "streamed: local frame does not decode to the shard"
    else:
        fr = codec.encode(xt)
        spz._lib.check(L.spz_frame_gather_dev(comm.handle, 0, total, S, fr.buf.data_ptr(),
                                              gbuf.data_ptr() if rank == 0 else None, cap, st), "spz_frame_gather_dev")
        torch.cuda.synchronize()
    spz.set_kernel_path(0)
    if rank != 0:
        return
    gf = spz.Frame(gbuf, pred, esz, total, D, chunk_rows, None)
    h = gf.header()
    clen = chunk_rows * D
    nch = max(1, -(-total // clen))
    assert h.len == total and h.nchunks == nch and h.ndims == D and h.chunk_rows == chunk_rows and h.predictor == pred
    off, nb, po = gf.directory()
    assert off[0] == 0 and np.array_equal(np.diff(off), (nb.astype(np.int64) + 15) // 16 * 16), "directory is not an exact scan"
    assert h.payload_bytes == off[-1] and h.frame_bytes == po + off[-1] + 16
    buf = gbuf[: int(h.frame_bytes)].cpu().numpy()
    for c in range(nch):
        want, ret = checker.compress(pred, xg[c * clen: (c + 1) * clen], D)
        got = buf[po + off[c]: po + off[c] + nb[c]]
        nref = ret * esz
        assert len(got) in (nref, nref + esz - 1) and np.array_equal(got[:nref], want[:nref]), f"chunk {c} differs from the reference"
    out = spz.frame_decode(gf)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy().view(xg.dtype), xg), "decode(gathered frame) != global stream"


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    comm = spz.Communicator.from_torch(rank, world)
    checker = Ref() if Ref.available() else Oracle()
    ncases = 0
    # (pred, esz, D, chunk_rows, nrows, ragged elems, force lane kernels)
    shapes = [(1, 2, 16, 64, 64 * 101 + 37, 0, 0), (1, 1, 8, 32, 32 * 257, 5, 2), (0, 1, 80, 64, 64 * 64 + 16, 0, 0),
              (0, 2, 5, 16, 16 * 333, 3, 0), (1, 2, 16, 16, 16 * 2000, 0, 2)]
    for pred, esz, D, chunk_rows, nrows, ragged, force in shapes:
        nch = -(-(nrows * D - ragged) // (chunk_rows * D))
        for S, streamed in ((0, False), (max(7, -(-nch // (20 * world))), False), (max(1, nch // (3 * world)), True),
                            (max(1, nch // (8 * world) + 1), True), (max(1, nch // (3 * world)), "p2p"),
                            (max(1, nch // (8 * world) + 1), "p2p")):
            run_case(comm, rank, world, checker, pred, esz, D, chunk_rows, nrows, ragged, S, streamed, force)
            ncases += 1
            dist.barrier()
    comm.close()
    dist.barrier()
    if rank == 0:
        print(f"MGPU OK {ncases} cases world={world} checker={checker.name}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
