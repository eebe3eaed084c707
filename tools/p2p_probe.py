"""Multi-GPU probe (torchrun, one process per GPU): encode + peer-memory gather against the number of stripes per rank.
usage: torchrun ... tools/p2p_probe.py c3|c4|c2 [rounds ...]"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import sprintz_b200 as spz  # noqa: E402


def main():
    cfg_name = sys.argv[1] if len(sys.argv) > 1 else "c3"
    rounds_list = [int(a) for a in sys.argv[2:]] or [1, 2, 4, 8, 16]
    env = bench.Env()
    dist.init_process_group("nccl", device_id=env.dev)
    env.comm = spz.Communicator.from_torch(env.rank, env.world)
    cfg = bench.CONFIGS[cfg_name]
    chunk_rows = cfg.get("chunk_rows", 512)
    esz, D, pred = cfg["esz"], cfg["D"], cfg["pred"]
    x, rows = bench.make_stream(cfg, env.dev, seed=12345 + env.rank, multiple=chunk_rows * 16)
    codec = spz.FrameCodec(pred, esz, x.numel(), D, chunk_rows, device=env.dev)
    fr = codec.encode(x)
    torch.cuda.synchronize()
    payload = int(fr.header().payload_bytes)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    env.barrier()
    e0.record()
    for _ in range(5):
        codec.encode(x)
    e1.record()
    env.barrier()
    t_enc = e0.elapsed_time(e1) / 5
    for rounds in rounds_list:
        if codec.nchunks % rounds:
            continue
        g = spz.FrameGather(env.comm, codec, dst=0, stripe_chunks=codec.nchunks // rounds)
        g.enable_p2p()
        g.encode_and_gather_p2p(x)
        torch.cuda.synchronize()
        env.barrier()
        ok = g.verify_against_shards(g.last_frame, x) if env.rank == 0 else None
        env.barrier()
        e0.record()
        for _ in range(5):
            g.encode_and_gather_p2p(x)
        e1.record()
        env.barrier()
        t = env.max_over_ranks([e0.elapsed_time(e1) / 5])[0]
        if env.rank == 0:
            print(json.dumps({"config": cfg_name, "world": env.world, "rounds": rounds, "encode_ms": round(t_enc, 3),
                              "encode_gather_p2p_ms": round(t, 3), "payload_per_rank_gb": round(payload / 1e9, 3),
                              "ingress_gbs_if_transfer_bound": round((env.world - 1) * payload / t / 1e6, 1), "ok": ok}), flush=True)
        del g
    env.comm.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
