"""Runs only the fused query-on-compressed launches of a bench config (for ncu): encode the frame once, then
REDUCE_SUM and REDUCE_MAX without materialising.  usage: python tools/query_probe.py c2|c3 [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
import sprintz_b200 as spz  # noqa: E402

cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
x, rows = bench.make_stream(cfg, torch.device("cuda", 0), seed=12345)
fr = spz.frame_encode(cfg["pred"], x, cfg["D"], cfg["chunk_rows"])
torch.cuda.synchronize()
for op, name in ((spz.Q_SUM, "REDUCE_SUM"), (spz.Q_MAX, "REDUCE_MAX")):
    for _ in range(reps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        res, _ = spz.frame_query(fr, op, materialize=False)
        ev[1].record()
        torch.cuda.synchronize()
    print(name, "fused" if spz.last_query_fused() else "decode+reduce", f"{ev[0].elapsed_time(ev[1]):.3f} ms", res[:4].tolist())
