"""Probe of the EXPERIMENTAL overlapped frame encode (spz_chase_kernel) at BASELINE sizes -- the experiment round 1 ran
out of GPU budget for.  Mode 2 lifts the 148-tile cap (unverified there: waits are bounded, a failed run flags the frame
with magic 0 and takes seconds, it cannot hang).  Prints classic vs overlapped encode time and checks the round trip.
usage: timeout 120 python tools/overlap_probe.py [c2|c3] [mib]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
import sprintz_b200 as spz  # noqa: E402

cfg = dict(bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "c2"])
if len(sys.argv) > 2:
    cfg["nbytes"] = int(sys.argv[2]) << 20
x, rows = bench.make_stream(cfg, torch.device("cuda", 0), seed=12345)
codec = spz.FrameCodec(cfg["pred"], cfg["esz"], x.numel(), cfg["D"], cfg["chunk_rows"])
out = torch.empty_like(x)
for mode in (0, 2, 0, 2):
    spz.set_encode_overlap(mode)
    fr = codec.encode(x)                      # the first launch of an instantiation always takes the classic path
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    t0 = time.time()
    ev[0].record()
    for _ in range(3):
        fr = codec.encode(x)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 3
    ok = None
    try:
        codec.decode(fr, out)
        torch.cuda.synchronize()
        codec.check_status()
        ok = bool(torch.equal(out, x)) and fr.header().magic == 0x315A5053
    except Exception as e:  # noqa: BLE001
        ok = f"{type(e).__name__}: {e}"
    print(f"mode {mode}: overlapped={spz.last_encode_overlapped()} nchunks={codec.nchunks} encode {ms:.3f} ms "
          f"({x.numel() * cfg['esz'] / ms / 1e6:.0f} GB/s) wall {time.time() - t0:.2f} s round trip {ok}", flush=True)
spz.set_encode_overlap(0)
