"""chunk_rows sweep of one BASELINE config on one GPU: ratio, codec-kernel / scan / compaction / decode times.
The lane kernels run whole chunks per lane, so the number of resident-lane "rounds" (nchunks * lanes-per-chunk /
(148 SMs * CTAs/SM * 128)) quantises the run time; this tool finds the chunk_rows where that quantisation and the
per-chunk ratio cost balance.  usage: python tools/rows_sweep.py c3 256,384,512,768,896,1024 [--mb N]"""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
import sprintz_b200 as spz  # noqa: E402

cfg_name = sys.argv[1]
rows_list = [int(v) for v in sys.argv[2].split(",")]
cfg = dict(bench.CONFIGS[cfg_name])
if "--mb" in sys.argv:
    cfg["nbytes"] = int(sys.argv[sys.argv.index("--mb") + 1]) << 20
peak, _ = bench.peaks()
L = spz.lib()
if "--pipeline" in sys.argv:
    L.spz_set_encode_pipeline(int(sys.argv[sys.argv.index("--pipeline") + 1]))
dev = torch.device("cuda", 0)
x, rows = bench.make_stream(cfg, dev, seed=12345)
n = x.numel()
esz, D, pred = cfg["esz"], cfg["D"], cfg["pred"]
raw = n * esz
out = torch.empty_like(x)
FAM = ("generic", "team(hd)", "lane", "wide")
for R in rows_list:
    codec = spz.FrameCodec(pred, esz, n, D, R, device=dev)
    for _ in range(3):
        fr = codec.encode(x); codec.decode(fr, out)
    torch.cuda.synchronize()
    ok = bool(torch.equal(out, x))
    payload = int(fr.header().payload_bytes)
    reps = 10
    L.spz_profile_enable(1)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    for _ in range(reps):
        fr = codec.encode(x)
    ev[1].record()
    for _ in range(reps):
        codec.decode(fr, out)
    ev[2].record()
    torch.cuda.synchronize()
    ms3 = (C.c_float * 3)(); nc = C.c_int(0)
    L.spz_profile_collect(ms3, C.byref(nc)); L.spz_profile_enable(0)
    te, td = ev[0].elapsed_time(ev[1]) / reps, ev[1].elapsed_time(ev[2]) / reps
    k = max(1, nc.value)
    alg = raw + payload
    print(json.dumps({"config": cfg_name, "chunk_rows": R, "nchunks": codec.nchunks, "slabs": int(L.spz_last_encode_slabs()), "ratio": round(raw / payload, 5), "ok": ok,
                      "enc_family": FAM[L.spz_last_kernel_family(0)], "dec_family": FAM[L.spz_last_kernel_family(1)],
                      "enc_ms": round(te, 4), "enc_codec_ms": round(ms3[0] / k, 4), "scan_ms": round(ms3[1] / k, 4),
                      "gather_ms": round(ms3[2] / k, 4), "dec_ms": round(td, 4),
                      "enc_frac_codec": round(alg / (ms3[0] / k / 1e3) / 1e9 / peak, 4),
                      "enc_frac_whole": round(alg / (te / 1e3) / 1e9 / peak, 4),
                      "dec_frac": round(alg / (td / 1e3) / 1e9 / peak, 4)}), flush=True)
    del codec, fr
    torch.cuda.empty_cache()
