"""BASELINE config 5 at the reference's block size (8 samples): frame encode / decode throughput over ndims x
element width x predictor on one GPU, 256 MiB of synthetic AR(1) per point, chunk_rows = 1024.
Prints one JSON line per point (kernel family, GB/s of raw data, roofline fraction N*(1+1/r)/t/peak).
usage: python tools/dims_sweep.py [--mb 256] [--dims 8,16] [--force 0|1|2] > profiles/<round>_dims_sweep.jsonl"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
import sprintz_b200 as spz  # noqa: E402

mb = int(sys.argv[sys.argv.index("--mb") + 1]) if "--mb" in sys.argv else 256
dims = ([int(v) for v in sys.argv[sys.argv.index("--dims") + 1].split(",")] if "--dims" in sys.argv
        else [1, 2, 3, 4, 5, 8, 16, 24, 32, 48, 64, 80, 128, 256])
force = int(sys.argv[sys.argv.index("--force") + 1]) if "--force" in sys.argv else 0   # spz_set_kernel_path
peak, _ = bench.peaks()
L = spz.lib()
L.spz_set_kernel_path(force)
FAM = ("generic", "team(hd)", "lane")
dev = torch.device("cuda", 0)
for esz in (1, 2):
    for D in dims:
        cfg = dict(esz=esz, D=D, nbytes=mb << 20)
        x, rows = bench.make_stream(cfg, dev, seed=12345)
        n = x.numel()
        raw = n * esz
        out = torch.empty_like(x)
        for pred, pname in ((0, "delta"), (1, "xff")):
            codec = spz.FrameCodec(pred, esz, n, D, 1024, device=dev)
            for _ in range(2):
                fr = codec.encode(x); codec.decode(fr, out)
            torch.cuda.synchronize()
            ok = bool(torch.equal(out, x))
            payload = int(fr.header().payload_bytes)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            reps = 5
            ev[0].record()
            for _ in range(reps):
                fr = codec.encode(x)
            ev[1].record()
            for _ in range(reps):
                codec.decode(fr, out)
            ev[2].record()
            torch.cuda.synchronize()
            te, td = ev[0].elapsed_time(ev[1]) / reps / 1e3, ev[1].elapsed_time(ev[2]) / reps / 1e3
            alg = raw + payload
            print(json.dumps({"elem_bits": 8 * esz, "ndims": D, "predictor": pname, "raw_mib": raw >> 20, "nchunks": codec.nchunks, "kernel_path": force,
                              "ratio": round(raw / payload, 4), "roundtrip_ok": ok,
                              "encode_kernels": FAM[L.spz_last_kernel_family(0)], "decode_kernels": FAM[L.spz_last_kernel_family(1)],
                              "encode_gbs": round(raw / te / 1e9, 1), "decode_gbs": round(raw / td / 1e9, 1),
                              "encode_frac": round(alg / te / 1e9 / peak, 3), "decode_frac": round(alg / td / 1e9 / peak, 3)}),
                  flush=True)
            del codec, fr
        del x, out
        torch.cuda.empty_cache()
