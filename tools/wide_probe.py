"""One frame encode + decode per (element width, predictor) on the wide-row kernels, for an ncu capture:
  ncu --set full --clock-control none --import-source on -k regex:spz_.*code_wide -f -o gpurun_out/wide python tools/wide_probe.py
usage: python tools/wide_probe.py [--dims 80] [--mb 1024] [--chunk-rows 512]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
import sprintz_b200 as spz  # noqa: E402


def arg(name, default):
    return sys.argv[sys.argv.index(name) + 1] if name in sys.argv else default


D, mb, chunk_rows = int(arg("--dims", 80)), int(arg("--mb", 1024)), int(arg("--chunk-rows", 512))
dev = torch.device("cuda:0")
for esz in (1, 2):
    x, rows = bench.make_stream(dict(esz=esz, D=D, nbytes=mb << 20), dev, seed=12345, multiple=1024)
    out = torch.empty_like(x)
    for pred in (0, 1):
        codec = spz.FrameCodec(pred, esz, x.numel(), D, chunk_rows, device=dev)
        fr = codec.encode(x)
        codec.decode(fr, out)
        torch.cuda.synchronize()
        assert torch.equal(out, x)
        print(esz, pred, "ratio", x.numel() * esz / int(fr.header().payload_bytes), flush=True)
