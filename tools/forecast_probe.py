"""Throughput of the stand-alone forecaster transforms (the paper's "FIRE speed" figure,
communicate/results.tex:186): delta / double delta (element-wise encode, chained-scan decode) and the FIRE transform
(predict.h:15-30 -- ONE stream is serial in rows, so it offers ndims-way parallelism only; the chunked frame codec is that
transform at full parallelism).  One JSON line per point.  usage: python tools/forecast_probe.py [--mb 256]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import sprintz_b200 as spz  # noqa: E402

mb = int(sys.argv[sys.argv.index("--mb") + 1]) if "--mb" in sys.argv else 256
dev = torch.device("cuda", 0)
L = spz.lib()
peak, _ = bench.peaks()


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps / 1e3


for esz, D in ((1, 8), (2, 16), (1, 80)):
    x, rows = bench.make_stream(dict(esz=esz, D=D, nbytes=mb << 20), dev, seed=12345)
    raw = x.numel() * esz
    for order, name in ((1, "delta"), (2, "doubledelta")):
        y = spz.forecast_encode_dev(order, x, D)
        back = spz.forecast_decode_dev(order, y, D)
        torch.cuda.synchronize()
        ok = bool(torch.equal(back, x))
        work = torch.empty(int(L.spz_forecast_workspace_bytes(x.numel(), D)), dtype=torch.uint8, device=dev)
        te = timed(lambda: spz.forecast_encode_dev(order, x, D))
        td = timed(lambda: spz.forecast_decode_dev(order, y, D, work))
        print(json.dumps({"transform": name, "elem_bits": 8 * esz, "ndims": D, "raw_mib": raw >> 20, "roundtrip_ok": ok,
                          "encode_gbs": round(raw / te / 1e9, 1), "decode_gbs": round(raw / td / 1e9, 1),
                          "encode_frac_of_copy_peak": round(2 * raw / te / 1e9 / peak, 3),
                          "decode_frac_of_copy_peak": round(2 * raw / td / 1e9 / peak, 3)}), flush=True)
    # the stand-alone FIRE transform: one serial stream (16 MiB is plenty to time it)
    n = min(x.numel(), (16 << 20) // esz)
    n -= n % D
    xs = x[:n].contiguous()
    ys = torch.empty_like(xs)
    zs = torch.empty_like(xs)

    def enc():
        spz._lib.check(L.spz_xff_forecast_dev(1, esz, xs.data_ptr(), n, D, ys.data_ptr(), spz.api._stream_ptr()), "xff enc")

    def dec():
        spz._lib.check(L.spz_xff_forecast_dev(0, esz, ys.data_ptr(), n, D, zs.data_ptr(), spz.api._stream_ptr()), "xff dec")

    te, td = timed(enc, 2), timed(dec, 2)
    print(json.dumps({"transform": "xff (stand-alone, one serial stream)", "elem_bits": 8 * esz, "ndims": D,
                      "raw_mib": (n * esz) >> 20, "roundtrip_ok": bool(torch.equal(zs, xs)),
                      "encode_gbs": round(n * esz / te / 1e9, 3), "decode_gbs": round(n * esz / td / 1e9, 3),
                      "note": "ndims-way parallelism only; the frame codec (FIRE + bit packing over independent chunks) runs the same transform at 1.2-2.5 TB/s"}),
          flush=True)
