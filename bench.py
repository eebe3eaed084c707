#!/usr/bin/env python
"""bench.py -- encode+decode throughput of the Sprintz frame path on B200 (see DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W [--config c2] [--impl reference]

A *step* is one pass of the hot path over one batch: frame-encode the rank's stream, then frame-decode it.
metric  = raw bytes / (t_encode + t_decode), whole job over N GPUs (weak scaling: every rank owns a full
          copy of the per-GPU workload, sharded by chunks with no data-path collective).
value   : inputs and outputs resident in HBM, CUDA events around exactly K steps, max over ranks.
e2e     : same metric through the host-buffer C ABI (spz_frame_compress_host / spz_frame_decompress_host):
          pinned host input -> H2D -> kernels -> D2H, every step.
roofline: dominant codec kernel, algorithmic bytes N*(1+1/ratio) / its CUDA-event time, against the measured
          HBM copy bandwidth in MEASURED_PEAKS.json.
cpu_baseline / --impl reference: the UNMODIFIED reference compiled from its own sources (oracle/_ref) on
          the box's host cores, all threads, over the same chunk slices.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# (name, elem_sz, ndims, predictor, bytes per GPU, default chunk_rows)
CONFIGS = {
    "c1": dict(desc="int8, 1-dim, delta+bitpack, 64 KB", esz=1, D=1, pred=0, nbytes=64 << 10, chunk_rows=2048),
    # 512-row chunks: 262144 lane tasks = 3.5 waves of resident lanes instead of 1.7 (+4 % decode), -0.8 % ratio
    "c2": dict(desc="int8, 8-dim rowmajor, FIRE+bitpack, 1 GB synthetic AR(1)", esz=1, D=8, pred=1,
               nbytes=1 << 30, chunk_rows=512),
    "c3": dict(desc="int16, 16-dim rowmajor, FIRE+bitpack, 4 GB synthetic AR(1)", esz=2, D=16, pred=1,
               nbytes=4 << 30, chunk_rows=1024),
    "c4": dict(desc="int8, 80-dim, delta+bitpack, 4 GB per GPU (32 GB over 8)", esz=1, D=80, pred=0,
               nbytes=4 << 30, chunk_rows=1024),
    # the non-RLE sibling codec compress_rowmajor_delta_8b (frame codec id 2), generic kernels
    "f1": dict(desc="int8, 8-dim rowmajor, delta+bitpack without RLE (compress_rowmajor_delta_8b), 1 GB synthetic AR(1)",
               esz=1, D=8, pred=2, nbytes=1 << 30, chunk_rows=1024),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(cfg_name, kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture of this config (or None)."""
    for name in ("r1b_traffic.json", "r1_traffic.json"):        # latest capture first
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                return json.load(f)["configs"][cfg_name][kernel]["traffic"]
        except Exception:
            continue
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_stream(cfg, device, seed):
    from sprintz_b200.synth import ar1_stream
    rows = cfg["nbytes"] // (cfg["esz"] * cfg["D"])
    rows -= rows % 16
    return ar1_stream(rows, cfg["D"], cfg["esz"], seed=seed, device=device), rows


# ------------------------------------------------------------------ CPU reference timing (oracle/_ref)

def time_reference(cfg, x_host, chunk_rows, nthreads, repeats=3):
    """Times the unmodified reference on host cores over the same chunk slices. -> dict"""
    import numpy as np
    from oracle import Ref
    R = Ref()
    esz, D, pred = cfg["esz"], cfg["D"], cfg["pred"]
    n = x_host.size
    chunk_len = chunk_rows * D
    nchunks = (n + chunk_len - 1) // chunk_len
    stride = R.max_bytes(esz, chunk_len, D) + 64
    stride = (stride + 63) // 64 * 64
    src = np.concatenate([x_host.view(np.uint8).reshape(-1), np.zeros(4096, np.uint8)])
    comp = np.zeros(nchunks * stride + 4096, np.uint8)
    dec = np.zeros(n * esz + 4096, np.uint8)
    rets = np.zeros(nchunks, np.int64)
    rets2 = np.zeros(nchunks, np.int64)
    te, td = [], []
    for _ in range(repeats):
        te.append(R.L.ref_compress_chunks(pred, esz, src.ctypes.data, n, chunk_len, D, comp.ctypes.data, stride,
                                          rets.ctypes.data, nthreads))
        td.append(R.L.ref_decompress_chunks(pred, esz, comp.ctypes.data, stride, nchunks, n, chunk_len,
                                            dec.ctypes.data, rets2.ctypes.data, nthreads))
    ok = bool(np.array_equal(dec[: n * esz], src[: n * esz]))
    raw = n * esz
    comp_bytes = int(rets.sum()) * esz
    return dict(t_enc=min(te), t_dec=min(td), t_enc_all=te, t_dec_all=td, raw=raw, ratio=raw / max(1, comp_bytes),
                roundtrip_ok=ok, nchunks=int(nchunks))


def run_reference(args, cfg, cfg_name):
    import numpy as np
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    ncores = os.cpu_count() or 1
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    x, rows = make_stream(cfg, dev, seed=12345)
    xh = x.cpu().numpy()
    xh = xh.view(np.uint8 if cfg["esz"] == 1 else np.uint16)
    del x
    steps, warm = args.steps, args.warmup
    chunk_rows = args.chunk_rows or cfg["chunk_rows"]
    ts = []
    res = None
    for i in range(warm + steps):
        res = time_reference(cfg, xh, chunk_rows, ncores, repeats=1)
        if i >= warm:
            ts.append(res["t_enc"] + res["t_dec"])
    raw = res["raw"]
    total = sum(ts)
    val = raw * steps / total / 1e9
    line = {
        "impl": "reference", "metric": "encode+decode GB/s", "value": val, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": 1e3 * total / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8" if cfg["esz"] == 1 else "u16", "data": "synthetic",
        "config": {"workload": cfg_name + ": " + cfg["desc"], "elem_bits": 8 * cfg["esz"], "ndims": cfg["D"],
                   "predictor": ("delta", "xff", "rowmajor_delta", "rowmajor_plain")[cfg["pred"]], "bytes": raw, "chunk_rows": chunk_rows,
                   "threads": ncores},
        "cpu_baseline": {"value": val, "unit": "GB/s", "cores": ncores, "kind": "reference",
                         "sample": f"whole {raw >> 20} MiB workload per step, {res['nchunks']} chunk slices, "
                                   f"std::thread pool over oracle/_ref (unmodified reference, g++ -O3 -mavx2 -mbmi2)",
                         "encode_gbs": raw / min(t for t in [res["t_enc"]]) / 1e9,
                         "decode_gbs": raw / res["t_dec"] / 1e9, "ratio": res["ratio"],
                         "roundtrip_ok": res["roundtrip_ok"]},
        "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line, args.real_stdout)
    return 0


# ------------------------------------------------------------------ the GPU arm

def run_ours(args, cfg, cfg_name):
    import numpy as np
    import torch
    import torch.distributed as dist
    import sprintz_b200 as spz

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the Sprintz B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    L = spz.lib()
    esz, D, pred = cfg["esz"], cfg["D"], cfg["pred"]
    chunk_rows = args.chunk_rows or cfg["chunk_rows"]
    x, rows = make_stream(cfg, dev, seed=12345 + rank)
    n = x.numel()
    raw = n * esz
    codec = spz.FrameCodec(pred, esz, n, D, chunk_rows, device=dev)
    out = torch.empty_like(x)
    steps, warm = args.steps, args.warmup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- correctness of exactly what is timed (round trip; per-chunk byte parity is tests/'s job)
    fr = codec.encode(x)
    codec.decode(fr, out)
    torch.cuda.synchronize()
    codec.check_status()
    if not torch.equal(out, x):
        raise SystemExit("bench.py: decode(encode(x)) != x")
    hdr = fr.header()
    payload = int(hdr.payload_bytes)
    frame_bytes = int(hdr.frame_bytes)
    ratio = raw / payload

    # clocks are sampled from the start of warm-up to the end of the timed region (the timed region alone can
    # be shorter than one nvidia-smi period); warm-up is extended until a few samples under load exist
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(warm):
        codec.decode(codec.encode(x), out)
    t_w = time.time()
    while len(sampler.lines) < 3 and time.time() - t_w < 4.0:
        codec.decode(codec.encode(x), out)
        torch.cuda.synchronize()
    barrier()

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * steps + 1)]
    L.spz_profile_enable(1)
    barrier()
    ev[0].record()
    for i in range(steps):
        fr = codec.encode(x)
        ev[2 * i + 1].record()
        codec.decode(fr, out)
        ev[2 * i + 2].record()
    barrier()
    clocks = sampler.stop()
    t_total = ev[0].elapsed_time(ev[2 * steps]) / 1e3
    t_enc = sum(ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(steps)) / 1e3
    t_dec = sum(ev[2 * i + 1].elapsed_time(ev[2 * i + 2]) for i in range(steps)) / 1e3
    ms3 = (C.c_float * 3)()
    ncalls = C.c_int(0)
    L.spz_profile_collect(ms3, C.byref(ncalls))
    L.spz_profile_enable(0)
    fam = ("spz_%s_kernel", "spz_%s_hd_kernel", "spz_%s_lane_kernel")      # which kernels the timed loop launched
    k_enc = fam[L.spz_last_kernel_family(0)] % "encode"
    k_dec = fam[L.spz_last_kernel_family(1)] % "decode"
    t_enc_kernel = ms3[0] / 1e3 / max(1, ncalls.value)     # per launch
    t_scan = ms3[1] / 1e3 / max(1, ncalls.value)
    t_gather = ms3[2] / 1e3 / max(1, ncalls.value)
    t_dec_kernel = t_dec / steps

    tt = torch.tensor([t_total, t_enc, t_dec], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_total, t_enc_m, t_dec_m = (float(v) for v in tt.tolist())

    # ---- query-on-compressed, outside the timed region: per-dim REDUCE_SUM / REDUCE_MAX over the
    # frame without materialising it (fused into the lane decoder where the shape allows); reads N/r bytes only
    query_info = None
    try:
        qres = torch.empty(D, dtype=torch.int64, device=dev)
        wb = int(L.spz_frame_query_workspace_bytes(esz, n, 0))
        qwork = out.view(torch.uint8)               # decode+reduce shapes decode into this (>= n * esz bytes)
        xu = x.view(-1, D)[: (n // D)].to(torch.int32) & ((1 << (8 * esz)) - 1) if n % D == 0 else None
        query_info = {}
        for opname, op in (("REDUCE_SUM", 2), ("REDUCE_MAX", 1)):
            def q():
                spz._lib.check(L.spz_frame_query_dev(pred, esz, op, 0, fr.buf.data_ptr(), fr.buf.numel(), n, D, chunk_rows,
                                                     None, 0, codec.status.data_ptr(), qres.data_ptr(), qwork.data_ptr(),
                                                     qwork.numel(), spz.api._stream_ptr()), "spz_frame_query_dev")
            for _ in range(2):
                q()
            torch.cuda.synchronize()
            qe = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            qe[0].record()
            for _ in range(5):
                q()
            qe[1].record()
            torch.cuda.synchronize()
            tq = qe[0].elapsed_time(qe[1]) / 5e3
            ok = None
            if xu is not None:
                want = xu.sum(0, dtype=torch.int64) if op == 2 else xu.max(0).values.to(torch.int64)
                ok = bool(torch.equal(qres, want))
            fused = bool(L.spz_last_query_fused())
            query_info[opname] = {"ms": 1e3 * tq, "fused": fused, "raw_gbs": raw / tq / 1e9,
                                  "hbm_bytes_read": payload, "hbm_read_frac": payload / tq / 1e9 / peaks()[0],
                                  "matches_reduction_of_input": ok}
        del xu
    except Exception as e:  # noqa: BLE001
        query_info = {"error": f"{type(e).__name__}: {e}"}

    # ---- optional exchange step: gather every rank's compressed payload + directory on rank 0 over
    # NCCL / NVLink.  Not part of `value` (the codec path itself needs no collective); reported beside it.
    gather_info = None
    if world > 1:
        from sprintz_b200.dist import gather_frames, shard_plan
        plan = shard_plan(world * n, D, chunk_rows, world)
        po = int(hdr.payload_off)
        pay = fr.buf[po: po + payload]
        nchunks = int(hdr.nchunks)
        sizes = fr.buf[64 + 8 * (nchunks + 1): 64 + 8 * (nchunks + 1) + 4 * nchunks].view(torch.int32).to(torch.int64)
        for it in range(5):                      # 2 warm-up + 3 timed
            if it == 2:
                barrier(); t_g0 = time.perf_counter()
            got = gather_frames(pay, sizes, plan, dst=0)
            torch.cuda.synchronize()
        barrier()
        t_g = torch.tensor([(time.perf_counter() - t_g0) / 3], dtype=torch.float64, device=dev)
        dist.all_reduce(t_g, op=dist.ReduceOp.MAX)
        moved = (world - 1) * payload            # bytes that crossed NVLink into rank 0 (payloads are ~equal)
        gather_info = {"ms": 1e3 * float(t_g.item()), "payload_bytes_total": world * payload,
                       "gbs_into_rank0": moved / float(t_g.item()) / 1e9, "backend": "nccl send/recv",
                       "note": "all-gather of chunk sizes -> exclusive scan -> variable-size send/recv to rank 0"}
        del got, pay

    # ---- end to end through the host-buffer C ABI (pinned host memory), every step H2D + D2H
    e2e_steps = max(1, min(steps, args.e2e_steps))
    xh = torch.empty(n, dtype=x.dtype, pin_memory=True)
    xh.copy_(x)
    fcap = int(L.spz_frame_max_bytes(esz, n, D, chunk_rows))
    fh = torch.empty(fcap, dtype=torch.uint8, pin_memory=True)
    oh = torch.empty(n, dtype=x.dtype, pin_memory=True)
    del codec, out
    torch.cuda.empty_cache()

    def e2e_step():
        fb = L.spz_frame_compress_host(pred, esz, xh.data_ptr(), n, D, chunk_rows, fh.data_ptr(), fcap)
        spz._lib.check(fb, "spz_frame_compress_host")
        ne = L.spz_frame_decompress_host(fh.data_ptr(), fb, oh.data_ptr(), n)
        spz._lib.check(ne, "spz_frame_decompress_host")
        return fb

    fb = e2e_step()
    if not torch.equal(oh, xh):
        raise SystemExit("bench.py: host-buffer round trip mismatch")
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        fb = e2e_step()
    torch.cuda.synchronize()
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    t_e2e = float(t_e2e.item())
    L.spz_release_workspaces()

    peak, peak_src = peaks()
    alg_bytes = raw * (1.0 + 1.0 / ratio)               # N*(1 + 1/r), both directions
    dom_is_dec = t_dec_kernel >= t_enc_kernel
    dom_t = t_dec_kernel if dom_is_dec else t_enc_kernel
    full_size = not args.bytes_mb
    roof = lambda t, kern: {"bound": "hbm", "achieved": alg_bytes / t / 1e9, "peak": peak, "unit": "GB/s",  # noqa: E731
                            "frac": alg_bytes / t / 1e9 / peak,
                            "traffic": ncu_traffic(cfg_name, kern) if full_size else None,
                            "alg_bytes_per_launch": alg_bytes, "ms_per_launch": 1e3 * t, "peak_source": peak_src,
                            "kernel": kern}

    value = world * raw * steps / t_total / 1e9
    line = {
        "metric": "encode+decode GB/s", "value": value, "unit": "GB/s", "n_gpus": world, "steps": steps,
        "warmup": warm, "ms_per_step": 1e3 * t_total / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8" if esz == 1 else "u16", "data": "synthetic",
        "config": {"workload": cfg_name + ": " + cfg["desc"], "elem_bits": 8 * esz, "ndims": D,
                   "predictor": ("delta", "xff", "rowmajor_delta", "rowmajor_plain")[pred], "bytes_per_gpu": raw, "chunk_rows": chunk_rows,
                   "nchunks_per_gpu": int(hdr.nchunks), "ratio": ratio,
                   "l2": "inputs (>= 1 GiB) exceed the 126 MB L2; no explicit flush",
                   "parallelism": f"chunk-sharded x{world}, no data-path collective"},
        "encode_gbs": world * raw * steps / t_enc_m / 1e9, "decode_gbs": world * raw * steps / t_dec_m / 1e9,
        "roofline": roof(dom_t, k_dec if dom_is_dec else k_enc),
        "roofline_decode": roof(t_dec_kernel, k_dec),
        "roofline_encode": roof(t_enc_kernel, k_enc),
        # the stricter read-only form the north star names: HBM bytes READ per launch / kernel time / peak
        "hbm_read_roofline": {"decode_frac": raw / ratio / t_dec_kernel / 1e9 / peak,
                              "encode_frac": raw / t_enc_kernel / 1e9 / peak,
                              "decode_raw_gbs_kernel": raw / t_dec_kernel / 1e9,
                              "encode_raw_gbs_kernel": raw / t_enc_kernel / 1e9},
        # overlapped: the directory scan + compaction run on a second stream WHILE the lane encoder runs; then
        # "compaction" is only the part left exposed after the encoder's last wave and "directory_scan" is ~0
        "encode_breakdown_ms": {"codec_kernel": 1e3 * t_enc_kernel, "directory_scan": 1e3 * t_scan,
                                "compaction": 1e3 * t_gather, "overlapped": bool(L.spz_last_encode_overlapped())},
        "e2e": {"value": world * raw * e2e_steps / t_e2e / 1e9, "unit": "GB/s",
                "h2d_bytes_per_step": raw + int(fb), "d2h_bytes_per_step": int(fb) + raw, "steps": e2e_steps,
                "api": "spz_frame_compress_host + spz_frame_decompress_host (pinned host buffers)"},
        "query": query_info,
        "gather": gather_info,
        "gpu_launches": steps * 4,     # encode, directory scan, compaction, decode per step
        "clocks": clocks,
    }

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            ncores = os.cpu_count() or 1
            xs = xh.numpy().view(np.uint8 if esz == 1 else np.uint16)
            sample_elems = min(n, (args.cpu_sample_mb << 20) // esz // (chunk_rows * D) * (chunk_rows * D))
            r = time_reference(cfg, xs[:sample_elems], chunk_rows, ncores, repeats=3)
            line["cpu_baseline"] = {
                "value": r["raw"] / (r["t_enc"] + r["t_dec"]) / 1e9, "unit": "GB/s", "cores": ncores,
                "kind": "reference",
                "sample": f"first {r['raw'] >> 20} MiB of the same stream, same {chunk_rows}-row chunk slices, "
                          f"best of 3, oracle/_ref (unmodified reference) on {ncores} threads",
                "encode_gbs": r["raw"] / r["t_enc"] / 1e9, "decode_gbs": r["raw"] / r["t_dec"] / 1e9,
                "ratio": r["ratio"], "roundtrip_ok": r["roundtrip_ok"]}
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "unit": "GB/s", "cores": 0, "kind": "reference",
                                    "sample": f"unavailable: {type(e).__name__}: {e}"}
    if rank == 0:
        emit(line, args.real_stdout)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    # Only the JSON line may reach stdout (NCCL / torchrun print banners there): park fd 1 on stderr until the end.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        return _main(real_stdout)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)


def emit(line, real_stdout):
    sys.stdout.flush()
    os.write(real_stdout, (json.dumps(line) + "\n").encode())


def _main(real_stdout):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--chunk-rows", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-sample-mb", type=int, default=1024)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--bytes-mb", type=int, default=0, help="override the per-GPU stream size (testing only)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    cfg = dict(CONFIGS[args.config])
    if args.bytes_mb:
        cfg["nbytes"] = args.bytes_mb << 20
        cfg["desc"] += f" [size overridden to {args.bytes_mb} MiB]"
    args.real_stdout = real_stdout
    if args.impl == "reference":
        return run_reference(args, cfg, args.config)
    return run_ours(args, cfg, args.config)


if __name__ == "__main__":
    sys.exit(main())
