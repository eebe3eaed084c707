#!/usr/bin/env python
"""bench.py -- encode+decode throughput of the Sprintz frame path on B200 (see DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W [--config c3] [--impl reference]

A *step* is one pass of the hot path over one batch: frame-encode the rank's stream, then frame-decode it.
The default workload is BASELINE.json's north-star configuration c3 (int16, 16-dim row-major, FIRE+bitpack,
4 GiB per GPU -- also the largest single-GPU configuration); c2 and c4 ride along as sub-records in the same line.

metric  = raw bytes / (t_encode + t_decode), whole job over N GPUs (weak scaling: every rank owns a full
          copy of the per-GPU workload, sharded by chunks with no data-path collective).
value   : inputs and outputs resident in HBM, CUDA events around exactly K steps, max over ranks.
e2e     : same metric through the host-buffer C ABI (spz_frame_compress_host / spz_frame_decompress_host):
          pinned host input -> H2D -> kernels -> D2H, every step.
roofline: dominant codec kernel, algorithmic bytes N*(1+1/ratio) / its CUDA-event time, against the measured
          HBM copy bandwidth in MEASURED_PEAKS.json (SURVEY 8d's form; the stricter read-only form the north star
          names is reported beside it as hbm_read_roofline).
parity  : EVERY chunk stream of the timed frame is compared with the bytes the unmodified reference
          (oracle/_ref) produces for the same slice (N = 1: the whole stream; N > 1: a bounded sample on rank 0).
cpu_baseline / --impl reference: the UNMODIFIED reference compiled from its own sources (oracle/_ref) on
          the box's host cores, all threads, over the same chunk slices; plus one thread on one whole stream
          (the paper's setting).
N > 1   : sub-record c4 shards the 32 GB int8 x 80 stream (4 GiB per GPU) and gathers the compressed frames on
          rank 0 over NCCL (ncclSend/ncclRecv inside libsprintz_b200.so): `value_with_gather` has the collective
          inside the timed region.
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

PRED_NAMES = ("delta", "xff", "rowmajor_delta", "rowmajor_plain", "rowmajor_xff")

# (name, elem_sz, ndims, predictor, bytes per GPU, default chunk_rows)
CONFIGS = {
    "c1": dict(desc="int8, 1-dim, delta+bitpack, 64 KB", esz=1, D=1, pred=0, nbytes=64 << 10, chunk_rows=2048),
    "c2": dict(desc="int8, 8-dim rowmajor, FIRE+bitpack, 1 GB synthetic AR(1)", esz=1, D=8, pred=1,
               nbytes=1 << 30, chunk_rows=512),
    # chunk_rows: the lane kernels run whole chunks per lane, so run time is quantised in machine "rounds"; 512 rows give
    # 6.9 (decode) rounds instead of 3.5 at 1024 -- decode 0.623 -> 0.644 of the roofline for 0.5 % of compression ratio
    # (profiles/r2_rows_sweep_c3.jsonl)
    "c3": dict(desc="int16, 16-dim rowmajor, FIRE+bitpack, 4 GB synthetic AR(1)", esz=2, D=16, pred=1,
               nbytes=4 << 30, chunk_rows=512),
    "c4": dict(desc="int8, 80-dim, delta+bitpack, 4 GB per GPU (32 GB over 8)", esz=1, D=80, pred=0,
               nbytes=4 << 30, chunk_rows=512),
    # the non-RLE sibling codec compress_rowmajor_delta_8b (frame codec id 2)
    "f1": dict(desc="int8, 8-dim rowmajor, delta+bitpack without RLE (compress_rowmajor_delta_8b), 1 GB synthetic AR(1)",
               esz=1, D=8, pred=2, nbytes=1 << 30, chunk_rows=1024),
}
MAIN_CONFIG = "c3"
SUB_CONFIGS = ("c2", "c4")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(cfg_name, kernel, chunk_rows):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture of this config AND chunk_rows
    (profiles/r2_traffic.json); None when no capture of exactly this launch shape exists."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
            rec = json.load(f)["configs"][cfg_name]
        if int(rec.get("chunk_rows", -1)) != int(chunk_rows):
            return None
        return rec[kernel]["traffic"]
    except Exception:
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


GATHER_ROUNDS = 8       # stripes per rank of the streamed multi-GPU gather (NCCL)
P2P_ROUNDS = 16         # stripes per rank of the peer-memory gather (profiles/r2_p2p_probe_n2.jsonl: 16 > 8 > 4)


def make_stream(cfg, device, seed, multiple=16):
    from sprintz_b200.synth import ar1_stream
    rows = cfg["nbytes"] // (cfg["esz"] * cfg["D"])
    rows -= rows % multiple
    return ar1_stream(rows, cfg["D"], cfg["esz"], seed=seed, device=device), rows


def config_record(cfg_name, cfg, raw, chunk_rows, world):
    """The `config` object of the JSON line: identical for both arms of the same workload."""
    clen = chunk_rows * cfg["D"]
    return {"workload": cfg_name + ": " + cfg["desc"], "elem_bits": 8 * cfg["esz"], "ndims": cfg["D"],
            "predictor": PRED_NAMES[cfg["pred"]], "bytes_per_gpu": raw, "chunk_rows": chunk_rows,
            "nchunks_per_gpu": (raw // cfg["esz"] + clen - 1) // clen,
            "l2": "inputs (>= 1 GiB) exceed the 126 MB L2; no explicit flush",
            "parallelism": f"chunk-sharded x{world}, no data-path collective"}


# ------------------------------------------------------------------ CPU reference (oracle/_ref): timing + parity source

def reference_chunks(cfg, x_host, chunk_rows, nthreads, repeats=1, keep=False):
    """Runs the unmodified reference on host cores over the chunk slices of x_host (numpy). -> dict
    keep=True also returns the compressed bytes (comp, stride, rets) for the parity check."""
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
    out = dict(t_enc=min(te), t_dec=min(td), t_enc_all=te, t_dec_all=td, raw=raw, comp_bytes=comp_bytes,
               ratio=raw / max(1, comp_bytes), roundtrip_ok=ok, nchunks=int(nchunks))
    if keep:
        out.update(comp=comp, stride=stride, rets=rets)
    return out


def reference_single_stream(cfg, x_host, max_mb=512):
    """ONE thread, ONE reference call on one whole stream (the paper's setting): the first
    `max_mb` MiB of the workload as a single sprintz_compress_* / sprintz_decompress_* call."""
    import numpy as np
    from oracle import Ref
    R = Ref()
    esz, D, pred = cfg["esz"], cfg["D"], cfg["pred"]
    n = min(x_host.size, (max_mb << 20) // esz // (16 * D) * (16 * D), 0xffffff00 // (16 * D) * (16 * D))
    src = np.concatenate([x_host[:n].view(np.uint8).reshape(-1), np.zeros(4096, np.uint8)])
    comp = np.zeros(R.max_bytes(esz, n, D) + 4096, np.uint8)
    dec = np.zeros(n * esz + 4096, np.uint8)
    best_e, best_d, ret = 1e30, 1e30, 0
    for _ in range(2):
        t0 = time.perf_counter()
        ret = int(R.L.ref_compress(pred, esz, src.ctypes.data, n, comp.ctypes.data, D, 1))
        t1 = time.perf_counter()
        R.L.ref_decompress(pred, esz, comp.ctypes.data, dec.ctypes.data)
        t2 = time.perf_counter()
        best_e, best_d = min(best_e, t1 - t0), min(best_d, t2 - t1)
    raw = n * esz
    return {"threads": 1, "sample": f"first {raw >> 20} MiB of the stream as ONE reference call (no chunking), best of 2",
            "encode_gbs": raw / best_e / 1e9, "decode_gbs": raw / best_d / 1e9,
            "value": raw / (best_e + best_d) / 1e9, "ratio": raw / max(1, ret * esz),
            "roundtrip_ok": bool(np.array_equal(dec[: raw], src[: raw]))}


def parity_vs_reference(esz, refrun, payload, off, nb, first_chunk=0):
    """Compares chunk streams [first_chunk, first_chunk + refrun.nchunks) of the GPU frame (payload bytes + directory
    off/nb, numpy) with the reference's bytes for the same slices.  The reference returns floor(bytes/2) elements
    for 16-bit streams, so a GPU stream may be one byte longer than ret*esz; everything the reference
    counts must be identical."""
    import numpy as np
    comp, stride, rets = refrun["comp"], refrun["stride"], refrun["rets"]
    nchunks = refrun["nchunks"]
    equal, first_bad, gpu_bytes = 0, None, 0
    for c in range(nchunks):
        want_n = int(rets[c]) * esz
        a = int(off[first_chunk + c]); n = int(nb[first_chunk + c])
        gpu_bytes += n
        ok = (n == want_n or (esz == 2 and n == want_n + 1)) and \
            np.array_equal(payload[a: a + want_n], comp[c * stride: c * stride + want_n])
        if ok:
            equal += 1
        elif first_bad is None:
            first_bad = first_chunk + c
    return {"chunks_compared": nchunks, "chunks_equal": equal, "all_equal": equal == nchunks,
            "first_differing_chunk": first_bad, "gpu_stream_bytes": gpu_bytes, "reference_stream_bytes": refrun["comp_bytes"],
            "against": "oracle/_ref (unmodified reference) on the same chunk slices"}


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
        res = reference_chunks(cfg, xh, chunk_rows, ncores, repeats=1)
        if i >= warm:
            ts.append(res["t_enc"] + res["t_dec"])
    raw = res["raw"]
    total = sum(ts)
    val = raw * steps / total / 1e9
    line = {
        "impl": "reference", "metric": "encode+decode GB/s", "value": val, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": 1e3 * total / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8" if cfg["esz"] == 1 else "u16", "data": "synthetic",
        "config": config_record(cfg_name, cfg, raw, chunk_rows, max(1, args.gpus)),
        "ratio": res["ratio"],
        "cpu_baseline": {"value": val, "unit": "GB/s", "cores": ncores, "kind": "reference",
                         "sample": f"whole {raw >> 20} MiB workload per step, {res['nchunks']} chunk slices, "
                                   f"std::thread pool over oracle/_ref (unmodified reference, g++ -O3 -mavx2 -mbmi2)",
                         "encode_gbs": raw / res["t_enc"] / 1e9,
                         "decode_gbs": raw / res["t_dec"] / 1e9, "ratio": res["ratio"],
                         "roundtrip_ok": res["roundtrip_ok"]},
        "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line, args.real_stdout)
    return 0


# ------------------------------------------------------------------ the GPU arm

class Env:
    """Process-wide state of the GPU arm (one process per GPU)."""

    def __init__(self):
        import torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the Sprintz B200 path has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.comm = None

    def barrier(self):
        import torch
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, vals):
        import torch
        import torch.distributed as dist
        t = torch.tensor(list(vals), dtype=torch.float64, device=self.dev)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]


def roof_record(alg_bytes, t, kern, cfg_name, chunk_rows, full_size):
    peak, peak_src = peaks()
    return {"bound": "hbm", "achieved": alg_bytes / t / 1e9, "peak": peak, "unit": "GB/s",
            "frac": alg_bytes / t / 1e9 / peak,
            "traffic": ncu_traffic(cfg_name, kern, chunk_rows) if full_size else None,
            "alg_bytes_per_launch": alg_bytes, "ms_per_launch": 1e3 * t, "peak_source": peak_src,
            "formula": "N*(1+1/ratio) / kernel time / measured HBM copy bandwidth (SURVEY 8d)", "kernel": kern}


def timed_codec_loop(env, codec, x, out, steps):
    """K steps of encode + decode, CUDA events on the launching stream. -> per-rank times (s) and kernel splits."""
    import torch
    import sprintz_b200 as spz
    L = spz.lib()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * steps + 1)]
    L.spz_profile_enable(1)
    launches0 = int(L.spz_kernel_launches())
    env.barrier()
    ev[0].record()
    fr = None
    for i in range(steps):
        fr = codec.encode(x)
        ev[2 * i + 1].record()
        codec.decode(fr, out)
        ev[2 * i + 2].record()
    env.barrier()
    launches = int(L.spz_kernel_launches()) - launches0
    t_total = ev[0].elapsed_time(ev[2 * steps]) / 1e3
    t_enc = sum(ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(steps)) / 1e3
    t_dec = sum(ev[2 * i + 1].elapsed_time(ev[2 * i + 2]) for i in range(steps)) / 1e3
    ms3 = (C.c_float * 3)()
    ncalls = C.c_int(0)
    L.spz_profile_collect(ms3, C.byref(ncalls))
    L.spz_profile_enable(0)
    k = max(1, ncalls.value)
    return dict(fr=fr, t_total=t_total, t_enc=t_enc, t_dec=t_dec, launches=launches,
                t_enc_kernel=ms3[0] / 1e3 / k, t_scan=ms3[1] / 1e3 / k, t_gather=ms3[2] / 1e3 / k,
                t_dec_kernel=t_dec / steps)


def kernel_names():
    import sprintz_b200 as spz
    L = spz.lib()
    fam = ("spz_%s_kernel", "spz_%s_hd_kernel", "spz_%s_lane_kernel", "spz_%s_wide_kernel")
    return fam[L.spz_last_kernel_family(0)] % "encode", fam[L.spz_last_kernel_family(1)] % "decode"


def frame_to_host(fr, hdr):
    """-> (payload bytes, chunk_off, chunk_nbytes) as numpy, for the parity check."""
    off, nb, po = fr.directory()
    payload = fr.buf[po: po + int(hdr.payload_bytes)].cpu().numpy()
    return payload, off, nb


def measure_config(env, args, cfg_name, cfg, steps, warm, main):
    """Device-timed encode+decode of one configuration on this rank's shard. `main` adds clocks, query, e2e."""
    import numpy as np
    import torch
    import sprintz_b200 as spz
    L = spz.lib()
    esz, D, pred = cfg["esz"], cfg["D"], cfg["pred"]
    chunk_rows = (args.chunk_rows if main and args.chunk_rows else 0) or cfg["chunk_rows"]
    world, rank, dev = env.world, env.rank, env.dev
    # N > 1: every rank's shard is GATHER_ROUNDS whole stripes of whole chunks (the gathered frame needs whole chunks
    # everywhere but at the end of the global stream)
    x, rows = make_stream(cfg, dev, seed=12345 + rank, multiple=16 if world == 1 else chunk_rows * max(GATHER_ROUNDS, P2P_ROUNDS))
    n = x.numel()
    raw = n * esz
    codec = spz.FrameCodec(pred, esz, n, D, chunk_rows, device=dev)
    out = torch.empty_like(x)
    full_size = not args.bytes_mb

    # ---- correctness of exactly what is timed: the round trip here, byte parity with the reference below
    fr = codec.encode(x)
    codec.decode(fr, out)
    torch.cuda.synchronize()
    codec.check_status()
    if not torch.equal(out, x):
        raise SystemExit(f"bench.py: {cfg_name}: decode(encode(x)) != x")
    hdr = fr.header()
    payload = int(hdr.payload_bytes)

    # clocks are sampled from the start of warm-up to the end of the timed region (the timed region alone can
    # be shorter than one nvidia-smi period); warm-up is extended until a few samples under load exist
    sampler = None
    if main:
        sampler = ClockSampler(env.local_rank)
        sampler.start()
    for _ in range(warm):
        codec.decode(codec.encode(x), out)
    if main:
        t_w = time.time()
        while len(sampler.lines) < 3 and time.time() - t_w < 4.0:
            codec.decode(codec.encode(x), out)
            torch.cuda.synchronize()
    r = timed_codec_loop(env, codec, x, out, steps)
    clocks = sampler.stop() if sampler else None
    fr = r["fr"]
    k_enc, k_dec = kernel_names()
    t_total, t_enc_m, t_dec_m = env.max_over_ranks([r["t_total"], r["t_enc"], r["t_dec"]])

    # exact sizes from the frame's own directory: ratio_exact = raw / sum(chunk_nbytes) is the reference's ratio
    # (payload_bytes also counts the 16-byte alignment padding between chunk streams)
    sizes = fr.buf[64 + 8 * (codec.nchunks + 1): 64 + 8 * (codec.nchunks + 1) + 4 * codec.nchunks].view(torch.int32)
    stream_bytes = int(sizes.to(torch.int64).sum().item())
    ratio = raw / payload
    alg_bytes = raw * (1.0 + 1.0 / ratio)               # N*(1 + 1/r), both directions
    peak, _ = peaks()
    rec = {
        "config": config_record(cfg_name, cfg, raw, chunk_rows, world),
        "value": world * raw * steps / t_total / 1e9, "unit": "GB/s", "ms_per_step": 1e3 * t_total / steps,
        "encode_gbs": world * raw * steps / t_enc_m / 1e9, "decode_gbs": world * raw * steps / t_dec_m / 1e9,
        "ratio": ratio, "ratio_exact": raw / stream_bytes, "stream_bytes": stream_bytes, "payload_bytes": payload,
        "roofline_decode": roof_record(alg_bytes, r["t_dec_kernel"], k_dec, cfg_name, chunk_rows, full_size),
        "roofline_encode": roof_record(alg_bytes, r["t_enc_kernel"], k_enc, cfg_name, chunk_rows, full_size),
        # the whole frame encode (codec kernel + directory scan + compaction) against the same algorithmic bytes
        "roofline_encode_whole": {"frac": alg_bytes / (r["t_enc"] / steps) / 1e9 / peak, "ms": 1e3 * r["t_enc"] / steps},
        # the stricter read-only form the north star names: HBM bytes READ per launch / kernel time / peak
        "hbm_read_roofline": {"decode_frac": raw / ratio / r["t_dec_kernel"] / 1e9 / peak,
                              "encode_frac": raw / r["t_enc_kernel"] / 1e9 / peak,
                              "decode_raw_gbs_kernel": raw / r["t_dec_kernel"] / 1e9,
                              "encode_raw_gbs_kernel": raw / r["t_enc_kernel"] / 1e9},
        "encode_breakdown_ms": {"codec_kernel": 1e3 * r["t_enc_kernel"], "directory_scan": 1e3 * r["t_scan"],
                                "compaction": 1e3 * r["t_gather"], "slabs": int(L.spz_last_encode_slabs())},
        "gpu_launches": r["launches"],
    }
    dom_is_dec = r["t_dec_kernel"] >= r["t_enc_kernel"]
    rec["roofline"] = rec["roofline_decode"] if dom_is_dec else rec["roofline_encode"]
    if clocks is not None:
        rec["clocks"] = clocks

    # ---- the slab-pipelined frame encode (three slabs, thin compaction CTAs beside the next slab's codec kernel): it pays for
    # the wide-row encoder (many machine rounds of small CTAs), not for the lane encoders -- reported beside the classic time
    old_slabs = spz.set_encode_pipeline(3)
    try:
        for _ in range(2):
            codec.encode(x)
        pe = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        env.barrier()
        pe[0].record()
        for _ in range(5):
            fr_p = codec.encode(x)
        pe[1].record()
        env.barrier()
        codec.decode(fr_p, out)
        torch.cuda.synchronize()
        t_pipe = env.max_over_ranks([pe[0].elapsed_time(pe[1]) / 5e3])[0]
        rec["encode_pipelined"] = {"slabs": int(L.spz_last_encode_slabs()), "ms": 1e3 * t_pipe, "encode_gbs": world * raw / t_pipe / 1e9,
                                   "classic_ms": 1e3 * r["t_enc"] / steps, "roundtrip_ok": bool(torch.equal(out, x))}
    except Exception as e:  # noqa: BLE001
        rec["encode_pipelined"] = {"error": f"{type(e).__name__}: {e}"}
    finally:
        spz.set_encode_pipeline(old_slabs)
        fr = codec.encode(x)                                # the classic frame again (what the checks below read)
        torch.cuda.synchronize()

    # ---- the frame on the host, for the parity check (before the buffers go away)
    want_parity = rank == 0 and not args.no_cpu_baseline
    host_frame = frame_to_host(fr, hdr) if want_parity else None

    # ---- query-on-compressed, outside the timed region
    if main:
        rec["query"] = measure_query(env, codec, fr, x, out, cfg, chunk_rows, payload)

    # ---- entropy stage over the frame payload (BASELINE config 3's "Huffman header"), outside the timed region
    if main:
        rec["entropy_stage"] = measure_entropy(env, fr, hdr, raw, payload)

    # ---- exchange step: gather every rank's frame on rank 0 over NCCL, inside the timed region
    if world > 1 and env.comm is not None:
        rec.update(measure_gather(env, args, codec, fr, x, out, cfg, chunk_rows, raw, payload, r, steps))

    # ---- end to end through the host-buffer C ABI (pinned host memory), every step H2D + D2H
    xh = None
    if main:
        e2e_steps = max(1, min(steps, args.e2e_steps))
        xh = torch.empty(n, dtype=x.dtype, pin_memory=True)
        xh.copy_(x)
        fcap = int(L.spz_frame_max_bytes(esz, n, D, chunk_rows))
        fh = torch.empty(fcap, dtype=torch.uint8, pin_memory=True)
        oh = torch.empty(n, dtype=x.dtype, pin_memory=True)
        del codec, out, fr
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
        env.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            fb = e2e_step()
        torch.cuda.synchronize()
        t_e2e = env.max_over_ranks([time.perf_counter() - t0])[0]
        L.spz_release_workspaces()
        rec["e2e"] = {"value": world * raw * e2e_steps / t_e2e / 1e9, "unit": "GB/s",
                      "h2d_bytes_per_step": raw + int(fb), "d2h_bytes_per_step": int(fb) + raw, "steps": e2e_steps,
                      "api": "spz_frame_compress_host + spz_frame_decompress_host (pinned host buffers)"}
        del fh, oh
    else:
        del codec, out, fr

    # ---- the reference on the host cores beside it + byte parity of the whole frame (rank 0)
    if want_parity:
        try:
            ncores = os.cpu_count() or 1
            xs = (xh.numpy() if xh is not None else x.cpu().numpy()).view(np.uint8 if esz == 1 else np.uint16)
            clen = chunk_rows * D
            # N = 1, main config: the WHOLE stream; otherwise a bounded sample of leading chunks
            if main and world == 1 and not args.cpu_sample_mb:
                sample_elems = n
            else:
                mb = args.cpu_sample_mb or 256
                if not (main and world == 1):
                    mb = min(mb, 256)
                sample_elems = min(n, max(clen, (mb << 20) // esz // clen * clen))
            rr = reference_chunks(cfg, xs[:sample_elems], chunk_rows, ncores, repeats=1, keep=True)
            pay_h, off_h, nb_h = host_frame
            rec["parity"] = parity_vs_reference(esz, rr, pay_h, off_h, nb_h)
            rec["parity"]["whole_stream"] = sample_elems == n
            if sample_elems == n:
                rec["parity"]["ratio_reference"] = rr["ratio"]
                rec["parity"]["ratio_equal"] = rec["parity"]["gpu_stream_bytes"] == rr["comp_bytes"]
            cb = {"value": rr["raw"] / (rr["t_enc"] + rr["t_dec"]) / 1e9, "unit": "GB/s", "cores": ncores,
                  "kind": "reference",
                  "sample": f"first {rr['raw'] >> 20} MiB of the same stream, same {chunk_rows}-row chunk slices, "
                            f"oracle/_ref (unmodified reference) on {ncores} threads",
                  "encode_gbs": rr["raw"] / rr["t_enc"] / 1e9, "decode_gbs": rr["raw"] / rr["t_dec"] / 1e9,
                  "ratio": rr["ratio"], "roundtrip_ok": rr["roundtrip_ok"]}
            del rr
            if main:
                cb["single_thread_whole_stream"] = reference_single_stream(cfg, xs)
            rec["cpu_baseline"] = cb
        except Exception as e:  # noqa: BLE001
            rec["cpu_baseline"] = {"value": None, "unit": "GB/s", "cores": 0, "kind": "reference",
                                   "sample": f"unavailable: {type(e).__name__}: {e}"}
    del x, xh, host_frame
    torch.cuda.empty_cache()
    return rec


def measure_query(env, codec, fr, x, out, cfg, chunk_rows, payload):
    """per-dim REDUCE_SUM / REDUCE_MAX over the frame without materialising it (fused into the lane decoder where
    the shape allows); reads N/r bytes only"""
    import torch
    import sprintz_b200 as spz
    L = spz.lib()
    esz, D, pred = cfg["esz"], cfg["D"], cfg["pred"]
    n = x.numel()
    raw = n * esz
    try:
        qres = torch.empty(D, dtype=torch.int64, device=env.dev)
        qwork = out.view(torch.uint8)               # decode+reduce shapes decode into this (>= n * esz bytes)
        info = {}
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
            fused = bool(L.spz_last_query_fused())
            ok = None
            if n % D == 0:
                want = torch.zeros(D, dtype=torch.int64, device=env.dev)
                rows_per = max(1, (1 << 26) // D)
                xv = x.view(-1, D)
                for a in range(0, xv.shape[0], rows_per):       # slab by slab: no 4x-sized temporaries
                    s = xv[a: a + rows_per].to(torch.int32) & ((1 << (8 * esz)) - 1)
                    if op == 2:
                        want += s.sum(0, dtype=torch.int64)
                    else:
                        want = torch.maximum(want, s.max(0).values.to(torch.int64))
                ok = bool(torch.equal(qres, want))
            info[opname] = {"ms": 1e3 * tq, "fused": fused, "raw_gbs": raw / tq / 1e9,
                            "hbm_bytes_read": payload, "hbm_read_frac": payload / tq / 1e9 / peaks()[0],
                            "matches_reduction_of_input": ok}
        return info
    except Exception as e:  # noqa: BLE001
        return {"error": f"{type(e).__name__}: {e}"}


def measure_entropy(env, fr, hdr, raw, payload):
    """Byte-wise canonical Huffman over the frame's payload (spz_huff_*_dev).  PARITY UNPINNED: Huff0 is not in the reference
    tree (cpp/Compress/entropy.cpp is a stub); the stage is checked against its own CPU specification in tests/."""
    import torch
    import sprintz_b200 as spz
    try:
        po = int(hdr.payload_off)
        pay = fr.buf[po: po + payload]
        h = spz.HuffCodec(payload, device=env.dev)
        back = torch.empty(payload, dtype=torch.uint8, device=env.dev)
        for _ in range(2):
            cont = h.encode(pay)
            h.decode(cont, back)
        torch.cuda.synchronize()
        h.check_status()
        ok = bool(torch.equal(back, pay))
        size = h.size()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        reps = 3
        ev[0].record()
        for _ in range(reps):
            h.encode(pay)
        ev[1].record()
        for _ in range(reps):
            h.decode(cont, back)
        ev[2].record()
        torch.cuda.synchronize()
        te, td = ev[0].elapsed_time(ev[1]) / reps / 1e3, ev[1].elapsed_time(ev[2]) / reps / 1e3
        del h, back
        return {"what": "byte-wise canonical Huffman over the frame payload (4 KB segments, 11-bit limit); parity unpinned",
                "payload_bytes": payload, "container_bytes": size, "stage_ratio": payload / size,
                "ratio_with_entropy_stage": raw / (po + size), "encode_ms": 1e3 * te, "decode_ms": 1e3 * td,
                "encode_gbs_of_payload": payload / te / 1e9, "decode_gbs_of_payload": payload / td / 1e9, "roundtrip_ok": ok}
    except Exception as e:  # noqa: BLE001
        return {"error": f"{type(e).__name__}: {e}"}


def measure_gather(env, args, codec, fr, x, out, cfg, chunk_rows, raw, payload, r, steps):
    """Encode + NCCL gather of every rank's frame into ONE decodable frame on rank 0 (spz_frame_gather_dev), with the
    collective inside the timed region; rank 0 decodes the gathered frame and checks it against the inputs."""
    import torch
    import torch.distributed as dist
    import sprintz_b200 as spz
    world, rank, dev = env.world, env.rank, env.dev
    try:
        g = spz.FrameGather(env.comm, codec, dst=0, stripe_chunks=codec.nchunks // GATHER_ROUNDS)
        gf = g.gather(fr)                                   # warm-up + correctness
        torch.cuda.synchronize()
        env.barrier()
        # rank 0 decodes the gathered frame: its own shard element by element, every other rank's through checksums
        ok = g.verify_against_shards(gf if rank == 0 else None, x, all_ranks=True)
        env.barrier()
        reps = max(3, min(steps, 10))
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        for _ in range(reps):
            g.gather(fr)
        ev[1].record()
        for _ in range(reps):                               # encode + gather, streamed (send of range i overlaps encode of i+1)
            g.encode_and_gather(x)
        ev[2].record()
        env.barrier()
        t_g = ev[0].elapsed_time(ev[1]) / reps / 1e3
        t_eg = ev[1].elapsed_time(ev[2]) / reps / 1e3
        t_g, t_eg = env.max_over_ranks([t_g, t_eg])
        ok2 = g.verify_against_shards(g.last_frame if rank == 0 else None, x, all_ranks=True)
        moved = (world - 1) * payload                       # bytes that crossed NVLink into rank 0 (payloads are ~equal)
        t_enc_step = r["t_enc"] / steps
        t_dec_step = r["t_dec"] / steps
        rec = {"gather": {"ms": 1e3 * t_g, "payload_bytes_total": world * payload, "gbs_into_rank0": moved / t_g / 1e9,
                          "backend": "ncclSend/ncclRecv grouped, libsprintz_b200.so (spz_frame_gather_dev)",
                          "encode_plus_gather_ms": 1e3 * t_eg, "encode_ms": 1e3 * t_enc_step,
                          "overlap_efficiency": max(t_enc_step, t_g) / t_eg,
                          "gathered_frame_decodes_to_input": ok, "streamed_frame_decodes_to_input": ok2},
               "value_with_gather": world * raw / (t_eg + t_dec_step) / 1e9,
               "encode_with_gather_gbs": world * raw / t_eg / 1e9}
        # ---- the same exchange over peer memory: every stripe's compaction kernel stores straight into rank 0's frame
        try:
            del g, gf                                       # the NCCL destination buffer (a whole gathered frame) goes first
            torch.cuda.empty_cache()                        # the shared regions are cudaMalloc'ed by the library, outside torch's pool
            g = spz.FrameGather(env.comm, codec, dst=0, stripe_chunks=codec.nchunks // P2P_ROUNDS, alloc=False)
            g.enable_p2p()
            g.encode_and_gather_p2p(x)                      # warm-up + correctness
            torch.cuda.synchronize()
            env.barrier()
            ok3 = g.verify_against_shards(g.last_frame if rank == 0 else None, x, all_ranks=True)
            env.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                g.encode_and_gather_p2p(x)
            e1.record()
            env.barrier()
            t_p2p = env.max_over_ranks([e0.elapsed_time(e1) / reps / 1e3])[0]
            ok4 = g.verify_against_shards(g.last_frame if rank == 0 else None, x, all_ranks=True)
            # what bounds it: rank 0 ingress (the other ranks' payload over NVLink) or the encode itself
            rec["gather_p2p"] = {"encode_plus_gather_ms": 1e3 * t_p2p, "gbs_into_rank0": moved / t_p2p / 1e9,
                                 "backend": "fused compaction + NVLink peer stores (spz_frame_encode_gather_p2p_dev), no NCCL in the data path",
                                 "overlap_efficiency": max(t_enc_step, t_g) / t_p2p, "stripes_per_rank": P2P_ROUNDS,
                                 "frame_decodes_to_input": bool(ok3) and bool(ok4) if rank == 0 else None}
            if t_p2p < t_eg:
                rec["value_with_gather"] = world * raw / (t_p2p + t_dec_step) / 1e9
                rec["encode_with_gather_gbs"] = world * raw / t_p2p / 1e9
                rec["value_with_gather_backend"] = "p2p"
            else:
                rec["value_with_gather_backend"] = "nccl"
        except Exception as e:  # noqa: BLE001
            rec["gather_p2p"] = {"error": f"{type(e).__name__}: {e}"}
        return rec
    except Exception as e:  # noqa: BLE001
        return {"gather": {"error": f"{type(e).__name__}: {e}"}}


def run_ours(args, cfg, cfg_name):
    import torch
    import torch.distributed as dist
    import sprintz_b200 as spz

    env = Env()
    pin_to_gpu_numa_node(env.local_rank)
    if env.world > 1:
        dist.init_process_group("nccl", device_id=env.dev)
        try:
            env.comm = spz.Communicator.from_torch(env.rank, env.world)
        except Exception as e:  # noqa: BLE001
            print(f"bench.py: NCCL communicator unavailable: {e}", file=sys.stderr)
    steps, warm = args.steps, args.warmup
    rec = measure_config(env, args, cfg_name, cfg, steps, warm, main=True)
    dtype = "u8" if cfg["esz"] == 1 else "u16"
    line = {"metric": "encode+decode GB/s", "value": rec.pop("value"), "unit": rec.pop("unit"), "n_gpus": env.world,
            "steps": steps, "warmup": warm, "ms_per_step": rec.pop("ms_per_step"), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic"}
    line.update(rec)
    # ---- the other BASELINE configs as sub-records (shorter runs, same measurements)
    if not args.no_sub and not args.bytes_mb:
        subs = {}
        for name in SUB_CONFIGS:
            if name == cfg_name:
                continue
            try:
                subs[name] = measure_config(env, args, name, dict(CONFIGS[name]), max(3, min(steps, 10)), 3, main=False)
            except SystemExit:
                raise
            except Exception as e:  # noqa: BLE001
                subs[name] = {"error": f"{type(e).__name__}: {e}"}
        line["sub_records"] = subs
    if env.rank == 0:
        emit(line, args.real_stdout)
    if env.comm is not None:
        env.comm.close()
    if env.world > 1:
        dist.destroy_process_group()
    return 0


def pin_to_gpu_numa_node(local_rank):
    """Host buffers of the e2e path are first-touched by this process: keep it on the CPUs of its GPU's NUMA node."""
    try:
        import torch
        pci = torch.cuda.get_device_properties(local_rank).pci_bus_id if hasattr(torch.cuda.get_device_properties(local_rank), "pci_bus_id") else None
        if pci is None:
            out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local_rank)],
                                 capture_output=True, text=True, timeout=10).stdout.strip()
            pci = out.lower()[-12:] if out else None
        if not pci:
            return
        node_f = f"/sys/bus/pci/devices/{pci.lower() if len(pci) == 12 else '0000:' + pci.lower()[-7:]}/numa_node"
        if not os.path.exists(node_f):
            return
        node = int(open(node_f).read().strip())
        if node < 0:
            return
        cpus_f = f"/sys/devices/system/node/node{node}/cpulist"
        cpus = set()
        for part in open(cpus_f).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


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
    ap.add_argument("--config", default=MAIN_CONFIG, choices=sorted(CONFIGS))
    ap.add_argument("--chunk-rows", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-sample-mb", type=int, default=0,
                    help="bound the reference / parity run to the first N MiB (0 = the whole stream at N = 1)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="also skips the byte-parity check (it needs the reference bytes)")
    ap.add_argument("--no-sub", action="store_true", help="skip the c2 / c4 sub-records")
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
