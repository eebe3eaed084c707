"""BASELINE config 5: frame encode / decode throughput over block size x ndims x element width x predictor, on 1 GPU or --
under torchrun -- on N GPUs (weak scaling: every rank encodes / decodes its own 256 MiB shard of synthetic AR(1), no
data-path collective; the time of a point is the max over ranks, the GB/s are whole-job aggregates).
Prints one JSON line per point (kernel family, GB/s of raw data, roofline fraction N*(1+1/r)/t/peak per GPU).

block_rows 8 is the reference's format (bit-exact with the reference); 16 / 32 / 64 are the block-size axis: the oracle's
generalisation, PARITY UNPINNED (no reference build exists), generic kernels.

usage: [torchrun --nproc-per-node N] python tools/dims_sweep.py [--mb 256] [--dims 8,16] [--blocks 8,16,32,64] [--force 0|1|2] [--chunk-rows 1024]
       > profiles/<round>_dims_sweep[_nN].jsonl"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
import sprintz_b200 as spz  # noqa: E402


def arg(name, default):
    return sys.argv[sys.argv.index(name) + 1] if name in sys.argv else default


mb = int(arg("--mb", 256))
dims = [int(v) for v in arg("--dims", "1,2,3,4,5,8,16,24,32,48,64,80,128,256").split(",")]
blocks = [int(v) for v in arg("--blocks", "8").split(",")]
force = int(arg("--force", 0))          # spz_set_kernel_path
chunk_rows = int(arg("--chunk-rows", 1024))
env = bench.Env()
if env.world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=env.dev)
peak, _ = bench.peaks()
L = spz.lib()
L.spz_set_kernel_path(force)
FAM = ("generic", "team(hd)", "lane", "wide")
dev = env.dev
for R in blocks:
    for esz in (1, 2):
        for D in dims:
            if R > 8 and 2 * R * D + D > 0xffff:
                continue                                    # the tail length is stored as a u16 (format.h:35-45)
            cfg = dict(esz=esz, D=D, nbytes=mb << 20)
            x, rows = bench.make_stream(cfg, dev, seed=12345 + env.rank, multiple=1024)
            n = x.numel()
            raw = n * esz
            out = torch.empty_like(x)
            for pred, pname in ((0, "delta"), (1, "xff")):
                codec = spz.FrameCodec(pred, esz, n, D, chunk_rows, device=dev, block_rows=R)
                for _ in range(2):
                    fr = codec.encode(x); codec.decode(fr, out)
                torch.cuda.synchronize()
                ok = bool(torch.equal(out, x))
                payload = int(fr.header().payload_bytes)
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                reps = 5 if R == 8 else 3
                env.barrier()
                ev[0].record()
                for _ in range(reps):
                    fr = codec.encode(x)
                ev[1].record()
                for _ in range(reps):
                    codec.decode(fr, out)
                ev[2].record()
                env.barrier()
                te, td = ev[0].elapsed_time(ev[1]) / reps / 1e3, ev[1].elapsed_time(ev[2]) / reps / 1e3
                te, td, nok = env.max_over_ranks([te, td, 0.0 if ok else 1.0])
                alg = raw + payload
                if env.rank == 0:
                    print(json.dumps({"block_rows": R, "elem_bits": 8 * esz, "ndims": D, "predictor": pname, "n_gpus": env.world,
                                      "raw_mib_per_gpu": raw >> 20, "nchunks_per_gpu": codec.nchunks, "chunk_rows": chunk_rows, "kernel_path": force,
                                      "ratio": round(raw / payload, 4), "roundtrip_ok": nok == 0.0,
                                      "parity": "bit-exact with the reference" if R == 8 else "unpinned (oracle generalisation)",
                                      "encode_kernels": FAM[L.spz_last_kernel_family(0)], "decode_kernels": FAM[L.spz_last_kernel_family(1)],
                                      "encode_gbs": round(env.world * raw / te / 1e9, 1), "decode_gbs": round(env.world * raw / td / 1e9, 1),
                                      "encode_frac": round(alg / te / 1e9 / peak, 3), "decode_frac": round(alg / td / 1e9 / peak, 3)}),
                          flush=True)
                del codec, fr
            del x, out
            torch.cuda.empty_cache()
if env.world > 1:
    dist.barrier()
    dist.destroy_process_group()
