"""bench.py's JSON contract, on the CPU: the reference arm (`--impl reference`) runs the unmodified reference on host
threads and must print ONE JSON line with the agreed keys; the GPU arm must refuse to run without a CUDA device
(no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, timeout=300):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                          timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_one_contract_line(ref):
    r = _run(["--impl", "reference", "--bytes-mb", "8", "--steps", "2", "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "encode+decode GB/s" and d["unit"] == "GB/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    assert d["dtype"] == "u16" and d["data"] == "synthetic" and d["config"]["workload"].startswith("c3")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert cb["roundtrip_ok"] is True
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == "GB/s" and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_gpu_arm_refuses_without_cuda():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a CUDA device is present")
    except ImportError:
        pytest.skip("torch not importable")
    r = _run(["--bytes-mb", "8", "--steps", "1", "--warmup", "3"])
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)
    assert not [ln for ln in r.stdout.splitlines() if ln.strip().startswith("{")]
