"""GPU-side multi-rank parity (needs >= 2 GPUs on the box): torchrun launches tests/mgpu_worker.py with one process per
GPU; the ranks' frames are gathered on rank 0 through the library's NCCL path and the gathered frame must decode to the
global stream with every chunk byte-identical to the reference."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4])
def test_gathered_frame_is_the_global_stream(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, this box has {torch.cuda.device_count()}")
    env = dict(os.environ, NCCL_DEBUG="WARN")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(29500 + world),
                        os.path.join(ROOT, "tests", "mgpu_worker.py")], capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert r.returncode == 0 and "MGPU OK" in r.stdout, (r.stdout[-1500:], r.stderr[-3000:])
    assert "unbatched P2P" not in r.stderr
