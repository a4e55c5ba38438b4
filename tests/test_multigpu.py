"""Multi-GPU parity (NCCL, one rank per GPU) -- runs tests/mgpu_worker.py under torchrun on every
grid shape the visible GPUs allow (1x2, 2x1, 2x2, 2x4).  Skipped when fewer than 2 GPUs are visible."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world,height", [(2, 1), (2, 2), (4, 2), (8, 2)])
def test_el_on_grid(world, height):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world + height),
           os.path.join(ROOT, "tests", "mgpu_worker.py"), str(height)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert p.returncode == 0 and "MGPU OK" in p.stdout, p.stdout[-4000:] + p.stderr[-4000:]
