"""Sharded (one process per GPU) parity: spawns tests/multi_gpu_worker.py under torchrun.
Skipped unless at least two GPUs are visible."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_parity(world):
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world), os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    sys.stdout.write(proc.stdout[-6000:])
    sys.stderr.write(proc.stderr[-3000:])
    assert proc.returncode == 0 and "MULTI-GPU PARITY GREEN" in proc.stdout
