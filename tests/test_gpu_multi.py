"""Multi-GPU parity (BASELINE.json configs[4]): one process per GPU over NCCL,
spawned with torch.distributed.run exactly as the driver launches bench.py.
Skips cleanly on a 1-GPU box; the host-side logic is covered on CPU by
tests/test_sharded_gloo.py."""
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_gemv_and_in_kernel_allreduce(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, box has {torch.cuda.device_count()}")
    port = 29600 + os.getpid() % 1500 + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port", str(port),
           str(ROOT / "tests" / "multi_gpu_worker.py")]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-3000:]
    assert "multi-GPU ok" in proc.stdout
