"""Two-GPU SPMD run (torchrun, NCCL over NVLink): row-sharded fused kernel operator and replicated-state
PCG / ASkotch solves.  Skipped on boxes with a single GPU."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_two_gpu_sharded_operator_and_solvers():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(HERE, "spmd_gpu_worker.py")]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-2000:]
    assert "SPMD_GPU_OK" in proc.stdout, proc.stdout[-2000:] + proc.stderr[-2000:]
