"""Real multi-GPU parity (CUDA IPC windows, NVLink stores, NCCL fallback): spawns tools/dist_check.py under torchrun on 2
GPUs when the box has them (the single-GPU driver box skips; bench.py --gpus N carries the same bit-identity flags)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_peer_build_is_bit_identical():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "dist_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "FAIL" not in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("PASS") >= 6
