"""On-hardware correctness of the data-parallel path: a 2-rank torchrun job over NCCL.
Skipped on a single-GPU box; run it with `gpurun --gpus 2 -- python -m pytest tests/test_ddp_nccl_gpu.py -m gpu`."""
import socket
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_two_rank_nccl_gradients_equal_mean_of_single_gpu_runs():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(ROOT / "tests" / "ddp_nccl_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=str(ROOT))
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-4000:]
    assert "DDP_NCCL_OK 0" in out and "DDP_NCCL_OK 1" in out, out[-4000:]
    print("\n".join(line for line in out.splitlines() if "DDP_NCCL_OK" in line))
