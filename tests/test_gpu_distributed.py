"""Multi-GPU path on real GPUs (needs >= 2 B200s: `gpurun --gpus 2`): torchrun + NCCL + the C-ABI kernels, checked against
the oracle over the union of the shards. Skipped on a single-GPU box."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_scan_join_aggregate():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    command = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
               "127.0.0.1", "--master-port", "29631", os.path.join(REPO, "tests", "gpu_distributed_worker.py")]
    completed = subprocess.run(command, capture_output=True, text=True, timeout=900)
    assert completed.returncode == 0, completed.stdout[-3000:] + completed.stderr[-3000:]
    assert "distributed OK on 2 GPUs" in completed.stdout
