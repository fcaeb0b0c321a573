"""2-GPU data-parallel test (skipped with fewer than 2 GPUs): replicas bitwise identical after DP steps and equal to a
single-GPU run on the merged batch."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_data_parallel_matches_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "dp_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, "\n".join(l for l in (r.stdout + r.stderr).splitlines() if "DP " in l or "Error" in l)[-3000:]
    assert "replicas bitwise equal=True" in r.stdout
