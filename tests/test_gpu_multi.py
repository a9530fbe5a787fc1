"""Multi-GPU parity (needs >= 2 GPUs; skipped on a one-GPU box): sharded scoring with halo
exchange + the core-set rounds' candidate all-gather must reproduce the single-GPU pick list."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_two_gpus_match_one(built_lib):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tools", "multi_gpu_check.py")]
    # prune even these small shards so that the segment / tile bookkeeping is exercised with row offsets
    env = dict(os.environ, VATLQ_PRUNE_MIN_ROWS="0")
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0 and "MULTI_GPU_CHECK PASS" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
