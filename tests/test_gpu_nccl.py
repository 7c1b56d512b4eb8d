"""The product's multi-GPU API under NCCL on real hardware: DistributedRaycaster / SyncGradients /
FusedVolumeSGD with two ranks must reproduce the single-process gradients of the whole view batch.  Skipped on a one-GPU box."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_distributed_raycaster_matches_single_process_under_nccl():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(ROOT, "tools", "dist_check.py")], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("rel-L2 vs single process") == 2 and r.stdout.count("fused distributed volume step") == 2
