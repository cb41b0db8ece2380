"""Multi-GPU check of the face/tracer-sharded path (NCCL strip exchange): runs tests/mgpu_face_sharding_check.py under
torchrun on 2 GPUs when the box has them; skipped on a single-GPU box (the exchange schedule itself is covered on CPU by
tests/test_partition_gloo.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_face_sharded_step_matches_single_context():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", os.path.join(HERE, "mgpu_face_sharding_check.py")], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0 and "ok=True" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_submosaic_step_matches_single_context():
    """24 sub-domains (layout 2 x 2 per tile) over 2 GPUs, sub-tile contexts + gather-list halos over NCCL."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29519", os.path.join(HERE, "mgpu_submosaic_check.py")], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0 and "ok=True" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
