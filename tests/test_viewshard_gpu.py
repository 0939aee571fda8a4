"""Camera-view sharding with neighbour K/V exchange on real GPUs (needs >= 2 devices; skipped otherwise)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (run under gpurun --gpus 2)")
def test_view_sharded_step_matches_single_gpu():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "run_viewshard.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=420)
    assert "VIEWSHARD OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
