"""Sharded featurise path on real GPUs (needs >= 2 devices; skipped on the 1-GPU test box).
The host-side logic of the N > 1 path is covered on CPU by tests/test_dist_gloo.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_featurise_equals_single_gpu():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "scripts", "check_dist_parity.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "DIST PARITY PASS" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
