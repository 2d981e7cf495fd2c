"""-m gpu: multi-GPU parity (needs >= 2 visible GPUs; skipped on a 1-GPU box).  Launches tests/multigpu_check.py with torchrun."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_parity():
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", os.path.join(ROOT, "tests", "multigpu_check.py")], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "MULTIGPU OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
    for tag in ("ppo update parity OK (fused LL peer exchange)", "sac update parity OK", "td3 update parity OK",
                "lagrange ppo update parity OK (nccl all-reduce)", "lagrange ppo update parity OK (one-shot peer all-reduce)"):
        assert tag in out.stdout, f"missing: {tag}\n" + out.stdout[-2000:]
