"""Drop-in proof through the reference trainer's own step (DDP(find_unused_parameters=True) + autocast + GradScaler +
torch AdamW + DataLoader), restated in scripts/dropin_train_check.py because /root/reference does not travel to the GPU box."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(world, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "scripts", "dropin_train_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "reference train() step on mammoclip_b200 factories ok" in r.stdout


def test_reference_train_step_single_rank_ddp():
    _run(1, 29881)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_reference_train_step_two_ranks_ddp():
    _run(2, 29882)
