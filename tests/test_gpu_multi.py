"""Multi-GPU (single node) check of the fused NVLink gather + InfoNCE kernel; needs >= 2 visible B200s."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("lse_exchange", ["0", "1"])        # 1: every rank scores only its own rows/columns and the LSEs are exchanged (ABI 5)
def test_fused_p2p_loss_matches_nccl_path(lse_exchange):
    n = min(torch.cuda.device_count(), 8)
    n = 8 if n >= 8 else 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29917", os.path.join(ROOT, "scripts", "multi_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, MCLIP_LOSS_LSE=lse_exchange, MCLIP_CHECK_NO_SWEEP="1"))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "fused P2P loss == all_gather/reduce_scatter oracle" in r.stdout
