"""-m gpu, needs >= 2 GPUs (skipped otherwise): the data-parallel training step reproduces the single-process step on
the whole batch (tools/ddp_check.py under torchrun, one rank per GPU) -- in both forms: the default peer-memory
optimizer step (csrc/peer_optim.cu) and the NCCL all-reduce schedule (MAGB200_PEER_ADAMW=0)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("peer", ["1", "0"])
def test_two_rank_training_equals_single_process(peer):
    n = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr",
           "127.0.0.1", "--master-port", "29611" if peer == "1" else "29612", os.path.join(ROOT, "tools", "ddp_check.py")]
    env = dict(os.environ, MAGB200_PEER_ADAMW=peer, DDP_STEPS="3" if peer == "1" else "1")
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0 and "DDP_CHECK PASS" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
    assert ("peer-memory optimizer: True" in out.stdout) == (peer == "1"), out.stdout[-2000:]
