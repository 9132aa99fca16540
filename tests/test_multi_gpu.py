"""Velocity-space sharding over >= 2 GPUs (needs them: run with `gpurun --gpus 2`)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("mode", ["nccl", "callback"])
def test_two_rank_parity(mode, oracle_lib):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29617", os.path.join(HERE, "mgpu_worker.py"), mode]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert "MGPU_RESULT PASS" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
