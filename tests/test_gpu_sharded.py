"""Batch sharding on real GPUs: a torchrun world of 2 (skipped on a box with one GPU) drives
tools/shard_parity_worker.py, which checks the sharded cycle — exchanges over peer memory and over NCCL — against
the unsharded handle and the CPU oracle at 2^20 x 56 and at a batch that leaves the last block of a shard ragged."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("batch", [1 << 20, 40000])
def test_sharded_world2_matches_unsharded_and_oracle(batch):
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs (run with `gpurun --gpus 2`); profiles/ holds the last 2-GPU run")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "shard_parity_worker.py"), str(batch), "3"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "SHARD_PARITY_OK" in out.stdout, out.stdout[-3000:]
