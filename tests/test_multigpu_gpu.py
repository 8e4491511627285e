"""Multi-GPU parity (runs when >= 2 GPUs are visible; one process per GPU under torchrun):
replicas stay bit-identical, the reduced gradient is the sum of the local ones, the overlapped
two-bucket reduction and the in-library rt_learner_step_dp equal the plain all-reduce."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.skipif(_gpus() < 2, reason="needs >= 2 GPUs")
def test_data_parallel_replicas_and_in_library_allreduce():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29731", os.path.join(ROOT, "scripts", "dist_check.py")]
    r = subprocess.run(cmd, capture_output=True, timeout=600, cwd=ROOT)
    out = r.stdout.decode() + r.stderr.decode()
    assert r.returncode == 0, out[-3000:]
    assert "replicas identical" in out


@pytest.mark.gpu
@pytest.mark.skipif(_gpus() < 2, reason="needs >= 2 GPUs")
def test_sharded_replay_shard_parity():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29732", os.path.join(ROOT, "scripts", "shard_check.py")]
    r = subprocess.run(cmd, capture_output=True, timeout=600, cwd=ROOT)
    out = r.stdout.decode() + r.stderr.decode()
    assert r.returncode == 0, out[-3000:]
    assert "shard parity ok" in out
