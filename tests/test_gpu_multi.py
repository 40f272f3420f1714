"""Multi-GPU parity: launches tests/multi_gpu_check.py under torchrun -- the vertex-partitioned CSR
(parts in separate processes, mapped into each other through VMM file descriptors, read by the kernel's
MULTI variant) gives walks bit-identical to the replicated graph; data-parallel SGNS leaves identical
tables on every rank.  With >= 2 GPUs: one rank per GPU over NCCL / NVLink.  On a single-GPU box the same
check runs with both ranks on cuda:0 (gloo plumbing), so the partitioned path is never untested."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partitioned_walk_and_dp_sgns_two_ranks_one_gpu():
    import torch
    if torch.cuda.device_count() >= 2:
        pytest.skip("covered by the two-GPU test")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29534", os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env={**os.environ, "N2V_CHECK_ONE_GPU": "1"})
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "MULTI_GPU_CHECK OK" in out.stdout


def test_partitioned_walk_and_dp_sgns_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "MULTI_GPU_CHECK OK" in out.stdout
