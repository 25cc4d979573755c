"""N > 1 on real GPUs (skipped on a one-GPU box): the set-sharded path and the peer-store all-gather under NCCL,
two ranks, through tools/dist_check.py -- log_likelihood_distributed bit-identical to one GPU (set-sharded) / 1e-13
(time-sharded), 200 epochs of jxp_peer_allgather_f64 with ragged blocks and lagging ranks equal to ncclAllGather."""

import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_ranks_set_sharding_and_peer_all_gather():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "dist_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "wrong epochs 0 of 200" in out.stdout and "NOT available" not in out.stdout
