"""Multi-GPU checks of the peer-memory backend (skipped on a single-GPU box): the fused all-reduce + LAMB kernel
against NCCL all-reduce + the arena LAMB, overflow agreement, and the general peer all-reduce."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_fused_allreduce_lamb_matches_nccl_plus_lamb():
    n = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr",
           "127.0.0.1", "--master-port", "29655", os.path.join(ROOT, "tools", "peer_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-3000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["world"] == n
    assert out["max_abs_diff_mc0"] < 1e-5 and out["overflow_skipped_mc0"] is True
    assert out["allreduce_many_err_mc0"] < 1e-5
    # GEMM -> reduce-scatter push: the first optimizer step is a bitwise-level comparison (the arms then train apart)
    assert out["push_diff_per_step_mc0"][0] < 1e-5
    if "max_abs_diff_mc1" in out:                       # NVLS multicast available on this fabric
        assert out["max_abs_diff_mc1"] < 1e-5 and out["overflow_skipped_mc1"] is True
        assert out["allreduce_many_err_mc1"] < 1e-5


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_weight_gradient_push_into_owner_arena():
    """Kernel level: after one fp32-accumulate GEMM in push mode every owner's shard holds the sum over ranks of
    (locally accumulated + new tile)."""
    n = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr",
           "127.0.0.1", "--master-port", "29656", os.path.join(ROOT, "tools", "push_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-3000:]
    out = json.loads([l for l in res.stdout.splitlines() if l.startswith("{")][-1])
    for name in ("ffn", "qkv", "small"):
        for rec in out[name]:
            assert rec.get("n_bad", 0) == 0, (name, rec)
