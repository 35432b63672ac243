"""Multi-GPU exchange paths through the C ABI.  qc_decide_split_dev against the same batch
decided unsplit and against the NCCL all_gather formulation -- bit-identical rows, cosines and verdicts.
World size 1 runs on any GPU box (plumbing: multi-destination K2, self handshake, buffer parity); world size 2
needs two GPUs (`gpurun --gpus 2`) and is skipped otherwise."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _run(world: int, port: int, extra=(), script="split_pool_run.py", decisions=48):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(ROOT / "tools" / script),
           "--decisions", str(decisions), "--iters", "3", *extra]
    env = dict(os.environ, PYTHONPATH=str(ROOT))
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-2000:]
    line = [ln for ln in proc.stdout.splitlines() if ln.startswith("{")][-1]
    return json.loads(line)


def test_split_pool_world_1():
    out = _run(1, 29531)
    assert all(out["checks"].values()) and out["stable_after_timing"]


def test_split_pool_world_1_large_pool_tcgen05():
    out = _run(1, 29532, ("--pool", "16", "--tokens", "96"))
    assert all(out["checks"].values()) and out["stable_after_timing"]


def test_split_pool_world_2():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    out = _run(2, 29533)
    assert out["all_ranks_ok"] and all(out["checks"].values()) and out["stable_after_timing"]
    out = _run(2, 29534, ("--pool", "16", "--tokens", "200"))
    assert out["all_ranks_ok"] and all(out["checks"].values())


# ---- candidates sharded on T (qc_decide_tsplit_dev) ------------------------------------------------------------
TS = dict(script="tsplit_run.py", decisions=12)


def test_token_split_world_1():
    out = _run(1, 29541, ("--pool", "8", "--tokens", "512", "--dim", "768", "--ragged"), **TS)
    assert all(out["checks"].values()), out


def test_token_split_world_2():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    out = _run(2, 29542, ("--pool", "8", "--tokens", "512", "--dim", "768"), **TS)
    assert out["all_ranks_ok"] and all(out["checks"].values()), out
    out = _run(2, 29543, ("--pool", "16", "--tokens", "2048", "--dim", "1536", "--ragged"), **TS)
    assert out["all_ranks_ok"] and all(out["checks"].values()), out
