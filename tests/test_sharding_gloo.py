"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: verdict gathering, the split-pool all-gather path and the
consensus round loop over a tree sharded by agent.  The similarity function is injected (the oracle here, K3 on GPUs)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import numeric
    from quoracle_b200 import sharding, synth

    # 1. shard by agent, no collective on the data path; gather verdicts afterwards
    b, n, d = 37, 8, 64
    emb = synth.make_embeddings(b, n, d, seed=5)
    thr = np.full(b, 0.5, np.float32)
    lo, hi = sharding.shard_range(b, rank, world)
    ok, med, _, _ = numeric.decide_batch(emb[lo:hi], thr[lo:hi])
    all_ok = sharding.gather_verdicts(ok, b)
    all_med = sharding.gather_verdicts(med, b)
    ref_ok, ref_med, _, _ = numeric.decide_batch(emb, thr)
    assert np.array_equal(all_ok, ref_ok) and np.array_equal(all_med, ref_med)

    # 2. one agent's pool spans the two ranks: all-gather of the local rows, then the similarity kernel
    n_local = n // world
    local = torch.from_numpy(emb[:, rank * n_local:(rank + 1) * n_local].copy())

    def sim(full, t):
        o, m, _, _ = numeric.decide_batch(full.numpy(), t)
        return o, m

    (ok2, med2), full = sharding.decide_split_pool(local, thr, sim)
    assert full.shape == (b, n, d) and np.array_equal(full.numpy(), emb)
    assert np.array_equal(ok2, ref_ok) and np.array_equal(med2, ref_med)
    # 3. the consensus round loop over an agent tree sharded by rank (depth-first blocks: a subtree stays on one
    #    device): every rank sweeps its own agents (no collective), the union of the results is what one rank computes
    import pickle
    from quoracle_b200 import consensus_driver
    from quoracle_b200.collector import DecisionCollector
    from tests.test_host_logic import _tree_agents, _validate, make_emb, scripted_query_fn
    agents = _tree_agents(3, 4)                                   # 4 + 16 + 64 agents
    pool = {a["agent_id"]: len(a["state"]["model_histories"]) for a in agents}
    qf = scripted_query_fn(21, lambda i: pool[i])
    lo, hi = sharding.shard_range(len(agents), rank, world)
    e1, _ = make_emb()
    mine = consensus_driver.run_consensus_rounds(DecisionCollector(e1), agents[lo:hi], qf, _validate)
    mine.pop("__sweeps__")
    gathered = [None] * world
    dist.all_gather_object(gathered, pickle.dumps(mine))
    union = {}
    for blob in gathered:
        union.update(pickle.loads(blob))
    e2, _ = make_emb()
    whole = consensus_driver.run_consensus_rounds(DecisionCollector(e2), agents, qf, _validate)
    whole.pop("__sweeps__")
    assert set(union) == set(whole) and len(union) == 84
    for k, v in whole.items():
        u = union[k]
        assert u[0] == v[0]
        if v[0] == "ok":
            assert u[1]["kind"] == v[1]["kind"] and u[1]["action"] == v[1]["action"] and u[1]["confidence"] == v[1]["confidence"]
        else:
            assert u[1] == v[1]
    Path(out_dir, f"ok{rank}").write_text("ok")
    dist.destroy_process_group()


def test_two_rank_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_token_slices_partition_every_candidate():
    """Host helper of the T-sharded path: slices tile the token matrix, slice lengths add up to the length."""
    import numpy as np

    from quoracle_b200 import sharding
    rng = np.random.default_rng(3)
    tok = rng.integers(0, 1000, size=(37, 96), dtype=np.int32)
    lens = rng.integers(0, 97, size=37).astype(np.int32)
    lens[:3] = (0, 96, 24)
    for world in (1, 2, 4, 8):
        parts = sharding.token_slices(tok, lens, world)
        assert len(parts) == world
        assert np.array_equal(np.concatenate([p[0] for p in parts], axis=1), tok)
        assert np.array_equal(sum(p[1].astype(np.int64) for p in parts), lens)
        tl = 96 // world
        for p, (ids, ln) in enumerate(parts):
            assert ids.flags["C_CONTIGUOUS"] and ids.shape == (37, tl) and ln.dtype == np.int32
            assert np.all((ln >= 0) & (ln <= tl))
            # a slice is non-empty only if every earlier slice is full
            if p > 0:
                assert np.all((ln == 0) | (parts[p - 1][1] == tl))
    import pytest
    with pytest.raises(ValueError):
        sharding.token_slices(tok, lens, 5)
