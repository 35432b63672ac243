"""Multi-GPU layout of the hot path: shard decisions by agent, one process per GPU.

A decision depends only on its own N candidates (lib/quoracle/agent/consensus.ex:295-330 never looks across
agents), so the common path has NO collective: rank r owns a contiguous block of the depth-first agent
order (a subtree stays on one device) and writes its own verdicts.  The only exchange is the case the north
star names -- one agent's model pool spanning devices: each rank embeds its local candidates, ONE all-gather
of the [N_local x D] rows, then the similarity kernel on the gathered rows.
Works with any torch.distributed backend (NCCL on GPUs; the gloo tests run it on CPU with an injected
similarity function).
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced [lo, hi) block of n_items for `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def tree_agent_order(depth: int, fanout: int) -> list[tuple[int, ...]]:
    """Depth-first order of a recursive agent tree (BASELINE config 3: depth 4, fan-out 8 -> 4096 leaves;
    all levels: sum_{l=1..depth} fanout^l agents).  Paths are tuples of child indices."""
    order: list[tuple[int, ...]] = []

    def walk(path):
        if path:
            order.append(path)
        if len(path) < depth:
            for c in range(fanout):
                walk(path + (c,))

    walk(())
    return order


def shard_tree(depth: int, fanout: int, rank: int, world: int) -> list[tuple[int, ...]]:
    order = tree_agent_order(depth, fanout)
    lo, hi = shard_range(len(order), rank, world)
    return order[lo:hi]


def gather_verdicts(local: np.ndarray, n_total: int, group=None) -> np.ndarray:
    """All ranks end up with the [n_total] verdict vector (for a single consumer that wants them together);
    uneven shards are padded to the largest block.  Not on the timed path."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    buf = torch.zeros(width, dtype=torch.int32)
    buf[: local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local, dtype=np.int32))
    outs = [torch.zeros(width, dtype=torch.int32) for _ in range(world)]
    dist.all_gather(outs, buf, group=group)
    return np.concatenate([o.numpy()[: hi - lo] for o, (lo, hi) in zip(outs, sizes)])


def decide_split_pool(local_emb, thr, similarity_fn: Callable, group=None):
    """Pool split across ranks: `local_emb` [B, N_local, D] (torch tensor on this rank's device, unit rows).

    One all_gather of the rows -> [B, world*N_local, D] in rank order (pool order = rank-major), then
    `similarity_fn(emb [B,N,D] tensor, thr) -> (ref_ok, medoid)` -- on GPUs that is
    Context.similarity_batch_dev over the gathered device buffer."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    parts = [torch.empty_like(local_emb) for _ in range(world)]
    dist.all_gather(parts, local_emb.contiguous(), group=group)
    full = torch.cat(parts, dim=1).contiguous()
    return similarity_fn(full, thr), full


def connect_split_pool(ctx, max_decisions: int, pool_local: int, group=None):
    """Builds the peer-memory split-pool object for the calling rank: every rank exports its 64-byte handle,
    one all_gather_object exchanges them (control plane only), every rank maps its peers.  The data path
    (`SplitPool.decide_dev`) then runs without any collective call: K2 stores the rows over NVLink."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sp = ctx.split_pool(world, rank, max_decisions, pool_local)
    handles = [None] * world
    dist.all_gather_object(handles, sp.handle(), group=group)
    for p, h in enumerate(handles):
        if p != rank:
            sp.attach(p, h)
    dist.barrier(group=group)
    return sp


def connect_token_split(ctx, max_rows: int, group=None):
    """Same handshake as connect_split_pool for candidates sharded on T (`TokenSplit`)."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    ts = ctx.token_split(world, rank, max_rows)
    handles = [None] * world
    dist.all_gather_object(handles, ts.handle(), group=group)
    for p, h in enumerate(handles):
        if p != rank:
            ts.attach(p, h)
    dist.barrier(group=group)
    return ts


def token_slices(token_ids, lens, world: int):
    """Host helper: split [R, T] token ids (T divisible by world) and their lengths into per-rank slices:
    rank p gets columns [p*T/world, (p+1)*T/world) and len_p = clip(len - p*T/world, 0, T/world)."""
    import numpy as np
    r, t = token_ids.shape
    if t % world:
        raise ValueError("T must divide over the ranks")
    tl = t // world
    out = []
    for p in range(world):
        out.append((np.ascontiguousarray(token_ids[:, p * tl:(p + 1) * tl]),
                    np.clip(lens - p * tl, 0, tl).astype(np.int32)))
    return out
