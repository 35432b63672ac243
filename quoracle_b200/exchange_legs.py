"""The two cross-GPU exchange paths of the hot path as measurable, self-checking legs (one process per GPU).

  split_pool_leg  one agent's model pool spanning devices (north star: "a single all-gather of embedding rows over
                  NVLink only when one agent's model pool spans devices"): qc_decide_split_dev -- K1 on the local pool
                  members, K2 stores every normalised row into every rank's gathered buffer over peer memory, flag
                  handshake, K3 -- against the same path written with an NCCL all_gather and against one GPU deciding
                  the whole pool alone.  Rows, cosines and verdicts must be bit-identical in all three.
  tsplit_leg      long candidates sharded on T (BASELINE config 4 shape): qc_decide_tsplit_dev against the unsplit
                  encoder.  Embeddings agree to fp32 summation order; verdicts wherever the unsplit margin allows.

Used by bench.py (extra keys of the N > 1 line, so the driver's SCALE record carries them), by tools/split_pool_run.py
and tools/tsplit_run.py, and by tests/test_gpu_split_pool.py.  torch.distributed (NCCL) is plumbing only: the handle
exchange, the barrier and the max-over-ranks of the device timings.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import native, sharding, synth

EMB_TOL = 2e-6      # unit rows, fp32 association only (tsplit vs unsplit)
COS_TOL = 1e-5
MARGIN = 5e-6       # 10x the cosine noise between the two summation orders


def _time(fn, stream, iters, dev):
    import torch
    import torch.distributed as dist
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(iters):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


def split_pool_leg(ctx: native.Context, dev, decisions: int = 256, pool: int = 8, tokens: int = 512, iters: int = 20,
                   seed: int = synth.BASE_SEED + 77) -> dict:
    """All ranks call this with the same arguments; ``ctx`` is the rank's context (its encoder fixes dim)."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    b, n, t, d = decisions, pool, tokens, ctx.dim
    if n % world:
        raise ValueError("pool must divide over the ranks")
    nl = n // world
    batch = synth.make_decisions(b, n, t, seed=seed, thresholds=(0.85, 0.995, 0.9995))
    tok_all = torch.from_numpy(batch.token_ids.reshape(b, n, t)).to(dev)
    len_all = torch.from_numpy(batch.lens.reshape(b, n)).to(dev)
    tok_loc = tok_all[:, rank * nl:(rank + 1) * nl].contiguous()
    len_loc = len_all[:, rank * nl:(rank + 1) * nl].contiguous()
    thr = torch.from_numpy(batch.thr).to(dev)
    stream = torch.cuda.Stream(device=dev)

    def outs():
        return (torch.empty(b, dtype=torch.int32, device=dev), torch.empty(b, dtype=torch.int32, device=dev),
                torch.empty((b, n, n), dtype=torch.float32, device=dev))

    sp = sharding.connect_split_pool(ctx, b, nl)
    f_ok, f_med, f_cos = outs()
    n_ok, n_med, n_cos = outs()
    u_ok, u_med, u_cos = outs()
    emb_loc = torch.empty((b, nl, d), dtype=torch.float32, device=dev)
    emb_full = torch.empty((b, n, d), dtype=torch.float32, device=dev)
    u_emb = torch.empty((b, n, d), dtype=torch.float32, device=dev)
    parts = [torch.empty_like(emb_loc) for _ in range(world)]

    def fused():
        sp.decide_dev(tok_loc.data_ptr(), len_loc.data_ptr(), b, t, thr.data_ptr(), f_ok.data_ptr(), f_med.data_ptr(),
                      f_cos.data_ptr(), stream.cuda_stream)

    def nccl():
        ctx.embed_batch_dev(tok_loc.data_ptr(), len_loc.data_ptr(), b * nl, t, emb_loc.data_ptr(), stream.cuda_stream)
        with torch.cuda.stream(stream):
            dist.all_gather(parts, emb_loc)
            torch.cat(parts, dim=1, out=emb_full)
        ctx.similarity_batch_dev(emb_full.data_ptr(), b, n, d, thr.data_ptr(), n_ok.data_ptr(), n_med.data_ptr(),
                                 n_cos.data_ptr(), stream.cuda_stream)

    def unsplit():
        ctx.decide_batch_dev(tok_all.data_ptr(), len_all.data_ptr(), b, n, t, thr.data_ptr(), u_ok.data_ptr(),
                             u_med.data_ptr(), u_cos.data_ptr(), u_emb.data_ptr(), stream.cuda_stream)

    for fn in (fused, nccl, unsplit):
        fn()
    torch.cuda.synchronize()
    ctx.check_status()
    gp, gb = sp.gathered()
    assert gb == b
    g_emb = torch.empty((b, n, d), dtype=torch.float32, device=dev)
    ctypes.CDLL("libcudart.so").cudaMemcpy(ctypes.c_void_p(g_emb.data_ptr()), ctypes.c_void_p(gp),
                                           ctypes.c_size_t(g_emb.numel() * 4), 3)
    torch.cuda.synchronize()
    checks = {
        "gathered_rows_equal_unsplit_rows": bool(torch.equal(g_emb, u_emb)),
        "nccl_rows_equal_unsplit_rows": bool(torch.equal(emb_full, u_emb)),
        "fused_verdicts_equal_unsplit": bool(torch.equal(f_ok, u_ok) and torch.equal(f_med, u_med)),
        "fused_cos_equal_unsplit": bool(torch.equal(f_cos, u_cos)),
        "nccl_verdicts_equal_unsplit": bool(torch.equal(n_ok, u_ok) and torch.equal(n_med, u_med)),
    }
    times = {name: _time(fn, stream, iters, dev) for name, fn in
             (("fused", fused), ("nccl", nccl), ("unsplit", unsplit))} if iters else {}
    ctx.check_status()
    # the same repeated batch must keep giving the same answer (buffer parity, epochs)
    checks["stable_after_timing"] = bool(torch.equal(f_cos, u_cos) and torch.equal(f_med, u_med))
    flag = torch.tensor([int(all(checks.values()))], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    ok_hist = torch.bincount(u_ok.to(torch.int64), minlength=2).tolist()
    sp.close()
    return {"world": world, "decisions": b, "pool": n, "pool_local": nl, "tokens": t, "dim": d, "checks": checks,
            "bit_equal_to_unsplit": bool(flag.item()), "ref_ok_zeros": int(ok_hist[0]), "ref_ok_ones": int(ok_hist[1]),
            "ms": times.get("fused"), "nccl_ms": times.get("nccl"), "unsplit_ms": times.get("unsplit"),
            "decisions_per_s": b / (times["fused"] * 1e-3) if times else None,
            "how": "qc_decide_split_dev: K2 stores rows into every rank's gathered buffer over peer memory (no "
                   "collective call); nccl_ms = embed -> torch all_gather -> cat -> K3; unsplit_ms = one GPU alone; "
                   "CUDA events on the launching stream, max over ranks"}


def tsplit_leg(ctx: native.Context, dev, decisions: int = 32, pool: int = 16, tokens: int = 2048, iters: int = 10,
               ragged: bool = False, seed: int = synth.BASE_SEED + 91) -> dict:
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    b, n, t, d = decisions, pool, tokens, ctx.dim
    if t % world:
        raise ValueError("T must divide over the ranks")
    tl = t // world
    batch = synth.make_decisions(b, n, t, seed=seed, thresholds=(0.85, 0.995, 0.9995))
    lens = batch.lens.reshape(b * n).copy()
    if ragged:
        rng = np.random.default_rng(5)
        lens = rng.integers(1, t + 1, size=lens.shape).astype(np.int32)
        lens[::7] = rng.integers(1, tl + 1, size=lens[::7].shape)        # some end inside the first slice
        lens[3] = 0                                                        # an empty candidate
    tok = batch.token_ids.reshape(b * n, t)
    tok_s, len_s = sharding.token_slices(tok, lens, world)[rank]
    d_tok_all = torch.from_numpy(tok).to(dev)
    d_len_all = torch.from_numpy(lens).to(dev)
    d_tok_loc = torch.from_numpy(tok_s).to(dev)
    d_len_loc = torch.from_numpy(len_s).to(dev)
    thr = torch.from_numpy(batch.thr).to(dev)
    stream = torch.cuda.Stream(device=dev)

    def outs():
        return (torch.empty(b, dtype=torch.int32, device=dev), torch.empty(b, dtype=torch.int32, device=dev),
                torch.empty((b, n, n), dtype=torch.float32, device=dev),
                torch.empty((b * n, d), dtype=torch.float32, device=dev))

    ts = sharding.connect_token_split(ctx, b * n)
    s_ok, s_med, s_cos, s_emb = outs()
    u_ok, u_med, u_cos, u_emb = outs()

    def tsplit():
        ts.decide_dev(d_tok_loc.data_ptr(), d_len_loc.data_ptr(), b, n, tl, thr.data_ptr(), s_ok.data_ptr(),
                      s_med.data_ptr(), s_cos.data_ptr(), s_emb.data_ptr(), stream.cuda_stream)

    def unsplit():
        ctx.decide_batch_dev(d_tok_all.data_ptr(), d_len_all.data_ptr(), b, n, t, thr.data_ptr(), u_ok.data_ptr(),
                             u_med.data_ptr(), u_cos.data_ptr(), u_emb.data_ptr(), stream.cuda_stream)

    tsplit()
    unsplit()
    torch.cuda.synchronize()
    ctx.check_status()
    emb_err = float((s_emb - u_emb).abs().max())
    cos_err = float((s_cos - u_cos).abs().max())
    thr_margin = (u_cos[:, 0, 1:] - thr[:, None]).abs().min(dim=1).values
    means = u_cos.mean(dim=2)
    top2 = means.topk(2, dim=1).values
    vote_margin = top2[:, 0] - top2[:, 1]
    safe_t, safe_v = thr_margin > MARGIN, vote_margin > MARGIN
    mine = s_emb.clone()
    dist.broadcast(mine, src=0)                         # every rank must hold the same bits
    checks = {
        "emb_within_tol": emb_err < EMB_TOL,
        "cos_within_tol": cos_err < COS_TOL,
        "ref_ok_equal_where_safe": bool(torch.equal(s_ok[safe_t], u_ok[safe_t])),
        "medoid_equal_where_safe": bool(torch.equal(s_med[safe_v], u_med[safe_v])),
        "all_ranks_same_bits": bool(torch.equal(mine, s_emb)),
        "unit_or_zero_rows": bool(((s_emb.norm(dim=1) - 1).abs() < 1e-5).logical_or(s_emb.norm(dim=1) == 0).all()),
    }
    flag = torch.tensor([int(all(checks.values()))], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    times = {name: _time(fn, stream, iters, dev) for name, fn in (("tsplit", tsplit), ("unsplit", unsplit))} if iters else {}
    ctx.check_status()
    ok_hist = torch.bincount(u_ok.to(torch.int64), minlength=2).tolist()
    ts.close()
    return {"world": world, "decisions": b, "pool": n, "tokens": t, "tokens_local": tl, "dim": d, "ragged": ragged,
            "max_abs_emb_diff": emb_err, "max_abs_cos_diff": cos_err,
            "margin_rejected": {"thr": int((~safe_t).sum()), "vote": int((~safe_v).sum()), "margin": MARGIN},
            "indices_equal": bool(checks["ref_ok_equal_where_safe"] and checks["medoid_equal_where_safe"]),
            "ref_ok_zeros": int(ok_hist[0]), "ref_ok_ones": int(ok_hist[1]),
            "checks": checks, "all_ranks_ok": bool(flag.item()),
            "ms": times.get("tsplit"), "unsplit_ms": times.get("unsplit"),
            "decisions_per_s": b / (times["tsplit"] * 1e-3) if times else None,
            "how": "qc_decide_tsplit_dev: every rank encodes its token slice of every candidate, fp32 partial sums "
                   "cross over peer memory, rank-ordered sum, K3 on every rank; unsplit_ms = one GPU alone; CUDA events, "
                   "max over ranks"}
