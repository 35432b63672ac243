"""Long candidates sharded on T across GPUs (BASELINE config 4): parity and timing.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 \
      tools/tsplit_run.py [--decisions 32] [--pool 16] [--tokens 2048] [--dim 1536] [--ragged] [--iters 10]

  tsplit   qc_decide_tsplit_dev: every rank encodes its token slice of every candidate, partial sums cross over
           peer memory, every rank finalises and decides                         (no collective call)
  unsplit  qc_decide_batch_dev on the whole sequences on every rank (what one GPU computes alone)
Embeddings must agree to fp32 summation-order noise, verdicts wherever the unsplit margin allows."""
import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from quoracle_b200 import native, sharding, synth  # noqa: E402

EMB_TOL = 2e-6      # unit rows, fp32 association only
COS_TOL = 1e-5
MARGIN = 5e-6      # 10x the cosine noise between the two summation orders


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--decisions", type=int, default=32)
    ap.add_argument("--pool", type=int, default=16)
    ap.add_argument("--tokens", type=int, default=2048)
    ap.add_argument("--dim", type=int, default=1536)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--ragged", action="store_true")
    ap.add_argument("--json", default="")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29518")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    b, n, t, d = args.decisions, args.pool, args.tokens, args.dim
    assert t % world == 0
    tl = t // world
    w = synth.make_encoder(d, seed=synth.BASE_SEED + 4)
    ctx = native.Context(local, w.table_bf16, w.w_bf16)
    batch = synth.make_decisions(b, n, t, seed=synth.BASE_SEED + 91)
    lens = batch.lens.copy()
    if args.ragged:
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
    # every rank must hold the same bits
    mine = s_emb.clone()
    dist.broadcast(mine, src=0)
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

    def time_it(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.iters):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / args.iters], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    times = {name: time_it(fn) for name, fn in (("tsplit", tsplit), ("unsplit", unsplit))} if args.iters else {}
    ctx.check_status()
    out = {"world": world, "decisions": b, "pool": n, "tokens": t, "tokens_local": tl, "dim": d, "ragged": args.ragged,
           "max_abs_emb_diff": emb_err, "max_abs_cos_diff": cos_err,
           "margin_rejected": {"thr": int((~safe_t).sum()), "vote": int((~safe_v).sum())},
           "checks": checks, "all_ranks_ok": bool(flag.item()), "ms_per_step_max_over_ranks": times,
           "decisions_per_s": {k: b / (v * 1e-3) for k, v in times.items()}}
    if rank == 0:
        print(json.dumps(out))
        if args.json:
            Path(args.json).write_text(json.dumps(out, indent=1) + "\n")
    ts.close()
    ctx.close()
    dist.destroy_process_group()
    if not flag.item():
        sys.exit(1)


if __name__ == "__main__":
    main()
