"""Pool split across GPUs: parity check and timing, one process per GPU.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
      tools/split_pool_run.py [--decisions 256] [--pool 8] [--tokens 512] [--dim 768] [--iters 20] [--json out]

Three ways to decide the same batch, all through the C ABI:
  fused    qc_decide_split_dev: K1 -> K2 stores rows into every rank's gathered buffer (peer memory) -> flag
           handshake -> K3                                                    (no collective call)
  nccl     qc_embed_batch_dev -> torch.distributed.all_gather (NCCL) -> cat -> qc_similarity_batch_dev
  unsplit  qc_decide_batch_dev with the whole pool on every rank (what one GPU computes alone)
and asserts that verdicts, cosines and embedding rows agree bit for bit.  Works with world size 1 (plumbing)."""
import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from quoracle_b200 import native, sharding, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--decisions", type=int, default=256)
    ap.add_argument("--pool", type=int, default=8)
    ap.add_argument("--tokens", type=int, default=512)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--iters", type=int, default=20)
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
        os.environ.setdefault("MASTER_PORT", "29517")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    b, n, t, d = args.decisions, args.pool, args.tokens, args.dim
    assert n % world == 0, "pool must divide over the ranks"
    nl = n // world
    w = synth.make_encoder(d, seed=synth.BASE_SEED + 2)
    ctx = native.Context(local, w.table_bf16, w.w_bf16)
    batch = synth.make_decisions(b, n, t, seed=synth.BASE_SEED + 77)
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
    import ctypes
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
    flag = torch.tensor([int(all(checks.values()))], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)

    def time_it(fn):
        for _ in range(3):
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

    times = {name: time_it(fn) for name, fn in (("fused", fused), ("nccl", nccl), ("unsplit", unsplit))} if args.iters else {}
    ctx.check_status()
    # the same repeated batch must keep giving the same answer (buffer parity, epochs)
    again = bool(torch.equal(f_cos, u_cos) and torch.equal(f_med, u_med))
    out = {"world": world, "decisions": b, "pool": n, "pool_local": nl, "tokens": t, "dim": d, "checks": checks,
           "stable_after_timing": again, "all_ranks_ok": bool(flag.item()), "ms_per_step_max_over_ranks": times,
           "decisions_per_s": {k: b / (v * 1e-3) for k, v in times.items()}}
    if rank == 0:
        print(json.dumps(out))
        if args.json:
            Path(args.json).write_text(json.dumps(out, indent=1) + "\n")
    sp.close()
    ctx.close()
    dist.destroy_process_group()
    if not (flag.item() and again):
        sys.exit(1)


if __name__ == "__main__":
    main()
