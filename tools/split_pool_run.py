"""Pool split across GPUs: parity check and timing, one process per GPU.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
      tools/split_pool_run.py [--decisions 256] [--pool 8] [--tokens 512] [--dim 768] [--iters 20] [--json out]

CLI over quoracle_b200.exchange_legs.split_pool_leg (the same leg bench.py runs at N > 1): fused peer-memory path vs
NCCL all_gather vs one GPU alone; verdicts, cosines and rows must agree bit for bit.  Works with world size 1."""
import argparse
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from quoracle_b200 import exchange_legs, native, synth  # noqa: E402


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
    w = synth.make_encoder(args.dim, seed=synth.BASE_SEED + 2)
    ctx = native.Context(local, w.table_bf16, w.w_bf16)
    out = exchange_legs.split_pool_leg(ctx, dev, args.decisions, args.pool, args.tokens, args.iters)
    out["stable_after_timing"] = out["checks"]["stable_after_timing"]
    out["all_ranks_ok"] = out["bit_equal_to_unsplit"]
    if rank == 0:
        print(json.dumps(out))
        if args.json:
            Path(args.json).write_text(json.dumps(out, indent=1) + "\n")
    ctx.close()
    dist.destroy_process_group()
    if not out["bit_equal_to_unsplit"]:
        sys.exit(1)


if __name__ == "__main__":
    main()
