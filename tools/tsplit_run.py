"""Long candidates sharded on T across GPUs (BASELINE config 4): parity and timing.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 \
      tools/tsplit_run.py [--decisions 32] [--pool 16] [--tokens 2048] [--dim 1536] [--ragged] [--iters 10]

CLI over quoracle_b200.exchange_legs.tsplit_leg (the same leg bench.py runs at N > 1)."""
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
    w = synth.make_encoder(args.dim, seed=synth.BASE_SEED + 4)
    ctx = native.Context(local, w.table_bf16, w.w_bf16)
    out = exchange_legs.tsplit_leg(ctx, dev, args.decisions, args.pool, args.tokens, args.iters, ragged=args.ragged)
    if rank == 0:
        print(json.dumps(out))
        if args.json:
            Path(args.json).write_text(json.dumps(out, indent=1) + "\n")
    ctx.close()
    dist.destroy_process_group()
    if not out["all_ranks_ok"]:
        sys.exit(1)


if __name__ == "__main__":
    main()
