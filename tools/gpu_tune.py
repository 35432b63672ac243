"""K1 tuning sweep (not a test): times every tiling variant on BASELINE config 2 (or config 4 with
--dim 1536) and prints the per-role barrier stall profile, so one gpurun call ranks the variants."""
import argparse
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

from quoracle_b200 import native, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--variants", default="")
    ap.add_argument("--ragged", action="store_true", help="len ~ U[T/2, T] (flops counted on valid tokens only)")
    args = ap.parse_args()
    import torch
    dev = torch.device("cuda:0")
    if args.dim == 768:
        b, n, t = 256, 8, 512
        variants = [(128, 1, 64, 1), (128, 2, 64, 1), (112, 2, 64, 1), (104, 2, 64, 1),
                    (128, 2, 64, 1, 2), (112, 2, 64, 1, 2), (104, 2, 64, 1, 2), (128, 2, 64, 1, 4)]
    elif args.dim == 1536:
        b, n, t = 32, 16, 2048
        variants = [(64, 2, 64, 1), (56, 2, 64, 1), (48, 2, 64, 1), (128, 2, 64, 1), (112, 2, 64, 1), (96, 2, 64, 1),
                    (56, 2, 64, 1, 2), (56, 2, 64, 1, 4)]
    else:
        b, n, t = 512, 8, 512
        variants = [(128, 1, 64, 0), (128, 1, 64, 1), (256, 1, 64, 1)]
    if args.variants:
        variants = [tuple(int(x) for x in v.split(",")) for v in args.variants.split(";")]
    w = synth.make_encoder(args.dim, seed=synth.BASE_SEED + 2)
    ctx = native.Context(0, w.table_bf16, w.w_bf16)
    nb = 8
    batches = [synth.make_decisions(b, n, t, seed=synth.BASE_SEED + 50 + i, ragged=args.ragged) for i in range(nb)]
    d_tok = [torch.from_numpy(x.token_ids).to(dev) for x in batches]
    d_len = [torch.from_numpy(x.lens).to(dev) for x in batches]
    d_emb = torch.empty((b * n, args.dim), dtype=torch.float32, device=dev)
    stream = torch.cuda.Stream(device=dev)
    flops = 2.0 * float(np.mean([x.lens.sum() for x in batches])) * args.dim * args.dim   # valid tokens only
    ref = None
    print(f"dim={args.dim} B={b} N={n} T={t}  flops/launch={flops:.3e}")
    for var in variants:
        tb, cg, wbk, gm = var[:4]
        mc = var[4] if len(var) > 4 else 1
        try:
            ctx.set_variant(tb, cg, wbk, gm)
            ctx.set_multicast(mc)
            if ctx.multicast() != mc:
                raise native.QcError(-7, "no such multicast instantiation", "bad_request")
        except native.QcError as e:
            print(f"variant tb={tb} cg={cg} wbk={wbk} gm={gm} mc={mc}: unsupported ({e})")
            continue
        info = ctx.encoder_info()
        for i in range(3):
            ctx.embed_batch_dev(d_tok[i % nb].data_ptr(), d_len[i % nb].data_ptr(), b * n, t, d_emb.data_ptr(),
                                stream.cuda_stream)
        torch.cuda.synchronize()
        try:
            ctx.check_status()
        except native.QcError as e:
            print(f"variant tb={tb} cg={cg} wbk={wbk} gm={gm} mc={mc}: {e} hang={ctx.last_hang()}")
            continue
        ctx.profile_reset()
        ctx.profile_enable(True)
        ctx.stall_profile(True)
        for i in range(args.steps):
            ctx.embed_batch_dev(d_tok[i % nb].data_ptr(), d_len[i % nb].data_ptr(), b * n, t, d_emb.data_ptr(),
                                stream.cuda_stream)
        torch.cuda.synchronize()
        ctx.profile_enable(False)
        st = ctx.kernel_stats()["encode"]
        ms = st["ms"] / st["launches"]
        sp = ctx.stall_profile(False, read=True)
        # parity between variants on batch 0
        ctx.embed_batch_dev(d_tok[0].data_ptr(), d_len[0].data_ptr(), b * n, t, d_emb.data_ptr(), stream.cuda_stream)
        torch.cuda.synchronize()
        emb = d_emb.cpu().numpy()
        if ref is None:
            ref = emb
        diff = float(np.max(np.abs(emb - ref)))
        mt = max(sp["mma_total"], 1.0)
        print(f"tb={tb:3d} cg={cg} wbk={wbk:2d} gm={gm} mc={mc} stages={info['w_stages']}: {ms:7.4f} ms  {flops / ms / 1e9:7.1f} TFLOP/s"
              f"  | MMA wait W {100 * sp['mma_wait_w'] / mt:4.1f}% tok {100 * sp['mma_wait_tokens'] / mt:4.1f}%"
              f" acc {100 * sp['mma_wait_acc'] / mt:4.1f}%  | epi wait {100 * sp['epi_wait_acc'] / max(sp['epi_total'], 1):4.1f}%"
              f"  gather wait free {100 * sp['gather_wait_free'] / max(sp['gather_total'], 1):4.1f}% loads {100 * sp['gather_wait_loads'] / max(sp['gather_total'], 1):4.1f}% fence {100 * sp['gather_fence'] / max(sp['gather_total'], 1):4.1f}%"
              f"  wprod wait {100 * sp['wprod_wait_free'] / max(sp['wprod_total'], 1):4.1f}%"
              f"  | maxdiff vs first {diff:.2e}  | sm clock {sp['mma_total'] / ms / 1e3:.0f} MHz", flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
