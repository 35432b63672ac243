"""Ragged batches (len ~ U[T/2, T]): K1 time with the dense block walk (device entry points) against the compacted
block list (host entry points), BASELINE config 2 shape.  Prints K1 ms per launch and TFLOP/s on valid tokens."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from quoracle_b200 import native, synth  # noqa: E402


def main():
    import torch
    dev = torch.device("cuda:0")
    b, n, t, d = 256, 8, 512, 768
    w = synth.make_encoder(d, seed=synth.BASE_SEED + 2)
    ctx = native.Context(0, w.table_bf16, w.w_bf16)
    batches = [synth.make_decisions(b, n, t, seed=synth.BASE_SEED + 50 + i, ragged=True) for i in range(8)]
    flops = 2.0 * float(np.mean([x.lens.sum() for x in batches])) * d * d
    pins = []
    for x in batches:
        pt, pl, pth = native.PinnedArray((b, n, t), np.int32), native.PinnedArray((b, n), np.int32), native.PinnedArray((b,), np.float32)
        pt.array[...], pl.array[...], pth.array[...] = x.token_ids, x.lens, x.thr
        pins.append((pt, pl, pth))
    d_tok = [torch.from_numpy(x.token_ids).to(dev) for x in batches]
    d_len = [torch.from_numpy(x.lens).to(dev) for x in batches]
    d_emb = torch.empty((b * n, d), dtype=torch.float32, device=dev)
    outs = [native.DecideResult(native.PinnedArray((b,), np.int32).array, native.PinnedArray((b,), np.int32).array, None, None)
            for _ in range(4)]

    def dense(steps):
        for i in range(steps):
            ctx.embed_batch_dev(d_tok[i % 8].data_ptr(), d_len[i % 8].data_ptr(), b * n, t, d_emb.data_ptr())
        ctx.synchronize()

    tb = ctx.token_block(t)
    bpc = -(-t // tb)
    maps = []
    for x in batches:
        nb = -(-x.lens.reshape(-1) // tb)
        ids = np.concatenate([r * bpc + np.arange(k, dtype=np.int32) for r, k in enumerate(nb)]).astype(np.int32)
        maps.append(torch.from_numpy(ids).to(dev))

    def compact(steps):
        for i in range(steps):
            ctx.hint_block_map(maps[i % 8].data_ptr(), maps[i % 8].numel())
            ctx.embed_batch_dev(d_tok[i % 8].data_ptr(), d_len[i % 8].data_ptr(), b * n, t, d_emb.data_ptr())
        ctx.synchronize()

    for name, fn in (("dense walk, empty blocks skipped", dense), ("compacted block list", compact)):
        fn(6)
        ctx.profile_reset()
        ctx.profile_enable(True)
        fn(40)
        ctx.profile_enable(False)
        st = ctx.kernel_stats()["encode"]
        ms = st["ms"] / st["launches"]
        print(f"{name:36s} K1 {ms:.4f} ms/launch  {flops / ms / 1e9:7.1f} TFLOP/s on valid tokens "
              f"({st['launches']} launches)")

    # end to end through the pipelined host API (what a caller sees): wall clock, four batches in flight
    import time

    def host_pipeline(steps):
        for i in range(steps):
            k = i % 4
            if i >= 4:
                ctx.decide_batch_wait(k)
            pt, pl, pth = pins[i % 8]
            ctx.decide_batch_submit(k, pt.array, pl.array, pth.array, outs[k])
        for i in range(max(0, steps - 4), steps):
            ctx.decide_batch_wait(i % 4)

    for mode, name in ((1, "host API, block list (packing off)"), (2, "host API, packed token stream")):
        ctx.set_packing(mode)
        host_pipeline(8)
        t0 = time.perf_counter()
        host_pipeline(200)
        dt = time.perf_counter() - t0
        print(f"{name:36s} {dt / 200 * 1e3:.4f} ms/step  {200 * b / dt:9.0f} decisions/s  "
              f"{flops * 200 / dt / 1e12:7.1f} TFLOP/s on valid tokens")
    ctx.set_packing(0)
    # CSR entry point, pipelined: nothing padded on the host or on the wire
    csr = []
    for x in batches:
        lens = x.lens.reshape(-1)
        off = np.zeros(lens.shape[0] + 1, np.int32)
        np.cumsum(lens, out=off[1:])
        flat = np.concatenate([x.token_ids.reshape(-1, t)[r, :lens[r]] for r in range(lens.shape[0])]).astype(np.int32)
        pt, po, pth = native.PinnedArray(flat.shape, np.int32), native.PinnedArray(off.shape, np.int32), native.PinnedArray((b,), np.float32)
        pt.array[...], po.array[...], pth.array[...] = flat, off, x.thr
        csr.append((pt, po, pth))

    def csr_pipeline(steps):
        for i in range(steps):
            k = i % 4
            if i >= 4:
                ctx.decide_batch_wait(k)
            pt, po, pth = csr[i % 8]
            ctx.decide_ragged_submit(k, pt.array, po.array, n, pth.array, outs[k])
        for i in range(max(0, steps - 4), steps):
            ctx.decide_batch_wait(i % 4)

    csr_pipeline(8)
    t0 = time.perf_counter()
    csr_pipeline(200)
    dt = time.perf_counter() - t0
    print(f"{'host API, CSR input (pipelined)':36s} {dt / 200 * 1e3:.4f} ms/step  {200 * b / dt:9.0f} decisions/s  "
          f"{flops * 200 / dt / 1e12:7.1f} TFLOP/s on valid tokens; H2D {csr[0][0].array.nbytes / 2**20:.1f} MiB/step "
          f"against {b * n * t * 4 / 2**20:.1f} MiB padded")
    full = 2.0 * b * n * t * d * d
    print(f"valid tokens: {flops / full:.3f} of the padded shape; blocks of 128: "
          f"{np.mean([np.ceil(x.lens / 128).sum() for x in batches]) / (b * n * 4):.3f} non-empty")
    ctx.close()


if __name__ == "__main__":
    main()
