"""Single-call latency of the host C ABI (qc_decide_batch, pinned buffers) for small batches: what one agent
(or a small collector flush) pays, next to the batched throughput.  Not a test."""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from quoracle_b200 import native, synth  # noqa: E402

w = synth.make_encoder(768, seed=synth.BASE_SEED + 2)
ctx = native.Context(0, w.table_bf16, w.w_bf16)
print("| decisions per call | p50 us | p95 us | decisions/s |")
print("|---|---|---|---|")
for b in (1, 2, 8, 32, 64, 128, 256, 1024):
    batch = synth.make_decisions(b, 8, 512, seed=5)
    tok = native.PinnedArray((b, 8, 512), np.int32)
    tok.array[...] = batch.token_ids
    lens = native.PinnedArray((b, 8), np.int32)
    lens.array[...] = batch.lens
    thr = native.PinnedArray((b,), np.float32)
    thr.array[...] = batch.thr
    out = native.DecideResult(native.PinnedArray((b,), np.int32).array, native.PinnedArray((b,), np.int32).array, None, None)
    for _ in range(10):
        ctx.decide_batch(tok.array, lens.array, thr.array, want_cos=False, out=out)
    ts = []
    for _ in range(200):
        t0 = time.perf_counter()
        ctx.decide_batch(tok.array, lens.array, thr.array, want_cos=False, out=out)
        ts.append(time.perf_counter() - t0)
    ts = np.array(ts) * 1e6
    print(f"| {b} | {np.percentile(ts, 50):.0f} | {np.percentile(ts, 95):.0f} | {b / (np.percentile(ts, 50) * 1e-6):.3e} |", flush=True)
ctx.close()
