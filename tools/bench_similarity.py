"""BASELINE config 5: similarity-kernel sweep (pool 2-64, batch up to 65536, dim 384/768/1536) on one GPU.

Times K3 alone (qc_similarity_batch_dev, embeddings resident in HBM, CUDA events on the launching stream,
inputs larger than L2 or an L2 flush between iterations) and reports achieved algorithmic GB/s against the
measured HBM peak, next to the oracle's C port on the host cores for a bounded sample.  Not a test; writes a
markdown table to stdout (committed under profiles/)."""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from quoracle_b200 import native, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--cpu", action="store_true", help="also time the oracle C port (single thread) on a sample")
    ap.add_argument("--pools", default="2,4,8,16,32,64")
    ap.add_argument("--sim-path", type=int, default=0, help="qc_similarity_set_path (0 auto, 1 mma.sync, 2 tcgen05)")
    args = ap.parse_args()
    import torch
    dev = torch.device("cuda:0")
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    hbm = peaks.get("hbm_gbs", 6650.0)
    ctx = native.Context(0)
    ctx.set_similarity_path(args.sim_path)
    stream = torch.cuda.Stream(device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    pools = [int(x) for x in args.pools.split(",")]
    dims = [384, 768, 1536]
    batches = [256, 4096, 65536] if not args.quick else [4096, 65536]
    print(f"| N | D | B | bytes/decision | ms | decisions/s | GB/s | frac of {hbm:.0f} GB/s | cpu 1-thread decisions/s |")
    print("|---|---|---|---|---|---|---|---|---|")
    for n in pools:
        for d in dims:
            for b in batches:
                nbytes = 4 * n * d + 4 * n * n + 8
                if b * n * d * 4 > 8 * 2**30:
                    continue
                emb = torch.from_numpy(synth.make_embeddings(min(b, 2048), n, d, seed=n * d)).to(dev)
                if b > emb.shape[0]:
                    emb = emb.repeat((b + emb.shape[0] - 1) // emb.shape[0], 1, 1)[:b].contiguous()
                thr = torch.full((b,), 0.5, dtype=torch.float32, device=dev)
                ok = torch.empty(b, dtype=torch.int32, device=dev)
                med = torch.empty(b, dtype=torch.int32, device=dev)
                cos = torch.empty((b, n, n), dtype=torch.float32, device=dev)

                def run():
                    ctx.similarity_batch_dev(emb.data_ptr(), b, n, d, thr.data_ptr(), ok.data_ptr(), med.data_ptr(),
                                             cos.data_ptr(), stream.cuda_stream)

                for _ in range(3):
                    run()
                torch.cuda.synchronize()
                times = []
                for _ in range(5):
                    flush.zero_()                                   # evict the inputs from L2 between iterations
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    run()
                    e1.record(stream)
                    torch.cuda.synchronize()
                    times.append(e0.elapsed_time(e1))
                ms = float(np.median(times))
                gbs = b * nbytes / (ms * 1e-3) / 1e9
                cpu = ""
                if args.cpu and b == batches[-1]:
                    from oracle import cport
                    sample = emb[: min(b, max(8, 200000 // (n * n * d // 64 + 1)))].cpu().numpy()
                    t0 = time.perf_counter()
                    cport.similarity_batch(sample, np.full(sample.shape[0], 0.5), want_full=True, threads=1,
                                           want_cos=False, native=True)
                    cpu = f"{sample.shape[0] / (time.perf_counter() - t0):.0f}"
                print(f"| {n} | {d} | {b} | {nbytes} | {ms:.4f} | {b / (ms * 1e-3):.3e} | {gbs:.0f} | {gbs / hbm:.3f} | {cpu} |",
                      flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
