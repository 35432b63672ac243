"""Per-role stall profile of the tcgen05 K3 kernel (tuning aid; prints cycles per CTA)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from quoracle_b200 import native, synth  # noqa: E402


def main():
    import torch
    n, d, b = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (16, 768, 65536)))
    ctx = native.Context(0)
    dev = torch.device("cuda:0")
    emb = torch.from_numpy(synth.make_embeddings(2048, n, d, seed=3)).to(dev)
    emb = emb.repeat((b + 2047) // 2048, 1, 1)[:b].contiguous()
    thr = torch.full((b,), 0.5, dtype=torch.float32, device=dev)
    ok = torch.empty(b, dtype=torch.int32, device=dev)
    med = torch.empty(b, dtype=torch.int32, device=dev)
    cos = torch.empty((b, n, n), dtype=torch.float32, device=dev)
    stream = torch.cuda.Stream(device=dev)
    for with_cos in (True, False):
        ctx.stall_profile(True)
        ctx.similarity_batch_dev(emb.data_ptr(), b, n, d, thr.data_ptr(), ok.data_ptr(), med.data_ptr(),
                                 cos.data_ptr() if with_cos else 0, stream.cuda_stream)
        torch.cuda.synchronize()
        st = ctx.stall_profile(True, read=True)
        chunks = ((b + 128 // n - 1) // (128 // n)) * ((d + 31) // 32) / 148
        print(f"N={n} D={d} B={b} cos={with_cos}: chunks/CTA {chunks:.0f}")
        names = {"mma_total": "mma total", "mma_wait_w": "mma wait A operands", "mma_wait_acc": "mma wait acc free",
                 "epi_total": "epilogue total", "epi_wait_acc": "epilogue wait acc",
                 "gather_total": "splitter total", "gather_wait_free": "splitter wait raw",
                 "gather_wait_loads": "splitter wait tmem stage free", "wprod_total": "producer total",
                 "wprod_wait_free": "producer wait raw free",
                 "mma_issue": "epilogue tmem->windows", "mma_commit": "epilogue cosine loop", "gather_fence": "epilogue vote",
                 "aux14": "splitter row load + split", "aux15": "splitter tmem store"}
        st = dict(st)
        for k, label in names.items():
            print(f"  {label:26s} {st[k]:12.0f}  ({st[k] / chunks:8.1f} per chunk)")
    ctx.close()


if __name__ == "__main__":
    main()
