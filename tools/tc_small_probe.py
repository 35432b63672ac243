"""Probe for the tcgen05 K3 kernel on large shapes (tuning aid): prints after every shape, then the watchdog record."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from quoracle_b200 import native, synth

ctx = native.Context(0)
dev = torch.device("cuda:0")
shapes = [(16, 768, 65536), (16, 1536, 4096), (16, 1536, 65536), (32, 384, 4096), (32, 384, 65536)]
if len(sys.argv) > 1:
    shapes = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]]
for n, d, b in shapes:
    e = torch.from_numpy(synth.make_embeddings(min(b, 2048), n, d, seed=n + d)).to(dev)
    e = e.repeat((b + e.shape[0] - 1) // e.shape[0], 1, 1)[:b].contiguous()
    thr = torch.full((b,), 0.5, dtype=torch.float32, device=dev)
    ok = torch.empty(b, dtype=torch.int32, device=dev)
    med = torch.empty(b, dtype=torch.int32, device=dev)
    cos = torch.empty((b, n, n), dtype=torch.float32, device=dev)
    for rep in range(4):
        ctx.similarity_batch_dev(e.data_ptr(), b, n, d, thr.data_ptr(), ok.data_ptr(), med.data_ptr(), cos.data_ptr())
    ctx.synchronize()
    try:
        ctx.check_status()
        print(n, d, b, "ok", int(med[:3].sum()), flush=True)
    except native.QcError as err:
        print(n, d, b, "ERROR", err, ctx.last_hang(), flush=True)
        break
print("done")
