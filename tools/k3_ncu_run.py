"""One K3 launch shape for ncu (`ncu --set full -k regex:qc_sim ... python tools/k3_ncu_run.py N D B`): embeddings resident in
HBM, a few launches of qc_similarity_batch_dev.  Not a test."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from quoracle_b200 import native, synth  # noqa: E402

n, d, b = (int(x) for x in sys.argv[1:4])
import torch  # noqa: E402
dev = torch.device("cuda:0")
ctx = native.Context(0)
emb = torch.from_numpy(synth.make_embeddings(min(b, 2048), n, d, seed=n * d)).to(dev)
emb = emb.repeat((b + emb.shape[0] - 1) // emb.shape[0], 1, 1)[:b].contiguous()
thr = torch.full((b,), 0.5, dtype=torch.float32, device=dev)
ok = torch.empty(b, dtype=torch.int32, device=dev)
med = torch.empty(b, dtype=torch.int32, device=dev)
cos = torch.empty((b, n, n), dtype=torch.float32, device=dev)
stream = torch.cuda.Stream(device=dev)
torch.cuda.synchronize()
for _ in range(4):
    ctx.similarity_batch_dev(emb.data_ptr(), b, n, d, thr.data_ptr(), ok.data_ptr(), med.data_ptr(), cos.data_ptr(), stream.cuda_stream)
torch.cuda.synchronize()
ctx.check_status()
print("done", n, d, b, int(med[:4].sum()))
ctx.close()
