"""One K3 configuration for an ncu capture (BASELINE config-5 corner): k3_one.py [pool dim batch],
default pool 8, dim 768, batch 65536; pools above 8 run the tcgen05 kernel."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from quoracle_b200 import native, synth  # noqa: E402

n, d, b = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (8, 768, 65536)
dev = torch.device("cuda:0")
ctx = native.Context(0)
emb = torch.from_numpy(synth.make_embeddings(2048, n, d, seed=1)).to(dev).repeat((b + 2047) // 2048, 1, 1)[:b].contiguous()
thr = torch.full((b,), 0.5, dtype=torch.float32, device=dev)
ok = torch.empty(b, dtype=torch.int32, device=dev)
med = torch.empty(b, dtype=torch.int32, device=dev)
cos = torch.empty((b, n, n), dtype=torch.float32, device=dev)
s = torch.cuda.Stream(device=dev)
for _ in range(6):
    ctx.similarity_batch_dev(emb.data_ptr(), b, n, d, thr.data_ptr(), ok.data_ptr(), med.data_ptr(), cos.data_ptr(),
                             s.cuda_stream)
torch.cuda.synchronize()
print("done", int(ok.sum()))
ctx.close()
