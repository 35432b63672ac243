"""Small K1/K2/K3 run for compute-sanitizer (memcheck / racecheck / synccheck): a few decisions through
every kernel family, both CTA-group variants."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from quoracle_b200 import native, synth  # noqa: E402

w = synth.make_encoder(768, vocab=2048, seed=1)
with native.Context(0, w.table_bf16, w.w_bf16) as ctx:
    for variant in [(128, 2, 64, 1), (128, 1, 64, 0), (104, 2, 64, 1)]:
        ctx.set_variant(*variant)
        batch = synth.make_decisions(3, 8, 300, vocab=2048, seed=3, ragged=True)
        r = ctx.decide_batch(batch.token_ids, batch.lens, batch.thr, want_cos=True, want_emb=True)
        print(variant, r.ref_ok.tolist(), r.medoid.tolist(), float(np.abs(r.emb).sum()))
    emb = synth.make_embeddings(5, 24, 200, seed=2)
    print(ctx.similarity_batch(emb, np.full(5, 0.5, np.float32))[1].tolist())
    q = emb[:, 0]
    print(ctx.first_match_batch(q, emb, np.full(5, 0.9, np.float32))[0].tolist())
    print(ctx.chunk_mean(emb).shape)
print("sanitize run done")
