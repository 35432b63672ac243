"""Small K1/K2/K3 run for compute-sanitizer (memcheck / racecheck / synccheck): a few decisions through
every kernel family, both CTA-group variants."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from quoracle_b200 import native, synth  # noqa: E402

w = synth.make_encoder(768, vocab=2048, seed=1)
with native.Context(0, w.table_bf16, w.w_bf16) as ctx:
    for variant in [(128, 2, 64, 1), (128, 1, 64, 0), (104, 2, 64, 1), (128, 2, 64, 1, 2), (128, 2, 64, 1, 4)]:
        ctx.set_variant(*variant[:4])
        ctx.set_multicast(variant[4] if len(variant) > 4 else 1)      # round 2: W-tile TMA multicast clusters of 4 / 8
        batch = synth.make_decisions(3, 8, 300, vocab=2048, seed=3, ragged=True)
        r = ctx.decide_batch(batch.token_ids, batch.lens, batch.thr, want_cos=True, want_emb=True)
        print(variant, r.ref_ok.tolist(), r.medoid.tolist(), float(np.abs(r.emb).sum()))
    emb = synth.make_embeddings(5, 24, 200, seed=2)
    print(ctx.similarity_batch(emb, np.full(5, 0.5, np.float32))[1].tolist())
    q = emb[:, 0]
    print(ctx.first_match_batch(q, emb, np.full(5, 0.9, np.float32))[0].tolist())
    print(ctx.chunk_mean(emb).shape)
    # packed token stream (host entry point, forced), CSR entry points, tcgen05 K3 on a many-tile batch
    ctx.set_variant(0, 0, 0, 0)
    ctx.set_packing(2)
    batch = synth.make_decisions(5, 8, 200, vocab=2048, seed=4, ragged=True)
    r = ctx.decide_batch(batch.token_ids, batch.lens, batch.thr, want_cos=True, want_emb=True)
    print("packed", r.medoid.tolist(), float(np.abs(r.emb).sum()))
    ctx.set_packing(0)
    lens = np.array([0, 1, 33, 200, 7, 64, 65, 128], np.int32)
    off = np.zeros(9, np.int32)
    np.cumsum(lens, out=off[1:])
    toks = np.random.default_rng(1).integers(0, 2048, size=int(off[-1]), dtype=np.int32)
    print("ragged", float(np.abs(ctx.embed_ragged(toks, off)).sum()),
          ctx.decide_ragged(toks, off, 8, np.array([0.9], np.float32)).medoid.tolist())
    for n, d in ((9, 100), (16, 384), (43, 96), (64, 128)):
        e = synth.make_embeddings(40, n, d, seed=n)
        print("tc", n, d, ctx.similarity_batch(e, np.full(40, 0.5, np.float32))[1][:4].tolist())
w = synth.make_encoder(1536, vocab=2048, seed=2)
with native.Context(0, w.table_bf16, w.w_bf16) as ctx:            # 56-token blocks: 8-token chunks
    ctx.set_packing(2)
    batch = synth.make_decisions(2, 4, 150, vocab=2048, seed=5, ragged=True)
    print("packed 1536", ctx.decide_batch(batch.token_ids, batch.lens, batch.thr).medoid.tolist())
    ctx.set_packing(0)
    for tb in (128, 96):                                            # round 2: K-split walk (rows held one K half at a time)
        ctx.set_variant(tb, 2, 64, 1)
        batch = synth.make_decisions(2, 4, 300, vocab=2048, seed=6, ragged=True)
        print("k-split", tb, ctx.decide_batch(batch.token_ids, batch.lens, batch.thr).medoid.tolist())
w = synth.make_encoder(384, vocab=2048, seed=3)
with native.Context(0, w.table_bf16, w.w_bf16) as ctx:            # round 2: 256-token tile (the dim-384 default)
    batch = synth.make_decisions(3, 5, 300, vocab=2048, seed=7, ragged=True)
    print("384/256", ctx.decide_batch(batch.token_ids, batch.lens, batch.thr).medoid.tolist())
    cache = ctx.lesson_cache(16, 384)
    rows = np.random.default_rng(0).standard_normal((6, 384)).astype(np.float32)
    cache.append(rows)
    cache.gather([4, 2, 0])
    print("lesson cache", cache.match(rows[2], 0.99))
print("sanitize run done")
