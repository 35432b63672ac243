"""A stand-in for quoracle_b200.native.Context used by the CPU tests of the host logic: same method
signatures, arithmetic delegated to the oracle (tests may do that; the product never does)."""
import numpy as np

from oracle import encoder as oenc
from oracle import numeric


class FakeContext:
    def __init__(self, table_bf16, w_bf16):
        self.table, self.w = table_bf16, w_bf16
        self.dim = table_bf16.shape[1]
        self.calls = {"embed": 0, "similarity": 0, "first_match": 0, "chunk_mean": 0}

    def embed_batch(self, token_ids, lens):
        self.calls["embed"] += 1
        return oenc.encode(self.table, self.w, token_ids, lens).astype(np.float32)

    def embed_ragged(self, tokens, offsets):
        self.calls["embed"] += 1
        lens = np.diff(offsets).astype(np.int32)
        t = max(1, int(lens.max(initial=1)))
        ids = np.full((len(lens), t), -1, np.int32)
        for r, (o, n) in enumerate(zip(offsets[:-1], lens)):
            ids[r, :n] = tokens[o:o + n]
        return oenc.encode(self.table, self.w, ids, lens).astype(np.float32)

    def similarity_batch(self, emb, thr, want_cos=True):
        self.calls["similarity"] += 1
        ok, med, cos, _ = numeric.decide_batch(emb, np.asarray(thr, dtype=np.float64))
        return ok, med, cos.astype(np.float32) if want_cos else None

    def first_match_batch(self, query, rows, thr, want_cos=True):
        self.calls["first_match"] += 1
        idx = np.array([numeric.first_match(q.astype(np.float64).tolist(), r.astype(np.float64).tolist(), float(t))
                        for q, r, t in zip(query, rows, thr)], dtype=np.int32)
        return idx, None

    def chunk_mean(self, chunks):
        self.calls["chunk_mean"] += 1
        return np.stack([np.array(numeric.average_embeddings(g.astype(np.float64).tolist()), dtype=np.float32)
                         for g in chunks])
