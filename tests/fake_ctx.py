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

    def lesson_cache(self, capacity=128, dim=None):
        return FakeLessonCache(self, capacity, dim or self.dim)


class FakeLessonCache:
    """Host stand-in for native.LessonCache with the same methods (rows kept in a numpy array, oracle arithmetic)."""

    def __init__(self, ctx, capacity, dim):
        self.ctx, self.capacity, self.dim = ctx, capacity, dim
        self.rows = np.zeros((0, dim), np.float32)

    def __len__(self):
        return self.rows.shape[0]

    def put(self, index, rows):
        rows = np.asarray(rows, np.float32).reshape(-1, self.dim)
        assert 0 <= index <= len(self) and index + rows.shape[0] <= self.capacity
        keep = max(len(self), index + rows.shape[0])
        out = np.zeros((keep, self.dim), np.float32)
        out[: len(self)] = self.rows
        out[index:index + rows.shape[0]] = rows
        self.rows = out

    def append(self, rows):
        self.put(len(self), rows)

    def gather(self, src_index):
        idx = np.asarray(src_index, np.int64).reshape(-1)
        self.rows = self.rows[idx] if idx.size else np.zeros((0, self.dim), np.float32)

    def match(self, query, thr, want_cos=False):
        self.ctx.calls["first_match"] += 1
        if len(self) == 0:
            return -1, None
        i = numeric.first_match(np.asarray(query, np.float64).tolist(), self.rows.astype(np.float64).tolist(), float(thr))
        return int(i), None

    def close(self):
        pass
