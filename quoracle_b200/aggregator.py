"""Host-side mirror of Aggregator.calculate_semantic_similarity/3 (lib/quoracle/consensus/aggregator.ex:249-289) on the
GPU path: the reference embeds the two texts one after the other and takes their cosine on the CPU (0.0 when either
embedding is unavailable, :285-288 keeps the accumulator shape); here any number of text pairs costs ONE embedding
call (cache and in-batch dedup in front of it) and ONE two-row K3 launch.
"""
from __future__ import annotations

from typing import Any, Optional, Sequence

import numpy as np

from .embeddings import Embeddings


def semantic_similarity_many(emb: Embeddings, pairs: Sequence[tuple[Any, Any]]) -> list[float]:
    """[(text1, text2)] -> [cosine]; a pair with a failed embedding scores 0.0, as in the reference."""
    if not pairs:
        return []
    flat = [t for p in pairs for t in p]
    rs = emb.get_embeddings(flat)
    out = [0.0] * len(pairs)
    good = [i for i in range(len(pairs)) if rs[2 * i][0] == "ok" and rs[2 * i + 1][0] == "ok"]
    if not good:
        return out
    x = np.stack([np.stack([rs[2 * i][1]["embedding"], rs[2 * i + 1][1]["embedding"]]) for i in good]).astype(np.float32)
    _, _, cos = emb.ctx.similarity_batch(x, np.zeros(len(good), np.float32), want_cos=True)
    for k, i in enumerate(good):
        out[i] = float(cos[k, 0, 1])
    return out


def calculate_semantic_similarity(emb: Embeddings, text1: Any, text2: Any, cost_accumulator: Optional[Any] = None):
    """One pair; with an accumulator the result is `(float, acc)` (aggregator.ex:285-288), one entry per embedding that
    was actually computed (cached texts cost nothing, embeddings.ex:116-127)."""
    before = emb.gpu_calls
    rs = emb.get_embeddings([text1, text2])
    sim = 0.0
    if rs[0][0] == "ok" and rs[1][0] == "ok":
        x = np.stack([rs[0][1]["embedding"], rs[1][1]["embedding"]])[None].astype(np.float32)
        sim = float(emb.ctx.similarity_batch(x, np.zeros(1, np.float32), want_cos=True)[2][0, 0, 1])
    if cost_accumulator is None:
        return sim
    acc = cost_accumulator
    if emb.gpu_calls != before:
        for r in rs:
            if r[0] == "ok" and not r[1]["cached"]:
                acc = acc.add({"cost_type": "llm_embedding", "cost_usd": None, "chunks": r[1]["chunks"]})
    return (sim, acc)


def semantic_similarity_with_embeddings(emb: Embeddings, text1: Any, text2: Any, threshold: float) -> bool:
    """semantic_similarity_with_embeddings/4 (aggregator.ex:352-362): similarity >= threshold."""
    return calculate_semantic_similarity(emb, text1, text2) >= threshold
