"""Host-side mirror of the `{:semantic_similarity, threshold: t}` rule on the GPU path.

``apply_rule`` follows ConsensusRules.apply_rule/3 for this rule (lib/quoracle/actions/consensus_rules.ex:18-31,
249-299): identical values short-circuit with no embedding work (:251-256); otherwise all N values are
embedded (here: ONE batched GPU call instead of N sequential HTTP calls, :302-318), any failure yields
``("error", "no_consensus")`` so the caller's mode fallback fires (result.ex:352-358), and value_0 wins iff
every cos(e_0, e_i) >= threshold (:277-286) -- evaluated by K3 on the device.  ``decide_many`` is the batched
form for a whole agent tree: P parameter merges -> one qc_similarity_batch call.
"""
from __future__ import annotations

from typing import Any, Optional, Sequence

import numpy as np

from .embeddings import Embeddings
from .terms import uniq as _uniq

DEFAULT_THRESHOLD = 0.9   # consensus_rules.ex:19


def decide_many(emb: Embeddings, merges: Sequence[tuple[Sequence[Any], float]]):
    """merges: [(values, threshold)] -> [("ok", value_0) | ("error", "no_consensus") | ("error", "no_values")]
    plus, per merge, the medoid index (north-star vote) or None when no GPU comparison was needed."""
    results: list = [None] * len(merges)
    medoids: list = [None] * len(merges)
    pending = []
    for i, (values, _thr) in enumerate(merges):
        values = list(values)
        if not values:
            results[i] = ("error", "no_values")
            continue
        u = _uniq(values)
        if len(u) == 1:
            results[i] = ("ok", u[0])                       # no embedding call at all
            continue
        pending.append(i)
    if not pending:
        return results, medoids
    flat = [v for i in pending for v in merges[i][0]]
    embs = emb.get_embeddings(flat)
    pos = 0
    by_n: dict[int, list] = {}
    for i in pending:
        n = len(merges[i][0])
        rows = embs[pos:pos + n]
        pos += n
        if any(r[0] != "ok" for r in rows):
            results[i] = ("error", "no_consensus")          # :embedding_failed -> :no_consensus (:294-297, :28-29)
            continue
        by_n.setdefault(n, []).append((i, np.stack([r[1]["embedding"] for r in rows])))
    for n, items in by_n.items():                            # one K3 launch per pool size
        x = np.stack([m for _, m in items]).astype(np.float32)
        thr = np.array([merges[i][1] for i, _ in items], dtype=np.float32)
        ref_ok, medoid, _ = emb.ctx.similarity_batch(x, thr, want_cos=False)
        for (i, _), ok, med in zip(items, ref_ok, medoid):
            results[i] = ("ok", merges[i][0][0]) if ok else ("error", "no_consensus")
            medoids[i] = int(med)
    return results, medoids


def apply_rule(emb: Embeddings, values: Sequence[Any], threshold: float = DEFAULT_THRESHOLD,
               cost_accumulator: Optional[Any] = None):
    """One `{:semantic_similarity, threshold: t}` merge; with an accumulator the result is wrapped
    ``(result, acc)`` exactly as the reference does (consensus_rules.ex:24-29)."""
    (res,), _ = decide_many(emb, [(values, threshold)])
    if cost_accumulator is not None and len(_uniq(list(values))) > 1:
        for _ in values:
            cost_accumulator = cost_accumulator.add({"cost_type": "llm_embedding", "cost_usd": None})
        return (res, cost_accumulator)
    return res
