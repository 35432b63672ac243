"""Host-side mirror of LessonManager dedup (lib/quoracle/agent/lesson_manager.ex:50-173) on the GPU path.

The reference re-embeds every existing lesson for every new lesson and walks them one by one; here the new
lesson and all existing ones are embedded in one batched call (cache hits cost nothing) and the first match
>= threshold comes from ``qc_first_match_batch``.  Semantics kept: first match wins, merged lesson keeps the
NEW content with confidence = old + 1, an embedding failure of the new lesson adds it without dedup, failed
existing lessons are skipped, pruning keeps the highest confidences.
"""
from __future__ import annotations

import numpy as np

from .embeddings import Embeddings

DEFAULT_SIMILARITY_THRESHOLD = 0.90
DEFAULT_MAX_LESSONS = 100


def prune_lessons(lessons, max_count=DEFAULT_MAX_LESSONS):
    if len(lessons) <= max_count:
        return lessons
    return sorted(lessons, key=lambda l: -l["confidence"])[:max_count]


def deduplicate_lesson(emb: Embeddings, lesson, existing, threshold=DEFAULT_SIMILARITY_THRESHOLD):
    if not existing:
        return ("new", lesson)
    rs = emb.get_embeddings([lesson["content"]] + [e["content"] for e in existing])
    if rs[0][0] != "ok":
        return ("new", lesson)
    ok_idx = [i for i, r in enumerate(rs[1:]) if r[0] == "ok"]
    if not ok_idx:
        return ("new", lesson)
    q = rs[0][1]["embedding"][None].astype(np.float32)
    rows = np.stack([rs[1 + i][1]["embedding"] for i in ok_idx])[None].astype(np.float32)
    idx, _ = emb.ctx.first_match_batch(q, rows, np.array([threshold], np.float32), want_cos=False)
    if idx[0] < 0:
        return ("new", lesson)
    hit = existing[ok_idx[int(idx[0])]]
    return ("merged", dict(lesson, confidence=hit["confidence"] + 1), hit["content"])


def accumulate_lessons(emb: Embeddings, existing, new_lessons, threshold=DEFAULT_SIMILARITY_THRESHOLD,
                       max_lessons=DEFAULT_MAX_LESSONS):
    if not new_lessons:
        return ("ok", existing)
    acc, fresh = list(existing), []
    for nl in new_lessons:
        r = deduplicate_lesson(emb, nl, acc, threshold)
        if r[0] == "merged":
            acc = [r[1] if l["content"] == r[2] else l for l in acc]
        else:
            fresh.append(r[1])
    return ("ok", prune_lessons(acc + fresh, max_lessons))
