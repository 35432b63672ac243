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


class LessonStore:
    """One agent's lesson list with its embeddings RESIDENT on the GPU (qc_lesson_cache_*).

    Same results as ``accumulate_lessons`` above (and as lesson_manager.ex:50-173), but a dedup step embeds only the
    NEW lesson and uploads one query row: the rows of the existing lessons sit in HBM in list order (rows of lessons
    whose embedding failed are simply absent -- the reference skips those comparisons, :167-169), a merge replaces
    the matched row with the new lesson's (the merged lesson keeps the NEW content, :159-161), new lessons are
    appended after the whole call (they are not compared with each other, :55-69), pruning reorders the rows with
    one device-side gather."""

    def __init__(self, emb: Embeddings, lessons=(), threshold=DEFAULT_SIMILARITY_THRESHOLD,
                 max_lessons=DEFAULT_MAX_LESSONS, capacity=None):
        self.emb = emb
        self.threshold = float(threshold)
        self.max_lessons = int(max_lessons)
        # room for a full list plus one call's worth of new lessons before the prune
        self.cache = emb.ctx.lesson_cache(capacity or 2 * self.max_lessons + 64, emb.ctx.dim)
        self.lessons: list = []
        self.row_of: list[int] = []          # device row of lesson i, -1 when its embedding failed
        self.uploaded_rows = 0               # rows sent host -> device so far (the cost the cache removes)
        if lessons:
            self._load(list(lessons))

    def _vectors(self, contents):
        return [r[1]["embedding"] if r[0] == "ok" else None for r in self.emb.get_embeddings(contents)]

    def _load(self, lessons):
        vecs = self._vectors([l["content"] for l in lessons])
        self.lessons, self.row_of = list(lessons), []
        rows = []
        for v in vecs:
            self.row_of.append(len(rows) if v is not None else -1)
            if v is not None:
                rows.append(v)
        self.cache.gather([])
        if rows:
            self.cache.put(0, np.stack(rows).astype(np.float32))
            self.uploaded_rows += len(rows)

    def deduplicate(self, lesson):
        """-> ("merged", merged_lesson, old_content, list_index, query_row) | ("new", lesson, query_row | None)."""
        (vec,) = self._vectors([lesson["content"]])
        if vec is None or not self.lessons:
            return ("new", lesson, vec)
        self.uploaded_rows += 1
        idx, _ = self.cache.match(vec, self.threshold)
        if idx < 0:
            return ("new", lesson, vec)
        li = self.row_of.index(idx)
        hit = self.lessons[li]
        return ("merged", dict(lesson, confidence=hit["confidence"] + 1), hit["content"], li, vec)

    def accumulate(self, new_lessons):
        if not new_lessons:
            return ("ok", list(self.lessons))
        fresh = []
        for nl in new_lessons:
            r = self.deduplicate(nl)
            if r[0] == "new":
                fresh.append((r[1], r[2]))
                continue
            _, merged, old_content, _, vec = r
            hit_rows = [i for i, l in enumerate(self.lessons) if l["content"] == old_content]
            if any(self.row_of[i] < 0 for i in hit_rows):
                # a lesson with the same content but no row (its embedding failed earlier) now gets one: the row order
                # changes, so rebuild (rare; the vectors come out of the embedding cache)
                self.lessons = [merged if l["content"] == old_content else l for l in self.lessons]
                self._load(self.lessons)
                continue
            for i in hit_rows:                                   # replace_lesson_by_content: every lesson with that content
                self.lessons[i] = merged
                self.cache.put(self.row_of[i], vec[None])
                self.uploaded_rows += 1
        base = len(self.cache)
        rows = [v for _, v in fresh if v is not None]
        k = 0
        for lesson, v in fresh:
            self.lessons.append(lesson)
            self.row_of.append(base + k if v is not None else -1)
            k += v is not None
        if rows:
            self.cache.append(np.stack(rows).astype(np.float32))
            self.uploaded_rows += len(rows)
        if len(self.lessons) > self.max_lessons:                 # prune_lessons: stable sort by confidence, keep the top
            order = sorted(range(len(self.lessons)), key=lambda i: -self.lessons[i]["confidence"])[: self.max_lessons]
            src = [self.row_of[i] for i in order if self.row_of[i] >= 0]
            self.cache.gather(src)
            new_row, k = [], 0
            for i in order:
                new_row.append(k if self.row_of[i] >= 0 else -1)
                k += self.row_of[i] >= 0
            self.lessons = [self.lessons[i] for i in order]
            self.row_of = new_row
        return ("ok", list(self.lessons))

    def close(self):
        self.cache.close()
