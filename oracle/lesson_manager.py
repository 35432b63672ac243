"""Restatement of Quoracle.Agent.LessonManager -- reference lib/quoracle/agent/lesson_manager.ex:50-190.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Lessons are dicts {"type", "content", "confidence"}.
``embedding_fn`` is 1-arity here (the reference wraps 2-arity fns with cost opts, :110-123) and returns
("ok", {"embedding": vec}) | ("error", reason).
"""
from __future__ import annotations

from . import aggregator

DEFAULT_SIMILARITY_THRESHOLD = 0.90   # lesson_manager.ex:14
DEFAULT_MAX_LESSONS = 100             # lesson_manager.ex:15


def prune_lessons(lessons, max_count=None):
    """prune_lessons/2 -- lesson_manager.ex:138-147: stable sort by confidence desc, keep max_count."""
    max_count = DEFAULT_MAX_LESSONS if max_count is None else max_count
    if len(lessons) <= max_count:
        return lessons
    return sorted(lessons, key=lambda l: -l["confidence"])[:max_count]


def deduplicate_lesson(lesson, existing, opts=None):
    """deduplicate_lesson/3 -- lesson_manager.ex:95-130,150-173.

    -> ("merged", merged_lesson, old_content) | ("new", lesson)."""
    opts = opts or {}
    if not existing:
        return ("new", lesson)
    threshold = opts.get("similarity_threshold", DEFAULT_SIMILARITY_THRESHOLD)
    fn = opts["embedding_fn"]
    r = fn(lesson["content"])
    if not (isinstance(r, tuple) and r[0] == "ok"):
        return ("new", lesson)                       # embedding failed: add without dedup
    new_emb = r[1]["embedding"]
    for ex in existing:                              # reduce_while: stop at the FIRST match
        er = fn(ex["content"])
        if not (isinstance(er, tuple) and er[0] == "ok"):
            continue
        if aggregator.cosine_similarity(new_emb, er[1]["embedding"]) >= threshold:
            merged = dict(lesson, confidence=ex["confidence"] + 1)
            return ("merged", merged, ex["content"])
    return ("new", lesson)


def accumulate_lessons(existing, new_lessons, opts=None):
    """accumulate_lessons/3 -- lesson_manager.ex:50-75."""
    opts = opts or {}
    if not new_lessons:
        return ("ok", existing)
    acc, new_acc = list(existing), []
    for nl in new_lessons:
        r = deduplicate_lesson(nl, acc, opts)
        if r[0] == "merged":
            _, merged, old = r
            acc = [merged if l["content"] == old else l for l in acc]
        else:
            new_acc.append(r[1])
    return ("ok", prune_lessons(acc + new_acc, opts.get("max_lessons", DEFAULT_MAX_LESSONS)))
