"""Numeric core of the oracle: cosine, the row-0 threshold rule, the medoid vote, chunk mean.

Scalar functions use Python floats (IEEE f64) with strictly left-to-right accumulation, which is
what BEAM does on boxed floats; the ``*_batch`` helpers are numpy f64 restatements of the same
formulas for sizes where a Python loop would take minutes (they differ from the scalar versions
only in summation order, i.e. by a few ulp -- irrelevant at the 1e-3 tolerance the path states).
"""
from __future__ import annotations

import math
from typing import Sequence

import numpy as np


def cosine_similarity_rules(vec1: Sequence[float], vec2: Sequence[float]) -> float:
    """ConsensusRules.cosine_similarity/2 -- reference lib/quoracle/actions/consensus_rules.ex:189-210.

    Raises on length mismatch (:189-191); one zip+reduce pass over {dot, m1, m2}; 0.0 for a zero norm.
    """
    if len(vec1) != len(vec2):
        raise ValueError("Vectors must have the same length")
    dot = 0.0
    m1 = 0.0
    m2 = 0.0
    for v1, v2 in zip(vec1, vec2):
        dot = dot + v1 * v2
        m1 = m1 + v1 * v1
        m2 = m2 + v2 * v2
    mag1 = math.sqrt(m1)
    mag2 = math.sqrt(m2)
    if mag1 == 0.0 or mag2 == 0.0:
        return 0.0
    return dot / (mag1 * mag2)


def cosine_similarity_aggregator(vec1: Sequence[float], vec2: Sequence[float]) -> float:
    """Aggregator.cosine_similarity/2 -- reference lib/quoracle/consensus/aggregator.ex:335-350.

    Three separate passes; 0.0 on a zero norm; 0.0 (not an exception) on length mismatch (:350).
    """
    if len(vec1) != len(vec2):
        return 0.0
    dot = 0.0
    for a, b in zip(vec1, vec2):
        dot = dot + a * b
    s1 = 0.0
    for x in vec1:
        s1 = s1 + x * x
    s2 = 0.0
    for x in vec2:
        s2 = s2 + x * x
    mag1 = math.sqrt(s1)
    mag2 = math.sqrt(s2)
    if mag1 == 0.0 or mag2 == 0.0:
        return 0.0
    return dot / (mag1 * mag2)


def reference_rule(embeddings: Sequence[Sequence[float]], threshold: float) -> bool:
    """All i >= 1: cos(e_0, e_i) >= threshold -- reference consensus_rules.ex:277-286 (Enum.all?)."""
    first = embeddings[0]
    for e in embeddings[1:]:
        if not (cosine_similarity_rules(first, e) >= threshold):
            return False
    return True


def cosine_matrix(embeddings: Sequence[Sequence[float]]) -> list[list[float]]:
    """Full N x N matrix with the reference cosine; the north-star's similarity tile."""
    n = len(embeddings)
    m = [[0.0] * n for _ in range(n)]
    for i in range(n):
        for j in range(i, n):
            c = cosine_similarity_rules(embeddings[i], embeddings[j])
            m[i][j] = c
            m[j][i] = c
    return m


def medoid_vote(cos: Sequence[Sequence[float]]) -> int:
    """argmax_i mean_j cos[i][j] (j ascending, diagonal included); the first maximum wins."""
    n = len(cos)
    best, best_mean = 0, 0.0
    for i in range(n):
        s = 0.0
        for j in range(n):
            s = s + cos[i][j]
        m = s / n
        if i == 0 or m > best_mean:
            best, best_mean = i, m
    return best


def average_embeddings(embeddings: Sequence[Sequence[float]]) -> list[float]:
    """Embeddings.average_embeddings/1 -- reference lib/quoracle/models/embeddings.ex:371-383.

    Element-wise: integer-0 seeded left-to-right sum over the chunks, divided by the chunk count.
    The result is not re-normalised.
    """
    dimension = len(embeddings[0])
    out = []
    for i in range(dimension):
        s = 0
        for e in embeddings:
            s = s + e[i]
        out.append(s / len(embeddings))
    return out


def first_match(query: Sequence[float], rows: Sequence[Sequence[float]], threshold: float) -> int:
    """Arithmetic of LessonManager.find_similar_lesson/5 -- reference lesson_manager.ex:150-173."""
    for m, row in enumerate(rows):
        if cosine_similarity_aggregator(query, row) >= threshold:
            return m
    return -1


# ---- numpy f64 batch restatements ---------------------------------------------------------------


def cosine_matrix_batch(emb: np.ndarray) -> np.ndarray:
    """emb [B, N, D] (any float dtype) -> cos [B, N, N] f64, zero-norm rows/cols give 0.0."""
    e = np.asarray(emb, dtype=np.float64)
    gram = np.einsum("bid,bjd->bij", e, e)
    nrm = np.sqrt(np.einsum("bid,bid->bi", e, e))
    denom = nrm[:, :, None] * nrm[:, None, :]
    with np.errstate(divide="ignore", invalid="ignore"):
        cos = np.where(denom == 0.0, 0.0, gram / np.where(denom == 0.0, 1.0, denom))
    return cos


def decide_batch(emb: np.ndarray, thr: np.ndarray):
    """-> (ref_ok [B] int32, medoid [B] int32, cos [B,N,N] f64, margins dict).

    margins: 'thr' = min_i |cos(0,i) - thr| over i >= 1 (inf for N == 1), 'vote' = gap between the two
    largest row means (inf for N == 1) -- used by the harness to reject inputs whose verdict a
    float32 path could legitimately flip (hard parts).
    """
    cos = cosine_matrix_batch(emb)
    b, n, _ = cos.shape
    thr = np.asarray(thr, dtype=np.float64)
    if n > 1:
        row0 = cos[:, 0, 1:]
        ref_ok = np.all(row0 >= thr[:, None], axis=1).astype(np.int32)
        thr_margin = np.min(np.abs(row0 - thr[:, None]), axis=1)
    else:
        ref_ok = np.ones(b, dtype=np.int32)
        thr_margin = np.full(b, np.inf)
    # row means with j ascending, diagonal included (cumulative left-to-right sum)
    means = np.zeros((b, n))
    for j in range(n):
        means = means + cos[:, :, j]
    means = means / n
    medoid = np.argmax(means, axis=1).astype(np.int32)  # first maximum
    if n > 1:
        srt = np.sort(means, axis=1)
        vote_margin = srt[:, -1] - srt[:, -2]
    else:
        vote_margin = np.full(b, np.inf)
    return ref_ok, medoid, cos, {"thr": thr_margin, "vote": vote_margin}
