"""Synthetic inputs for the consensus hot path.

Everything is generated host-side with ``numpy.random.default_rng(seed)``; seed = 0xC0FFEE + config id.
There is no dataset or checkpoint in this environment: token ids, encoder weights and embedding
tensors are all synthetic, and every report says so.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

BASE_SEED = 0xC0FFEE
VOCAB = 32768
RESAMPLE_PROBS = (0.0, 0.05, 0.3, 1.0)   # identical / near / far / unrelated pool members


def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even float32 -> bf16 bit patterns (uint16)."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    rounding = ((u >> np.uint32(16)) & np.uint32(1)) + np.uint32(0x7FFF)
    return ((u + rounding) >> np.uint32(16)).astype(np.uint16)


@dataclass
class EncoderWeights:
    table_bf16: np.ndarray   # [vocab, dim] uint16
    w_bf16: np.ndarray       # [dim, dim] uint16 (in x out)
    seed: int


def make_encoder(dim: int, vocab: int = VOCAB, seed: int = BASE_SEED) -> EncoderWeights:
    """E ~ N(0,1), W ~ N(0, 1/dim), both rounded to bf16 (the build's encoder definition)."""
    rng = np.random.default_rng(seed)
    table = rng.standard_normal((vocab, dim), dtype=np.float32)
    w = rng.standard_normal((dim, dim), dtype=np.float32) / np.float32(np.sqrt(dim))
    return EncoderWeights(f32_to_bf16_bits(table), f32_to_bf16_bits(w), seed)


@dataclass
class DecisionBatch:
    token_ids: np.ndarray    # [B, N, T] int32, -1 beyond lens
    lens: np.ndarray         # [B, N] int32
    thr: np.ndarray          # [B] float32
    seed: int


def make_decisions(b: int, n: int, t: int, vocab: int = VOCAB, seed: int = BASE_SEED, ragged: bool = False,
                   thresholds=(0.85,)) -> DecisionBatch:
    """Per decision: a base token sequence; candidate k copies it, resampling each token with
    probability RESAMPLE_PROBS[k mod 4].  ``ragged`` draws len ~ U[T/2, T] and pads with -1."""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, vocab, size=(b, 1, t), dtype=np.int32)
    fresh = rng.integers(0, vocab, size=(b, n, t), dtype=np.int32)
    # identical / near / far / unrelated, nudged by 0.03 per wrap so no two pool members are exact
    # duplicates (exact duplicates tie the medoid vote to the last ulp, which no float path can pin)
    p = np.array([min(1.0, RESAMPLE_PROBS[k % len(RESAMPLE_PROBS)] + 0.03 * (k // len(RESAMPLE_PROBS)))
                  for k in range(n)], dtype=np.float32)
    # shuffle which pool member is identical / near / far / unrelated so index 0 is not special
    perm = np.stack([rng.permutation(n) for _ in range(b)])
    p_b = p[perm]                                            # [B, N]
    take_fresh = rng.random((b, n, t), dtype=np.float32) < p_b[:, :, None]
    tok = np.where(take_fresh, fresh, np.broadcast_to(base, (b, n, t))).astype(np.int32)
    if ragged:
        lens = rng.integers(max(1, t // 2), t + 1, size=(b, n), dtype=np.int32)
        tok[np.arange(t)[None, None, :] >= lens[:, :, None]] = -1
    else:
        lens = np.full((b, n), t, dtype=np.int32)
    thr = np.asarray(thresholds, dtype=np.float32)[rng.integers(0, len(thresholds), size=b)]
    return DecisionBatch(np.ascontiguousarray(tok), lens, thr.astype(np.float32), seed)


def make_embeddings(b: int, n: int, d: int, seed: int = BASE_SEED) -> np.ndarray:
    """C5 sweep inputs: normalised N(0,1) rows plus a per-decision shared component with mixing
    weight U[0,1] per row (so both verdicts and all medoid positions occur)."""
    rng = np.random.default_rng(seed)
    own = rng.standard_normal((b, n, d), dtype=np.float32)
    shared = rng.standard_normal((b, 1, d), dtype=np.float32)
    mix = rng.random((b, n, 1), dtype=np.float32)
    x = (1.0 - mix) * own + mix * shared
    x /= np.linalg.norm(x, axis=2, keepdims=True)
    return np.ascontiguousarray(x.astype(np.float32))
