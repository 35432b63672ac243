"""numpy f64 restatement of THIS BUILD's local encoder (the reference defines none).

    h = L2normalise( mean_{t < len} GELU_tanh( E[tok_t] . W ) )

E: bf16 [V, D], W: bf16 [D, D] (in x out).  Stands in for the provider call of
Embeddings.get_embedding/2 (reference lib/quoracle/models/embeddings.ex:277-298); the surrounding
cache / chunk / mean logic is in oracle/embeddings.py.  Embedding VALUES are "parity unpinned"
against the reference (see oracle/__init__.py); this file is what the CUDA encoder is checked against.
"""
from __future__ import annotations

import numpy as np


def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even f32 -> bf16 bit patterns (uint16)."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    rounding = ((u >> np.uint32(16)) & np.uint32(1)) + np.uint32(0x7FFF)
    return ((u + rounding) >> np.uint32(16)).astype(np.uint16)


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (np.asarray(b, dtype=np.uint16).astype(np.uint32) << np.uint32(16)).view(np.float32)


def gelu_tanh(x: np.ndarray) -> np.ndarray:
    return 0.5 * x * (1.0 + np.tanh(0.7978845608028654 * (x + 0.044715 * x * x * x)))


def encode(table_bf16: np.ndarray, w_bf16: np.ndarray, token_ids: np.ndarray, lens: np.ndarray,
           rows_per_call: int = 32) -> np.ndarray:
    """token_ids [R, T] int32, lens [R] -> normalised embeddings [R, D] f64."""
    table = np.asarray(table_bf16, dtype=np.uint16)
    vocab, dim = table.shape
    w = bf16_bits_to_f32(np.asarray(w_bf16, dtype=np.uint16).reshape(dim, dim)).astype(np.float64)
    token_ids = np.asarray(token_ids, dtype=np.int64)
    lens = np.asarray(lens, dtype=np.int64)
    r_total, t_slots = token_ids.shape
    out = np.zeros((r_total, dim), dtype=np.float64)
    for r0 in range(0, r_total, rows_per_call):
        r1 = min(r_total, r0 + rows_per_call)
        ids = token_ids[r0:r1]
        ln = lens[r0:r1]
        valid = (np.arange(t_slots)[None, :] < ln[:, None]) & (ids >= 0) & (ids < vocab)
        flat = np.where(valid, ids, 0).reshape(-1)
        x = bf16_bits_to_f32(table[flat]).astype(np.float64)
        x[~valid.reshape(-1)] = 0.0                      # padding contributes GELU(0) = 0
        g = gelu_tanh(x @ w).reshape(r1 - r0, t_slots, dim)
        pooled = g.sum(axis=1)
        pooled = np.where(ln[:, None] > 0, pooled / np.maximum(ln, 1)[:, None], 0.0)
        nrm = np.sqrt((pooled * pooled).sum(axis=1))
        out[r0:r1] = np.where(nrm[:, None] > 0, pooled / np.where(nrm > 0, nrm, 1.0)[:, None], 0.0)
    return out
