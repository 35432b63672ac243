"""Hashing tokenizer for the local encoder: token id = FNV-1a-32(word bytes) mod vocab.

The reference tokenizes only to COUNT tokens (tiktoken cl100k, lib/quoracle/agent/token_manager.ex:19-24);
a local encoder needs ids.  FNV-1a over whitespace-separated words is trivially reproducible from any
host language (the Elixir glue in elixir/quoracle_native.ex computes the same function), needs no
vocabulary file, and one word == one token keeps the reference's chunking rule (words + separators,
embeddings/token_chunker.ex:62-107) meaningful.
"""
from __future__ import annotations

import numpy as np

FNV_OFFSET = 0x811C9DC5
FNV_PRIME = 0x01000193
DEFAULT_VOCAB = 32768


def fnv1a32(data: bytes) -> int:
    h = FNV_OFFSET
    for b in data:
        h ^= b
        h = (h * FNV_PRIME) & 0xFFFFFFFF
    return h


def tokenize(text: str, vocab: int = DEFAULT_VOCAB) -> list[int]:
    return [fnv1a32(w.encode("utf-8")) % vocab for w in text.split()]


def count_tokens(text: str) -> int:
    return len(text.split())


def pack(texts: list[str], vocab: int = DEFAULT_VOCAB, max_tokens: int | None = None):
    """-> (token_ids [R, T] int32 padded with -1, lens [R] int32); T = longest row (>= 1)."""
    rows = [tokenize(t, vocab) for t in texts]
    if max_tokens is not None:
        rows = [r[:max_tokens] for r in rows]
    t = max(1, max((len(r) for r in rows), default=1))
    ids = np.full((len(rows), t), -1, dtype=np.int32)
    lens = np.zeros(len(rows), dtype=np.int32)
    for i, r in enumerate(rows):
        ids[i, : len(r)] = r
        lens[i] = len(r)
    return ids, lens


def pack_ragged(texts: list[str], vocab: int = DEFAULT_VOCAB, max_tokens: int | None = None):
    """-> (tokens [sum len] int32, offsets [R + 1] int32): the CSR form qc_embed_ragged / qc_decide_ragged take --
    no padding is ever materialised, on the host or on the wire."""
    rows = [tokenize(t, vocab) for t in texts]
    if max_tokens is not None:
        rows = [r[:max_tokens] for r in rows]
    offsets = np.zeros(len(rows) + 1, dtype=np.int32)
    np.cumsum([len(r) for r in rows], out=offsets[1:])
    tokens = np.fromiter((t for r in rows for t in r), dtype=np.int32, count=int(offsets[-1]))
    return tokens, offsets
