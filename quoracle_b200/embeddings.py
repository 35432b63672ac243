"""Host-side mirror of Quoracle.Models.Embeddings on top of the CUDA encoder.

Same contract as the reference's ``Embeddings.get_embedding/2`` (lib/quoracle/models/embeddings.ex:41-63):
non-binary input is JSON-encoded (:46), ``""`` -> ``("error", "invalid_input")`` (:49-50), sha256-keyed cache
with TTL 1 h / 1000 entries / evict the 100 oldest (:65-95,385-426), texts above the token limit are split
on whitespace into chunks (embeddings/token_chunker.ex:62-107), each chunk embedded and the vectors averaged
WITHOUT re-normalising (:371-383), and the 2-/3-tuple return shape with a cost accumulator (:116-127).
The provider HTTP call (:277-298) is replaced by ``qc_embed_batch`` (K1+K2); the chunk mean runs on the GPU
(``qc_chunk_mean``).  ``get_embeddings`` is the batched form the collector uses: every uncached text of a
whole consensus step goes to the GPU in ONE call.
"""
from __future__ import annotations

import hashlib
import json
import re
import time
from typing import Any, Optional, Sequence

import numpy as np

from . import tokenizer
from .native import QC_MAX_TOKENS, Context, QcError

DEFAULT_CACHE_TTL_MS = 3_600_000       # embeddings.ex:24
MAX_CACHE_ENTRIES = 1000               # embeddings.ex:25
DEFAULT_TOKEN_LIMIT = 8191             # token_chunker.ex:13
TOKEN_SAFETY_MARGIN = 0.9              # token_chunker.ex:15


class EmbeddingCache:
    """The ETS table owned by Quoracle.Models.EmbeddingCache (embedding_cache.ex:29-38), as a dict."""

    def __init__(self):
        self.table: dict[bytes, tuple[np.ndarray, int, int]] = {}

    def lookup(self, key: bytes, now_ms: int):
        hit = self.table.get(key)
        if hit is None:
            return None
        emb, ts, ttl = hit
        if now_ms - ts < ttl:
            return emb
        del self.table[key]
        return None

    def insert(self, key: bytes, emb: np.ndarray, now_ms: int, ttl: int) -> None:
        if len(self.table) >= MAX_CACHE_ENTRIES:
            for k, _ in sorted(self.table.items(), key=lambda kv: kv[1][1])[: MAX_CACHE_ENTRIES // 10]:
                del self.table[k]
        self.table[key] = (emb, now_ms, ttl)


def effective_token_limit(model_limit: Optional[int] = None) -> int:
    return int((model_limit or DEFAULT_TOKEN_LIMIT) * TOKEN_SAFETY_MARGIN)


def chunk_text_by_tokens(text: str, max_tokens: int) -> list[str]:
    """TokenChunker.chunk_text_by_tokens/2 with the hashing tokenizer's counts (one word = one token)."""
    if tokenizer.count_tokens(text) <= max_tokens:
        return [text]
    chunks, current, cur_tokens = [], [], 0
    for word in re.split(r"\s+", text):
        wt = tokenizer.count_tokens(word)
        new_tokens = cur_tokens + wt + (1 if current else 0)
        if new_tokens > max_tokens:
            if not current:
                chunks.append(word)
                current, cur_tokens = [], 0
            else:
                chunks.append(" ".join(current))
                current, cur_tokens = [word], wt
        else:
            current.append(word)
            cur_tokens = new_tokens
    if current:
        chunks.append(" ".join(current))
    return [c for c in chunks if c != ""]


class Embeddings:
    def __init__(self, ctx: Context, vocab: int = tokenizer.DEFAULT_VOCAB, cache: Optional[EmbeddingCache] = None,
                 model_token_limit: Optional[int] = None, now_ms=lambda: int(time.time() * 1000)):
        self.ctx = ctx
        self.vocab = vocab
        self.cache = cache or EmbeddingCache()
        self.limit = effective_token_limit(model_token_limit)
        self.now_ms = now_ms
        self.gpu_calls = 0

    @staticmethod
    def _stringify(value: Any) -> str:
        # Jason.encode!/1 (embeddings.ex:46): compact, keys in BEAM map order (ascending for small maps), UTF-8 as is
        return value if isinstance(value, str) else json.dumps(value, separators=(",", ":"), sort_keys=True, ensure_ascii=False)

    def get_embeddings(self, values: Sequence[Any], cache_ttl: int = DEFAULT_CACHE_TTL_MS) -> list:
        """Batched get_embedding: one GPU call for every uncached chunk of every value.

        -> per value ("ok", {"embedding": np.ndarray f32 [D], "cached": bool, "chunks": int}) | ("error", atom)."""
        texts = [self._stringify(v) for v in values]
        out: list = [None] * len(texts)
        now = self.now_ms()
        todo: dict[str, list[int]] = {}
        for i, text in enumerate(texts):
            if text == "":
                out[i] = ("error", "invalid_input")
                continue
            if cache_ttl != 0:
                hit = self.cache.lookup(hashlib.sha256(text.encode("utf-8")).digest(), now)
                if hit is not None:
                    out[i] = ("ok", {"embedding": hit, "cached": True, "chunks": 1})
                    continue
            todo.setdefault(text, []).append(i)            # in-batch dedup
        if not todo:
            return out
        uniq = list(todo.keys())
        chunk_lists = [self._chunks_for(t) for t in uniq]
        flat = [c for cl in chunk_lists for c in cl]
        tokens, offsets = tokenizer.pack_ragged(flat, self.vocab)     # texts differ in length: CSR, nothing padded
        try:
            emb = self.ctx.embed_ragged(tokens, offsets)
            self.gpu_calls += 1
        except QcError as e:
            for t in uniq:
                for i in todo[t]:
                    out[i] = ("error", e.atom)
            return out
        pos = 0
        for t, cl in zip(uniq, chunk_lists):
            if not cl:
                for i in todo[t]:
                    out[i] = ("error", "all_chunks_failed")
                continue
            rows = emb[pos:pos + len(cl)]
            pos += len(cl)
            vec = rows[0] if len(cl) == 1 else self.ctx.chunk_mean(rows[None])[0]   # unweighted, not re-normalised
            if cache_ttl > 0:
                self.cache.insert(hashlib.sha256(t.encode("utf-8")).digest(), vec, now, cache_ttl)
            for i in todo[t]:
                out[i] = ("ok", {"embedding": vec, "cached": False, "chunks": len(cl)})
        return out

    def _chunks_for(self, text: str) -> list:
        """get_embedding_with_chunking/2 (embeddings.ex:142-167) with the device in the provider's role: above the
        effective limit the text is chunked by the limit; a text within the limit that the encoder would still
        reject (more than QC_MAX_TOKENS token slots -- the reference's `:context_length_exceeded` answer) is
        re-chunked at HALF its own token count (:159-164); chunks that are still too long are the reference's
        failed chunks and are dropped (process_chunks/2, :170-192) -- an empty list means `:all_chunks_failed`."""
        count = tokenizer.count_tokens(text)
        if count > self.limit:
            chunks = chunk_text_by_tokens(text, self.limit)
        elif count > QC_MAX_TOKENS:
            chunks = chunk_text_by_tokens(text, max(count // 2, 1))
        else:
            return [text]
        return [c for c in chunks if tokenizer.count_tokens(c) <= QC_MAX_TOKENS]

    def get_embedding(self, value: Any, options: Optional[dict] = None):
        """Embeddings.get_embedding/2: 2-tuple, or 3-tuple with the accumulator when one is passed and the
        result was not cached (local embeddings cost nothing: the entry records chunks only)."""
        options = options or {}
        r = self.get_embeddings([value], cache_ttl=options.get("cache_ttl", DEFAULT_CACHE_TTL_MS))[0]
        acc = options.get("cost_accumulator")
        if r[0] == "ok" and acc is not None and not r[1]["cached"]:
            return ("ok", r[1], acc.add({"cost_type": "llm_embedding", "cost_usd": None, "chunks": r[1]["chunks"]}))
        return r

    def embedding_fn(self):
        """A 2-arity closure with the `:embedding_fn` contract (consensus_rules.ex:325-327)."""
        return lambda value, opts: self.get_embedding(value, opts)
