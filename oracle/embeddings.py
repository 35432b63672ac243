"""Restatement of Quoracle.Models.Embeddings (cache, chunking, chunk mean) and TokenChunker.

TEST INFRASTRUCTURE (see oracle/__init__.py).
Reference: lib/quoracle/models/embeddings.ex:41-63,65-95,107-132,142-192,371-426 and
lib/quoracle/models/embeddings/token_chunker.ex:13-15,48-107.

The provider call (embeddings.ex:277-298, an HTTP request) is a plug-in here: ``options["embedding_fn"]``
(the reference's own test seam, embeddings.ex:213-217) or the module default set with
``set_provider``.  Token counting (tiktoken cl100k in the reference, token_manager.ex:19-24) is an
injectable ``token_counter``; the default counts whitespace-separated words -- "parity unpinned".
"""
from __future__ import annotations

import hashlib
import json
import time
from typing import Any, Callable, Optional

from . import numeric
from .consensus_rules import Accumulator

DEFAULT_CACHE_TTL = 3_600_000        # embeddings.ex:24 (ms)
MAX_CACHE_ENTRIES = 1000             # embeddings.ex:25
DEFAULT_EMBEDDING_TOKEN_LIMIT = 8191  # token_chunker.ex:13
TOKEN_SAFETY_MARGIN = 0.9            # token_chunker.ex:15

_provider: Optional[Callable[[str], Any]] = None
_default_cache: dict = {}


def set_provider(fn: Optional[Callable[[str], Any]]) -> None:
    global _provider
    _provider = fn


def default_token_counter(text: str) -> int:
    return len(text.split())


def effective_token_limit(model_limit: Optional[int] = None) -> int:
    """effective_token_limit/1 -- token_chunker.ex:48-53: trunc(limit * 0.9)."""
    return int((model_limit or DEFAULT_EMBEDDING_TOKEN_LIMIT) * TOKEN_SAFETY_MARGIN)


def chunk_text_by_tokens(text: str, max_tokens: int, token_counter=default_token_counter) -> list:
    """chunk_text_by_tokens/2 -- token_chunker.ex:62-107 (running per-word counts, +1 per separator)."""
    if token_counter(text) <= max_tokens:
        return [text]
    import re
    words = re.split(r"\s+", text)
    chunks, current, current_tokens = [], [], 0
    for word in words:
        wt = token_counter(word)
        sep = 0 if not current else 1
        new_tokens = current_tokens + wt + sep
        if new_tokens > max_tokens:
            if not current:
                chunks.append(word)                 # a single word over the limit is its own chunk
                current, current_tokens = [], 0
            else:
                chunks.append(" ".join(current))
                current, current_tokens = [word], wt
        else:
            current, current_tokens = current + [word], new_tokens
    if current:
        chunks.append(" ".join(current))
    return [c for c in chunks if c != ""]


average_embeddings = numeric.average_embeddings     # embeddings.ex:371-383


def _now_ms() -> int:
    return int(time.time() * 1000)


def _cache_key(text: str) -> bytes:
    return hashlib.sha256(text.encode("utf-8")).digest()      # :crypto.hash(:sha256, text), embeddings.ex:67


def _enforce_cache_limit(table: dict) -> None:
    """enforce_cache_limit/1 -- embeddings.ex:402-426: at >= 1000 entries drop the 100 oldest."""
    if len(table) >= MAX_CACHE_ENTRIES:
        oldest = sorted(table.items(), key=lambda kv: kv[1][1])[: MAX_CACHE_ENTRIES // 10]
        for k, _ in oldest:
            del table[k]


def get_embedding(text: Any, options: Optional[dict] = None):
    """get_embedding/2 -- embeddings.ex:41-63.

    -> ("ok", {"embedding", "cached", "chunks"}) | ("ok", result, acc) | ("error", atom)."""
    options = dict(options or {})
    if not isinstance(text, str):
        # Jason.encode!/1 (compact), embeddings.ex:46.  Jason walks a map in BEAM iteration order, which for the small
        # maps that occur as parameter values is ascending key order: two equal maps give the same text whatever order
        # their keys were written in
        text = json.dumps(text, separators=(",", ":"), sort_keys=True, ensure_ascii=False)
    if text == "":
        return ("error", "invalid_input")
    table = options.get("cache_table", _default_cache)
    ttl = options.get("cache_ttl", DEFAULT_CACHE_TTL)
    now_fn = options.get("now_ms", _now_ms)
    key = _cache_key(text)
    if ttl != 0 and key in table:                                  # check_cache/2, embeddings.ex:65-95
        emb, ts, entry_ttl = table[key]
        if now_fn() - ts < entry_ttl:
            return ("ok", {"embedding": emb, "cached": True, "chunks": 1})
        del table[key]
    r = _get_embedding_with_chunking(text, options)
    if r[0] != "ok":
        return r
    _, embedding, chunks = r
    if ttl > 0:                                                    # cache_embedding/3, embeddings.ex:385-400
        _enforce_cache_limit(table)
        table[key] = (embedding, now_fn(), ttl)
    result = {"embedding": embedding, "cached": False, "chunks": chunks}
    acc = options.get("cost_accumulator")
    if isinstance(acc, Accumulator):                               # embeddings.ex:116-127
        return ("ok", result, acc.add({"cost_type": "llm_embedding", "chunks": chunks}))
    return ("ok", result)


def _call_api(text: str, options: dict):
    """call_embedding_api/2 -- embeddings.ex:211-252 (only the injectable branch is restated)."""
    fn = options.get("embedding_fn", _provider)
    if fn is None:
        return ("error", "not_configured")
    r = fn(text)
    if isinstance(r, tuple) and len(r) == 2 and r[0] == "ok":
        return ("ok", r[1])
    return r


def _get_embedding_with_chunking(text: str, options: dict):
    """get_embedding_with_chunking/2 + process_chunks/2 -- embeddings.ex:142-192."""
    counter = options.get("token_counter", default_token_counter)
    limit = effective_token_limit(options.get("model_token_limit"))
    count = counter(text)
    if count > limit:
        return _process_chunks(chunk_text_by_tokens(text, limit, counter), options)
    r = _call_api(text, options)
    if r[0] == "ok":
        return ("ok", r[1], 1)
    if r == ("error", "context_length_exceeded"):                  # :159-164: retry at half the token count
        return _process_chunks(chunk_text_by_tokens(text, max(count // 2, 1), counter), options)
    return r


def _process_chunks(chunks: list, options: dict):
    results = []
    for c in chunks:                                               # serial, failed chunks dropped
        r = _call_api(c, options)
        if r[0] == "ok":
            results.append(r[1])
    if not results:
        return ("error", "all_chunks_failed")
    return ("ok", average_embeddings(results), len(results))


def sha_test_embedding(text: str):
    """test_embedding_fn/1 -- lib/quoracle/agent/consensus/per_model_query/helpers.ex:69-76:
    sha256(text) bytes / 255 -> 32-dim vector (the reference's reproducible deterministic embedding)."""
    return ("ok", {"embedding": [b / 255 for b in hashlib.sha256(text.encode("utf-8")).digest()]})
