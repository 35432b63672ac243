"""Restatement of Quoracle.Consensus.Aggregator (clustering + cosine entry points).

TEST INFRASTRUCTURE (see oracle/__init__.py).
Reference: lib/quoracle/consensus/aggregator.ex:23-103,241-362,366-480 and
lib/quoracle/consensus/aggregator/param_matching.ex:9-91.
Responses are dicts {"action": str, "params": dict, "reasoning": str, "wait": ...}.
"""
from __future__ import annotations

import json
import re
from typing import Any, Optional

from . import numeric, schema_tables, terms
from .consensus_rules import Accumulator

cosine_similarity = numeric.cosine_similarity_aggregator   # aggregator.ex:335-350


def cluster_responses(responses: list) -> list:
    """cluster_responses/1 -- aggregator.ex:23-35.

    group_by fingerprint, then a STABLE sort by count desc.  The order among equal counts is BEAM map order over
    the fingerprint keys `{action, signature}`: for the <= 32 clusters a pool can produce that is ascending Erlang
    term order (flat maps keep their keys sorted), restated here as (action atom by name, term order of the
    signature); above 32 keys BEAM iterates in hash order and the tie order is unpinned."""
    if not responses:
        return []
    groups: dict = {}
    for r in responses:
        fp = action_fingerprint(r)
        groups.setdefault(_fp_key(fp), (fp, []))[1].append(r)
    ordered = sorted(groups.values(), key=lambda g: (str(g[0][0]).encode("utf-8"), terms.term_order(g[0][1])))
    clusters = [{"count": len(actions), "actions": actions, "representative": actions[0], "fingerprint": fp}
                for fp, actions in ordered]
    return sorted(clusters, key=lambda c: -c["count"])       # Python's sort is stable, like Enum.sort_by


def _fp_key(fp):
    action, sig = fp
    return (str(action), json.dumps(_jsonable(sig), sort_keys=True))


def _jsonable(v):
    if isinstance(v, dict):
        return {str(k): _jsonable(x) for k, x in v.items()}
    if isinstance(v, (list, tuple)):
        return [_jsonable(x) for x in v]
    return v


def find_majority_cluster(clusters: list, total_count: int, round_: int = 2):
    """find_majority_cluster/3 -- aggregator.ex:48-62: round 1 unanimous, rounds >= 2 strictly > 50%."""
    for c in clusters:
        if (round_ == 1 and c["count"] == total_count) or (round_ != 1 and c["count"] > total_count / 2):
            return ("majority", c)
    return ("no_majority", clusters)


def action_fingerprint(response: dict):
    """action_fingerprint/1 -- aggregator.ex:72-103."""
    action = response["action"]
    if action == "batch_async":
        return ("batch_async", sorted(_extract_action_type(a) for a in _get_batch_actions(response.get("params"))))
    if action == "batch_sync":
        return ("batch_sync", [_extract_action_type(a) for a in _get_batch_actions(response.get("params"))])
    schema = schema_tables.get_schema(action)
    if schema is None:
        return (action, "invalid")
    return (action, _create_action_signature(response, schema))


def _get_batch_actions(params):
    if isinstance(params, dict) and isinstance(params.get("actions"), list):
        return params["actions"]
    return []


def _extract_action_type(spec):
    if isinstance(spec, dict) and isinstance(spec.get("action"), str):
        return spec["action"]
    return "unknown"


def _create_action_signature(response, schema):
    """create_action_signature/2 -- aggregator.ex:366-387."""
    params = response.get("params") or {}
    sig = {}
    for param in schema["required"] + schema["optional"]:
        value = params.get(param)
        if value is None:
            continue
        sig[param] = _normalize_param_for_signature(value, schema["rules"].get(param))
    return sig


def _normalize_param_for_signature(value, rule):
    """normalize_param_for_signature/2 -- aggregator.ex:389-421."""
    name = rule[0] if rule else None
    if name == "exact_match":
        return value
    if name == "semantic_similarity":
        threshold = rule[1] if rule[1] else 0.9
        return normalize_semantic_string(value, threshold)
    if name == "mode_selection":
        return "_mode_mergeable"
    if name == "union_merge":
        return sorted(value, key=terms.term_order) if isinstance(value, list) else value
    if name == "structural_merge":
        return _deep_sort_map(value) if isinstance(value, dict) else value
    if name == "percentile":
        return "_percentile_mergeable"
    return value


_WS = re.compile(r"\s+", re.UNICODE)
_NOT_WORD_SPACE = re.compile(r"[^\w\s]", re.ASCII)   # Elixir ~r/[^\w\s]/ without the u flag: ASCII classes


def normalize_semantic_string(s, threshold):
    """normalize_semantic_string/2 + extract_key_terms/1 -- aggregator.ex:423-454."""
    if not isinstance(s, str):
        return s
    s = s.lower()
    if threshold < 0.95:
        s = _NOT_WORD_SPACE.sub("", s)
    s = re.sub(r"\s+", " ", s).strip()
    words = [w for w in s.split() if len(w) > 3]        # String.length/1 counts graphemes; len() code points
    words.sort(key=lambda w: w.encode("utf-8"))
    return "_".join(words[:5])


def _deep_sort_map(v):
    if isinstance(v, dict):
        return {k: _deep_sort_map(x) for k, x in sorted(v.items(), key=lambda kv: terms.term_order(kv[0]))}
    return v


# ---- actions_match? / param matching ------------------------------------------------------------------

def actions_match(r1: dict, r2: dict) -> bool:
    """actions_match?/2 -- aggregator.ex:109-122; ParamMatching.params_match?/3 -- param_matching.ex:9-91."""
    if r1["action"] != r2["action"]:
        return False
    schema = schema_tables.get_schema(r1["action"])
    if schema is None:
        return False
    p1, p2 = r1.get("params") or {}, r2.get("params") or {}
    for param, rule in schema["rules"].items():
        if not _param_values_match(p1.get(param), p2.get(param), rule):
            return False
    return True


def _param_values_match(v1, v2, rule):
    if v1 is None and v2 is None:
        return True
    if v1 is None or v2 is None:
        return False
    name = rule[0]
    if name == "exact_match":
        return terms.term_equal(v1, v2)
    if name == "semantic_similarity":
        threshold = rule[1] or 0.9
        if not (isinstance(v1, str) and isinstance(v2, str)):
            return False
        n1, n2 = _pm_normalize(v1), _pm_normalize(v2)
        if n1 == n2:
            return True
        t1, t2 = set(n1.split("_")), set(n2.split("_"))
        union = len(t1 | t2)
        if union == 0:
            return True
        return len(t1 & t2) / union >= threshold * 0.8
    if name in ("mode_selection", "percentile"):
        return True
    return terms.term_equal(v1, v2)


def _pm_normalize(s):
    """param_matching.ex:82-90."""
    s = re.sub(r"[^a-z0-9\s]", "", s.lower())
    words = sorted(s.split(), key=lambda w: w.encode("utf-8"))
    return "_".join(words[:5])


# ---- semantic similarity via embeddings ----------------------------------------------------------------

def calculate_semantic_similarity(text1, text2, opts: Optional[dict] = None):
    """calculate_semantic_similarity/3 -- aggregator.ex:249-289.

    opts["embedding_fn"]: 1-arity ``fn(text)`` -> ("ok", vec) | ("ok", {"embedding": vec}), or 2-arity
    ``fn(text, opts_map)`` which may also return ("ok", result, acc).  Arity is given explicitly via
    opts["embedding_fn_arity"] (default 1) since Python callables do not carry one.
    Returns a float, or (float, acc) iff an accumulator was passed."""
    opts = opts or {}
    raw = opts.get("embedding_fn")
    arity = opts.get("embedding_fn_arity", 1)
    acc0 = opts.get("cost_accumulator")
    cost_opts = {k: opts[k] for k in ("agent_id", "task_id", "pubsub", "cost_accumulator") if k in opts}
    tracks = True
    if raw is not None and arity == 2:
        fn = lambda text, m: raw(text, m)               # noqa: E731
    elif raw is not None:
        fn = lambda text, m: raw(text)                  # noqa: E731
        tracks = False
    else:
        from . import embeddings
        fn = lambda text, m: embeddings.get_embedding(text, m)   # noqa: E731

    def call(text, m, cur):
        r = fn(text, m)
        if not tracks:
            return r, cur
        if isinstance(r, tuple) and len(r) == 3 and r[0] == "ok" and isinstance(r[2], Accumulator):
            return ("ok", r[1]), r[2]
        if isinstance(r, tuple) and len(r) == 2 and r[0] == "ok":
            return r, cur
        return r, cur

    r1, acc1 = call(text1, dict(cost_opts), acc0)
    m2 = dict(cost_opts)
    if acc1 is not None:
        m2["cost_accumulator"] = acc1
    r2, acc2 = call(text2, m2, acc1)
    e1, e2 = _unwrap(r1), _unwrap(r2)
    sim = cosine_similarity(e1, e2) if (e1 is not None and e2 is not None) else 0.0
    return (sim, acc2) if isinstance(acc2, Accumulator) else sim


def _unwrap(r):
    if isinstance(r, tuple) and len(r) == 2 and r[0] == "ok":
        if isinstance(r[1], dict) and "embedding" in r[1]:
            return r[1]["embedding"]
        if isinstance(r[1], list):
            return r[1]
    return None


def semantic_similarity_with_embeddings(text1, text2, threshold, opts=None) -> bool:
    """aggregator.ex:359-362 (as the reference: a {float, acc} tuple would compare > any float; callers pass no acc)."""
    sim = calculate_semantic_similarity(text1, text2, opts)
    return (sim[0] if isinstance(sim, tuple) else sim) >= threshold


def get_cached_embedding(text, opts=None):
    """get_cached_embedding/2 -- aggregator.ex:315-328."""
    opts = opts or {}
    fn = opts.get("embedding_fn")
    if fn:
        r = fn(text)
        if isinstance(r, tuple) and r[0] == "ok":
            return ("ok", {"embedding": r[1], "cached": False, "chunks": 1})
        return r
    from . import embeddings
    return embeddings.get_embedding(text, opts)
