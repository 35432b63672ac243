"""Restatement of Quoracle.Actions.ConsensusRules -- reference lib/quoracle/actions/consensus_rules.ex.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Return shapes mirror the Elixir ones:
``("ok", v)`` / ``("error", reason)``, and ``(("ok", v), acc)`` / ``(("error", r), acc)`` when a cost
accumulator was threaded through embedding calls (consensus_rules.ex:24-29).
"""
from __future__ import annotations

from typing import Any, Callable, Optional

from . import numeric, schema_tables, terms


class Accumulator:
    """Quoracle.Costs.Accumulator (lib/quoracle/costs/accumulator.ex:22-46): an immutable entry list."""

    def __init__(self, entries=None):
        self.entries = list(entries or [])

    def add(self, entry) -> "Accumulator":
        return Accumulator(self.entries + [entry])

    def count(self) -> int:
        return len(self.entries)


EmbeddingFn = Callable[[Any, dict], Any]
_default_embedding: Optional[EmbeddingFn] = None


def set_default_embedding(fn: Optional[EmbeddingFn]) -> None:
    """Stands in for ModelQuery.get_embedding/2 (consensus_rules.ex:321-323) when no embedding_fn is given."""
    global _default_embedding
    _default_embedding = fn


cosine_similarity = numeric.cosine_similarity_rules   # consensus_rules.ex:189-210


# ---- apply_rule -----------------------------------------------------------------------------------

def apply_rule(rule, values: list, cost_opts: Optional[dict] = None):
    """apply_rule/2 and /3 -- consensus_rules.ex:18-163."""
    name = rule[0] if isinstance(rule, tuple) else rule
    if cost_opts is not None:
        if name == "semantic_similarity":                                  # :18-31
            opts = rule[1] if len(rule) > 1 and isinstance(rule[1], dict) else {"threshold": rule[1]} if len(rule) > 1 else {}
            threshold = opts.get("threshold", schema_tables.DEFAULT_SEMANTIC_THRESHOLD)
            embedding_fn = opts.get("embedding_fn") or cost_opts.get("embedding_fn")
            res = _semantic_similarity_check(values, threshold, opts.get("plug"), cost_opts, embedding_fn)
            return res
        if name == "batch_sequence_merge":                                 # :33-44
            return _batch_sequence_merge(values, cost_opts)
        return apply_rule(rule, values)                                    # :46
    # ---- 2-arity --------------------------------------------------------------------------------
    if name == "exact_match":                                              # :49-57
        if values == []:
            return ("error", "no_values")
        if len(values) == 1:
            return ("ok", values[0])
        u = terms.uniq(values)
        return ("ok", u[0]) if len(u) == 1 else ("error", "no_consensus")
    if name == "semantic_similarity":                                      # :59-68
        opts = rule[1] if len(rule) > 1 and isinstance(rule[1], dict) else {"threshold": rule[1]} if len(rule) > 1 else {}
        threshold = opts.get("threshold", schema_tables.DEFAULT_SEMANTIC_THRESHOLD)
        return _semantic_similarity_check(values, threshold, opts.get("plug"), {}, None)
    if name == "mode_selection":                                           # :70-76
        if values == []:
            return ("error", "no_values")
        return ("ok", terms.frequencies_max(values))
    if name == "union_merge":                                              # :78-81
        flat = _flatten(values)
        return ("ok", terms.uniq(flat))
    if name == "structural_merge":                                         # :83-91
        merged: Any = {}
        for v in values:
            merged = _deep_merge(merged, v)
        return ("ok", merged)
    if name == "percentile" and isinstance(values, list) and values != []:  # :93-105
        numeric_values = [v for v in values if terms.is_number(v)]
        if not numeric_values:
            return apply_rule(("mode_selection",), values)
        s = sorted(numeric_values)
        index = rule[1] / 100.0 * (len(s) - 1)                             # :342-344
        return ("ok", _percentile_value(s, index))
    if name == "batch_sequence_merge":                                     # :107-112
        return _batch_sequence_merge(values, {})
    if name == "wait_parameter":                                           # :114-160
        return _wait_parameter(values)
    if values == []:                                                       # :162
        return ("error", "no_values")
    return ("error", "unknown_rule")                                       # :163


def merge_params(action: str, param: str, values: list):
    """merge_params/3 -- consensus_rules.ex:168-182."""
    if param == "wait":
        return apply_rule(("wait_parameter",), values)
    schema = schema_tables.get_schema(action)
    if schema is None:
        return ("error", "unknown_action")
    rule = schema["rules"].get(param)
    if rule is None:
        return ("error", "unknown_param")
    return apply_rule(rule, values)


# ---- semantic similarity (the hot rule) -------------------------------------------------------------

def _semantic_similarity_check(values, threshold, plug, cost_opts, embedding_fn):
    """semantic_similarity_check/5 + the apply_rule result mapping -- consensus_rules.ex:24-30,225-247."""
    if values == []:
        return ("error", "no_consensus")
    res = _use_embeddings_similarity(values, threshold, plug, cost_opts, embedding_fn)
    if isinstance(res[0], tuple):                       # ({:ok,v}|{:error,r}, acc)
        (tag, payload), acc = res
        return (("ok", payload), acc) if tag == "ok" else (("error", "no_consensus"), acc)
    tag, payload = res
    return ("ok", payload) if tag == "ok" else ("error", "no_consensus")


def _use_embeddings_similarity(values, threshold, plug, cost_opts, embedding_fn):
    """use_embeddings_similarity/5 -- consensus_rules.ex:249-299."""
    u = terms.uniq(values)
    if not u:
        return ("error", "no_values")
    if len(u) == 1:
        return ("ok", u[0])                              # no embedding call at all (:251-256)
    embed_opts = dict(cost_opts)
    if plug:
        embed_opts = {"plug": plug, **embed_opts}
    initial_acc = cost_opts.get("cost_accumulator")
    pairs, final_acc = _get_embeddings_for_values(values, embed_opts, embedding_fn, initial_acc)
    if len(pairs) != len(values):
        return ("error", "embedding_failed")             # :294-297 (accumulator dropped, as in the reference)
    if len(pairs) == 1 and len(values) > 1:              # unreachable given the length check; kept for fidelity
        result = ("error", "no_consensus")
    else:
        first_value, first_embedding = pairs[0]
        ok = True
        for _, emb in pairs[1:]:                         # Enum.all?/2 short-circuits
            if not (cosine_similarity(first_embedding, emb) >= threshold):
                ok = False
                break
        result = ("ok", first_value) if ok else ("error", "no_consensus")
    if isinstance(final_acc, Accumulator):
        return (result, final_acc)
    return result


def _get_embeddings_for_values(values, embed_opts, embedding_fn, initial_acc):
    """get_embeddings_for_values/4 -- consensus_rules.ex:302-318: sequential, pool order, acc threaded."""
    results, acc = [], initial_acc
    for value in values:
        current = dict(embed_opts)
        if acc is not None:
            current["cost_accumulator"] = acc
        r = _call_embedding(value, current, embedding_fn)
        if isinstance(r, tuple) and len(r) == 3 and r[0] == "ok" and isinstance(r[1], dict) and "embedding" in r[1]:
            results.append((value, r[1]["embedding"]))
            acc = r[2]
        elif isinstance(r, tuple) and len(r) == 2 and r[0] == "ok" and isinstance(r[1], dict) and "embedding" in r[1]:
            results.append((value, r[1]["embedding"]))
        # anything else: dropped (:313-314)
    return results, acc


def _call_embedding(value, opts, embedding_fn):
    """call_embedding/3 -- consensus_rules.ex:321-327."""
    if embedding_fn is None:
        if _default_embedding is None:
            return ("error", "not_configured")
        return _default_embedding(value, opts)
    return embedding_fn(value, opts)


# ---- small rules ------------------------------------------------------------------------------------

def _flatten(values):
    out = []
    for v in values:
        if isinstance(v, list):
            out.extend(_flatten(v))
        else:
            out.append(v)
    return out


def _deep_merge(m1, m2):
    """deep_merge/2 -- consensus_rules.ex:329-340."""
    if isinstance(m1, dict) and isinstance(m2, dict):
        out = dict(m1)
        for k, v2 in m2.items():
            if k in out and isinstance(out[k], dict) and isinstance(v2, dict):
                out[k] = _deep_merge(out[k], v2)
            else:
                out[k] = v2
        return out
    if isinstance(m2, dict):
        return m2
    if isinstance(m1, dict):
        return m1
    raise ValueError("deep_merge: no clause matches two non-maps")   # FunctionClauseError in the reference


def _percentile_value(sorted_vals, index):
    """calculate_percentile_value/2 -- consensus_rules.ex:346-358 (linear interpolation, round/1)."""
    lower = int(index)                                   # trunc/1
    upper = min(lower + 1, len(sorted_vals) - 1)
    fraction = index - lower
    lo, hi = sorted_vals[lower], sorted_vals[upper]
    return terms.erl_round(lo + (hi - lo) * fraction)


def _median(sorted_list):
    """calculate_median/1 -- consensus_rules.ex:360-373 (integer div for the even case)."""
    count = len(sorted_list)
    mid = count // 2
    if count % 2 == 0:
        return terms.erl_div(sorted_list[mid - 1] + sorted_list[mid], 2)
    return sorted_list[mid]


def _wait_parameter(values):
    """apply_rule(:wait_parameter, values) -- consensus_rules.ex:114-160."""
    if values == []:
        return ("error", "no_values")
    booleans = [v for v in values if isinstance(v, bool)]
    integers = [v for v in values if not isinstance(v, bool)]
    if not integers and booleans and all(b is False for b in booleans):
        return ("ok", False)
    if not integers and booleans and all(b is True for b in booleans):
        return ("ok", True)
    if not integers and len(booleans) >= 3 and any(b is True for b in booleans):
        return ("ok", True)
    if not booleans and integers:
        return ("ok", _median(sorted(integers)))
    max_int = 30 if not integers else max(integers)
    converted = []
    for v in values:
        if v is False:
            converted.append(0)
        elif v is True:
            converted.append(max_int)
        elif isinstance(v, int):
            converted.append(v)
        else:
            raise ValueError("wait_parameter: no clause for non-integer value")   # FunctionClauseError
    return ("ok", _median(sorted(converted)))


# ---- batch_sync per-position merge ----------------------------------------------------------------

def _batch_sequence_merge(sequences, cost_opts):
    """:batch_sequence_merge -- consensus_rules.ex:33-44,375-471."""
    if sequences == []:
        return ("ok", [])
    if len(sequences) == 1:
        return ("ok", sequences[0])
    if len(set(len(s) for s in sequences)) > 1:
        return ("error", "sequence_length_mismatch")
    merged = []
    for position in zip(*sequences):                     # Enum.zip_with(sequences, & &1)
        r = _merge_position(list(position), cost_opts)
        if r[0] != "ok":
            return r
        merged.append(r[1])
    return ("ok", merged)


def _normalize_action_keys(action):
    a = action.get("action")
    params = action.get("params")
    return {"action": a, "params": params if params is not None else {}}


def _merge_position(actions, cost_opts):
    normalized = [_normalize_action_keys(a) for a in actions]
    types = terms.uniq([a["action"] for a in normalized])
    if len(types) != 1:
        return ("error", "sequence_mismatch")
    action_type = types[0]
    schema = schema_tables.get_schema(action_type)
    if schema is None:
        return ("error", "unknown_action")
    params_list = [a["params"] for a in normalized]
    all_keys = terms.uniq([k for p in params_list for k in p.keys()])
    merged = {}
    for key in all_keys:
        values = [p.get(key) for p in params_list]
        values = [v for v in values if v is not None]
        if not values:
            continue
        rule = schema["rules"].get(key, ("exact_match",))     # default differs from result.ex:334 (Appendix A.4)
        r = apply_rule(rule, values, cost_opts)
        if isinstance(r[0], tuple):
            # (({:ok,v}|{:error,r}), acc): the reference has no matching clause here;
            # oracle choice: treat as the inner result and note the divergence.
            r = r[0]
        if r[0] == "ok":
            merged[key] = r[1]
        else:
            return ("error", "no_consensus")
    return ("ok", {"action": action_type, "params": merged})
