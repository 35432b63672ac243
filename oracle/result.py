"""Restatement of Quoracle.Consensus.Result (+ Result.Scoring): winner selection, per-parameter merge
with the mode fallback, confidence.

TEST INFRASTRUCTURE (see oracle/__init__.py).
Reference: lib/quoracle/consensus/result.ex:30-42,58-130,133-248,261-372 and
lib/quoracle/consensus/result/scoring.ex:15-102.
"""
from __future__ import annotations

from typing import Optional

from . import consensus_rules as rules
from . import schema_tables, terms
from .consensus_rules import Accumulator


# ---- scoring.ex ---------------------------------------------------------------------------------------

def wait_score(value):
    """wait_score/1 -- scoring.ex:33-37: {true_count, finite_sum}; lower = more conservative."""
    if value is True:
        return (0, 0)
    if value is None:
        return (0, 1)
    if value is False or (isinstance(value, int) and not isinstance(value, bool) and value == 0):
        return (1, 0)
    if isinstance(value, int) and not isinstance(value, bool) and value > 0:
        return (0, 1 + value)
    raise ValueError("wait_score: no function clause")


def cluster_wait_score(cluster):
    """cluster_wait_score/1 -- scoring.ex:44-54: component-wise sum over the cluster's responses."""
    if "actions" not in cluster:
        return (0, 0)
    tc = fs = 0
    for a in cluster["actions"]:
        t, f = wait_score(a.get("wait"))
        tc, fs = tc + t, fs + f
    return (tc, fs)


def calculate_cluster_priority(cluster):
    """calculate_cluster_priority/1 -- scoring.ex:57-102: batches take the MAX priority of their members."""
    rep = cluster["representative"]
    action = rep["action"]
    if action in ("batch_async", "batch_sync"):
        actions = (rep.get("params") or {}).get("actions")
        if isinstance(actions, list) and actions:
            ps = []
            for a in actions:
                p = schema_tables.get_action_priority(a.get("action"))
                ps.append(999 if p is None else p)
            return max(ps)
        return 999
    p = schema_tables.get_action_priority(action)
    return 999 if p is None else p


def break_tie(tied_clusters):
    """break_tie/1 -- scoring.ex:15-26: stable sort by {priority, wait score}; the first wins."""
    if len(tied_clusters) == 1:
        return tied_clusters[0]
    return sorted(tied_clusters, key=lambda c: (calculate_cluster_priority(c), cluster_wait_score(c)))[0]


# ---- result.ex ----------------------------------------------------------------------------------------

def find_winner(clusters, total_count):
    """find_winner/2 -- result.ex:290-308."""
    for c in clusters:
        if c["count"] > total_count / 2:
            return ("majority", c)
    max_count = max(c["count"] for c in clusters)
    tied = [c for c in clusters if c["count"] == max_count]
    return ("plurality", break_tie(tied) if len(tied) > 1 else tied[0])


def calculate_confidence(cluster, total_count, round_num, opts: Optional[dict] = None) -> float:
    """calculate_confidence/4 -- result.ex:261-286."""
    max_rounds = (opts or {}).get("max_refinement_rounds", 4)
    frac = cluster["count"] / total_count
    bonus = 0.15 if frac > 0.8 else 0.10 if frac > 0.6 else 0.05 if frac > 0.5 else 0.0
    penalty = (round_num - max_rounds) * 0.1 if round_num > max_rounds else 0.0
    return min(max(frac + bonus - penalty, 0.1), 1.0)


def mode_value(values):
    """mode_value/1 -- result.ex:367-372."""
    return terms.frequencies_max(values)


def merge_params_by_rules(actions, schema, cost_opts: dict):
    """merge_params_by_rules/3 -- result.ex:311-365: schema order, nil values dropped, rule error -> mode."""
    all_params = [a.get("params") or {} for a in actions]
    cost_acc = cost_opts.get("cost_accumulator")
    merged = {}
    for param in schema["required"] + schema["optional"]:
        values = [p.get(param) for p in all_params]
        values = [v for v in values if v is not None]
        if not values:
            continue
        rule = schema["rules"].get(param, ("mode_selection",))
        cur = dict(cost_opts)
        if cost_acc is not None:
            cur["cost_accumulator"] = cost_acc
        r = rules.apply_rule(rule, values, cur)
        if isinstance(r[0], tuple):                       # ({:ok,v}|{:error,_}, acc)
            (tag, payload), acc = r
            cost_acc = acc
            merged[param] = payload if tag == "ok" else mode_value(values)
        elif r[0] == "ok":
            merged[param] = r[1]
        else:
            merged[param] = mode_value(values)
    return merged, cost_acc


def _merge_wait_parameter(values):
    """merge_wait_parameter/1 -- result.ex:240-248."""
    r = rules.merge_params("wait", "wait", values)
    return r[1] if r[0] == "ok" else values[0]


def _finish(result, actions):
    wait_values = [a.get("wait") for a in actions if a.get("wait") is not None]
    result["wait"] = _merge_wait_parameter(wait_values) if wait_values else False   # result.ex:113-121
    return result


def merge_cluster_params(cluster, cost_opts: Optional[dict] = None):
    """merge_cluster_params/2 -- result.ex:58-130 -> (action_map, accumulator | None) or ("error", reason)."""
    cost_opts = dict(cost_opts or {})
    actions = cluster["actions"]
    rep = cluster["representative"]
    if rep["action"] == "batch_sync":                                         # :58-84
        sequences = [((a.get("params") or {}).get("actions") or []) for a in actions]
        initial_acc = cost_opts.get("cost_accumulator")
        r = rules.apply_rule(("batch_sequence_merge",), sequences, cost_opts)
        if r[0] == "ok":
            merged_actions = r[1]
        elif r == ("error", "no_consensus"):
            merged_actions = _merge_sequences_with_mode(sequences)
        else:
            return r
        reasoning = next((a.get("reasoning") for a in actions if a.get("reasoning") not in ("", None)), "")
        return _finish({"action": "batch_sync", "params": {"actions": merged_actions}, "reasoning": reasoning},
                       actions), initial_acc
    schema = schema_tables.get_schema(rep["action"])
    if schema is None:                                                        # :125-129
        return rep, cost_opts.get("cost_accumulator")
    merged, final_acc = merge_params_by_rules(actions, schema, cost_opts)
    reasoning = next((a.get("reasoning") for a in actions if a.get("reasoning") != ""), "")
    return _finish({"action": rep["action"], "params": merged, "reasoning": reasoning}, actions), final_acc


def _merge_sequences_with_mode(sequences):
    """merge_sequences_with_mode/1 -- result.ex:163-228."""
    max_len = max((len(s) for s in sequences), default=0)
    out = []
    for idx in range(max_len):
        at = [s[idx] for s in sequences if idx < len(s) and s[idx] is not None]
        if not at:
            continue
        if len(at) == 1:
            out.append({"action": at[0].get("action"), "params": dict(at[0].get("params") or {})})
            continue
        action_type = mode_value([a.get("action") for a in at])
        matching = [a for a in at if a.get("action") == action_type]
        all_params = [a.get("params") or {} for a in matching]
        keys = terms.uniq([k for p in all_params for k in p.keys()])
        merged = {}
        for k in keys:
            vals = [p.get(k) for p in all_params if p.get(k) is not None]
            if vals:
                merged[k] = mode_value(vals)
        out.append({"action": action_type, "params": merged})
    return out


def format_result(clusters, total_count, round_num, cost_opts: Optional[dict] = None):
    """format_result/4 -- result.ex:30-47 -> (:consensus | :forced_decision, action, opts)."""
    cost_opts = dict(cost_opts or {})
    kind, cluster = find_winner(clusters, total_count)
    merged = merge_cluster_params(cluster, cost_opts)
    action, accumulator = merged
    confidence = calculate_confidence(cluster, total_count, round_num, cost_opts)
    opts = {"confidence": confidence}
    if accumulator is not None:
        opts["accumulator"] = accumulator
    return ("consensus" if kind == "majority" else "forced_decision", action, opts)


__all__ = ["Accumulator", "wait_score", "cluster_wait_score", "calculate_cluster_priority", "break_tie",
           "find_winner", "calculate_confidence", "mode_value", "merge_params_by_rules", "merge_cluster_params",
           "format_result"]
