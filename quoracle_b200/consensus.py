"""Batched per-agent decide step: cluster the N proposals of every pending decision, pick the winner,
merge its parameters -- with ALL semantic-similarity merges of the whole agent tree evaluated in one GPU
flush -- and score confidence.

This is the host-side mirror of ``Result.format_result/4`` after ``Aggregator.cluster_responses/1``
(reference lib/quoracle/consensus/result.ex:30-42, aggregator.ex:23-35), i.e. what
``Consensus.execute_consensus_process/3`` (lib/quoracle/agent/consensus.ex:295-330) does once the LLM
responses are in, restructured so the only heavy arithmetic (embedding + cosine of the semantic parameters)
is gathered across agents.  String/integer rules (fingerprints, mode, percentile, wait, tie-break,
confidence) run on the host exactly as in the reference; the rule/priority tables are data
(``data/action_schemas.json``, transcribed from the reference's schemas).
"""
from __future__ import annotations

import json
import math
import re
from collections import OrderedDict
from pathlib import Path
from typing import Any, Optional, Sequence

from .collector import DecisionCollector

_DATA = json.loads((Path(__file__).parent / "data" / "action_schemas.json").read_text())
SCHEMAS = _DATA["schemas"]
PRIORITIES = _DATA["priorities"]
DEFAULT_THRESHOLD = _DATA["default_semantic_threshold"]


from .terms import key as _key, mode_value, order as _order, uniq as _uniq  # noqa: E402


# ---- non-semantic rules (consensus_rules.ex:49-163) --------------------------------------------------------
def _median(sorted_ints):
    n = len(sorted_ints)
    if n % 2:
        return sorted_ints[n // 2]
    s = sorted_ints[n // 2 - 1] + sorted_ints[n // 2]
    return int(math.trunc(s / 2)) if abs(s) < 2**52 else (abs(s) // 2) * (1 if s >= 0 else -1)


def wait_rule(values):
    bools = [v for v in values if isinstance(v, bool)]
    ints = [v for v in values if not isinstance(v, bool)]
    if not ints and bools:
        if not any(bools):
            return False
        if all(bools) or len(bools) >= 3:
            return True
    if not bools and ints:
        return _median(sorted(ints))
    top = max(ints) if ints else 30
    return _median(sorted(0 if v is False else top if v is True else v for v in values))


def host_rule(rule, values):
    """-> ("ok", v) | ("error", reason) for every rule that needs no embeddings."""
    name = rule[0]
    if name == "exact_match":
        u = _uniq(values)
        return ("ok", u[0]) if len(u) == 1 else ("error", "no_consensus")
    if name == "mode_selection":
        return ("ok", mode_value(values))
    if name == "union_merge":
        flat = []
        stack = list(values)[::-1]
        while stack:
            v = stack.pop()
            if isinstance(v, list):
                stack.extend(v[::-1])
            else:
                flat.append(v)
        return ("ok", _uniq(flat))
    if name == "structural_merge":
        def merge(a, b):
            if isinstance(a, dict) and isinstance(b, dict):
                out = dict(a)
                for k, v in b.items():
                    out[k] = merge(out[k], v) if isinstance(out.get(k), dict) and isinstance(v, dict) else v
                return out
            return b if isinstance(b, dict) else a
        acc: Any = {}
        for v in values:
            acc = merge(acc, v)
        return ("ok", acc)
    if name == "percentile":
        nums = sorted(v for v in values if isinstance(v, (int, float)) and not isinstance(v, bool))
        if not nums:
            return ("ok", mode_value(values))
        idx = rule[1] / 100.0 * (len(nums) - 1)
        lo = int(idx)
        hi = min(lo + 1, len(nums) - 1)
        x = nums[lo] + (nums[hi] - nums[lo]) * (idx - lo)
        return ("ok", int(math.floor(x + 0.5)) if x >= 0 else -int(math.floor(-x + 0.5)))
    return ("error", "unknown_rule")            # :first_non_nil / :merge_maps have no clause -> mode fallback


# ---- fingerprints and clustering (aggregator.ex:23-103,366-454) ----------------------------------------------
_PUNCT = re.compile(r"[^\w\s]", re.ASCII)


def _key_terms(s: str, threshold: float) -> str:
    s = s.lower()
    if threshold < 0.95:
        s = _PUNCT.sub("", s)
    words = sorted((w for w in s.split() if len(w) > 3), key=lambda w: w.encode())
    return "_".join(words[:5])


def fingerprint(response: dict):
    """action_fingerprint/1 (aggregator.ex:72-103): (action, signature); the signature is the list of sub-action types
    for batches, the map {param: normalised value} otherwise (create_action_signature/2, :366-421)."""
    action = response["action"]
    params = response.get("params") or {}
    if action in ("batch_sync", "batch_async"):
        acts = params.get("actions") if isinstance(params, dict) and isinstance(params.get("actions"), list) else []
        seq = [a["action"] if isinstance(a, dict) and isinstance(a.get("action"), str) else "unknown" for a in acts]
        return (action, sorted(seq) if action == "batch_async" else seq)
    schema = SCHEMAS.get(action)
    if schema is None:
        return (action, "invalid")
    sig = {}
    for p in schema["required"] + schema["optional"]:
        v = params.get(p)
        if v is None:
            continue
        rule = schema["rules"].get(p)
        name = rule[0] if rule else None
        if name == "semantic_similarity":
            v = _key_terms(v, rule[1] or 0.9) if isinstance(v, str) else v
        elif name == "mode_selection":
            v = "_mode_mergeable"
        elif name == "percentile":
            v = "_percentile_mergeable"
        elif name == "union_merge" and isinstance(v, list):
            v = sorted(v, key=_order)
        sig[p] = v
    return (action, sig)


def cluster(responses: Sequence[dict]) -> list:
    """cluster_responses/1 (aggregator.ex:23-35): group by fingerprint, stable sort by count descending.  Equal counts
    keep BEAM map order over the fingerprint keys, which for the <= 32 clusters of a pool is ascending Erlang term
    order: action atom by name, then the signature (same restatement as the oracle)."""
    groups: "OrderedDict" = OrderedDict()
    for r in responses:
        fp = fingerprint(r)
        groups.setdefault((fp[0], _key(fp[1])), (fp, []))[1].append(r)
    ordered = sorted(groups.values(), key=lambda g: (str(g[0][0]).encode("utf-8"), _order(g[0][1])))
    clusters = [{"count": len(a), "actions": a, "representative": a[0], "fingerprint": fp} for fp, a in ordered]
    return sorted(clusters, key=lambda c: -c["count"])


def find_majority(clusters, total, round_num):
    for c in clusters:
        if (round_num == 1 and c["count"] == total) or (round_num != 1 and c["count"] > total / 2):
            return c
    return None


# ---- winner, tie-break, confidence (result.ex:261-308, scoring.ex:15-102) --------------------------------------
def _wait_score(v):
    if v is True:
        return (0, 0)
    if v is None:
        return (0, 1)
    if v is False or v == 0:
        return (1, 0)
    return (0, 1 + v)


def _priority(cluster):
    rep = cluster["representative"]
    if rep["action"] in ("batch_sync", "batch_async"):
        acts = (rep.get("params") or {}).get("actions") or []
        return max((PRIORITIES.get(a.get("action"), 999) for a in acts), default=999)
    return PRIORITIES.get(rep["action"], 999)


def pick_winner(clusters, total):
    for c in clusters:
        if c["count"] > total / 2:
            return "consensus", c
    top = max(c["count"] for c in clusters)
    tied = [c for c in clusters if c["count"] == top]
    def score(c):
        ws = [_wait_score(a.get("wait")) for a in c.get("actions", [])]
        return (_priority(c), (sum(w[0] for w in ws), sum(w[1] for w in ws)))
    return "forced_decision", sorted(tied, key=score)[0]


def confidence(count, total, round_num, max_rounds=4):
    frac = count / total
    bonus = 0.15 if frac > 0.8 else 0.10 if frac > 0.6 else 0.05 if frac > 0.5 else 0.0
    penalty = (round_num - max_rounds) * 0.1 if round_num > max_rounds else 0.0
    return min(max(frac + bonus - penalty, 0.1), 1.0)


# ---- the batched decide step -----------------------------------------------------------------------------------
# ---- :batch_sequence_merge (consensus_rules.ex:33-44,107-112,375-471) in two passes -----------------------------
def _plan_batch_merge(collector: DecisionCollector, agent_id: str, tag: str, sequences: list):
    """Host pass: shape checks and every rule that needs no embeddings; semantic parameters of the sub-actions are
    queued on the collector (they ride in the same GPU flush as everything else in the tree).
    -> ("done", result) when the outcome is already known, else ("plan", positions)."""
    if sequences == []:
        return ("done", ("ok", []))
    if len(sequences) == 1:
        return ("done", ("ok", sequences[0]))
    if len({len(q) for q in sequences}) > 1:
        return ("done", ("error", "sequence_length_mismatch"))
    positions = []
    for pos, actions in enumerate(zip(*sequences)):
        norm = [{"action": a.get("action"), "params": a.get("params") if a.get("params") is not None else {}} for a in actions]
        types = _uniq([a["action"] for a in norm])
        if len(types) != 1:
            positions.append({"error": "sequence_mismatch"})
            break                                             # later positions are never looked at (:385-396)
        schema = SCHEMAS.get(types[0])
        if schema is None:
            positions.append({"error": "unknown_action"})
            break
        plist = [a["params"] for a in norm]
        keys = _uniq([k for prm in plist for k in prm.keys()])
        entry = {"action": types[0], "keys": [], "merged": {}, "waiting": {}, "failed": False}
        for key in keys:
            vals = [prm.get(key) for prm in plist if prm.get(key) is not None]
            if not vals:
                continue
            entry["keys"].append(key)
            rule = schema["rules"].get(key, ["exact_match"])  # default differs from result.ex:334
            if rule[0] == "semantic_similarity":
                name = f"{tag}[{pos}].{key}"
                entry["waiting"][key] = name
                collector.submit(agent_id, name, vals, rule[1] if len(rule) > 1 else DEFAULT_THRESHOLD)
            else:
                r = host_rule(rule, vals)
                if r[0] == "ok":
                    entry["merged"][key] = r[1]
                else:
                    entry["failed"] = True                    # any failing rule -> {:error, :no_consensus} for the merge
        positions.append(entry)
    return ("plan", positions)


def _finish_batch_merge(positions: list, results: dict, agent_id: str):
    """After the flush: -> ("ok", merged_actions) | ("error", reason), the first failure in position order wins."""
    out = []
    for entry in positions:
        if "error" in entry:
            return ("error", entry["error"])
        merged = dict(entry["merged"])
        failed = entry["failed"]
        for key, name in entry["waiting"].items():
            r = results[agent_id][name]["result"]
            if r[0] == "ok":
                merged[key] = r[1]
            else:
                failed = True
        if failed:
            return ("error", "no_consensus")
        out.append({"action": entry["action"], "params": {k: merged[k] for k in entry["keys"]}})
    return ("ok", out)


def _merge_sequences_with_mode(sequences: list) -> list:
    """merge_sequences_with_mode/1 (result.ex:163-228): position by position, the most frequent action type and,
    among the actions of that type, the most frequent value of every parameter."""
    out = []
    for idx in range(max((len(q) for q in sequences), default=0)):
        at = [q[idx] for q in sequences if idx < len(q) and q[idx] is not None]
        if not at:
            continue
        if len(at) == 1:
            out.append({"action": at[0].get("action"), "params": dict(at[0].get("params") or {})})
            continue
        action_type = mode_value([a.get("action") for a in at])
        params = [a.get("params") or {} for a in at if a.get("action") == action_type]
        merged = {}
        for k in _uniq([k for prm in params for k in prm.keys()]):
            vals = [prm.get(k) for prm in params if prm.get(k) is not None]
            if vals:
                merged[k] = mode_value(vals)
        out.append({"action": action_type, "params": merged})
    return out


def decide_tree(collector: DecisionCollector, pending: Sequence[dict], round_num: int = 2,
                max_refinement_rounds: int = 4) -> list:
    """pending: [{"agent_id", "responses": [ {action, params, reasoning, wait?} x N ]}] ->
    [{"agent_id", "kind": "consensus" | "forced_decision" | "no_majority", "action", "confidence"}].

    Pass 1 (host): cluster, find the winner, merge every non-semantic parameter, enqueue every semantic one.
    One collector flush (GPU).  Pass 2 (host): fill the semantic parameters, mode fallback on error."""
    plans = []
    default_round, default_max = round_num, max_refinement_rounds
    for d in pending:
        responses = d["responses"]
        clusters = cluster(responses)
        # per-decision overrides (consensus_driver.run_consensus_rounds): the reference counts majorities against the
        # POOL size (consensus.ex:297), agents of one sweep sit at different rounds, and a dead refinement formats the
        # answers in hand whatever the majority says ("force")
        total = d.get("total", len(responses))
        round_num = d.get("round", default_round)
        max_refinement_rounds = d.get("max_refinement_rounds", default_max)
        if not d.get("force") and find_majority(clusters, total, round_num) is None and round_num <= max_refinement_rounds:
            plans.append({"agent_id": d["agent_id"], "kind": "no_majority", "clusters": clusters})
            continue
        kind, win = pick_winner(clusters, total)
        rep = win["representative"]
        schema = SCHEMAS.get(rep["action"])
        merged, waiting, batch = {}, {}, {}
        if rep["action"] == "batch_sync":                     # result.ex:58-84: the whole action is one sequence merge
            seqs = [((a.get("params") or {}).get("actions") or []) for a in win["actions"]]
            batch["__sync__"] = (_plan_batch_merge(collector, d["agent_id"], "batch_sync", seqs), seqs)
        elif schema is not None:
            for p in schema["required"] + schema["optional"]:
                vals = [v for v in ((a.get("params") or {}).get(p) for a in win["actions"]) if v is not None]
                if not vals:
                    continue
                rule = schema["rules"].get(p, ["mode_selection"])
                if rule[0] == "batch_sequence_merge":         # batch_async: a parameter like any other, mode on failure
                    batch[p] = (_plan_batch_merge(collector, d["agent_id"], f"batch:{p}", vals), vals)
                elif rule[0] == "semantic_similarity":
                    waiting[p] = vals
                    collector.submit(d["agent_id"], p, vals, rule[1] if len(rule) > 1 else DEFAULT_THRESHOLD)
                else:
                    r = host_rule(rule, vals)
                    merged[p] = r[1] if r[0] == "ok" else mode_value(vals)
        plans.append({"agent_id": d["agent_id"], "kind": kind, "win": win, "schema": schema, "merged": merged,
                      "waiting": waiting, "batch": batch, "total": total, "round": round_num,
                      "max_rounds": max_refinement_rounds})
    results = collector.flush()
    out = []
    for pl in plans:
        if pl["kind"] == "no_majority":
            out.append({"agent_id": pl["agent_id"], "kind": "no_majority", "action": None, "confidence": None})
            continue
        win, merged = pl["win"], pl["merged"]
        for p, vals in pl["waiting"].items():
            r = results[pl["agent_id"]][p]["result"]
            merged[p] = r[1] if r[0] == "ok" else mode_value(vals)
        rep = win["representative"]
        batch_results = {}
        for p, ((state, payload), vals) in pl["batch"].items():
            batch_results[p] = payload if state == "done" else _finish_batch_merge(payload, results, pl["agent_id"])
        waits = [a.get("wait") for a in win["actions"] if a.get("wait") is not None]
        if rep["action"] == "batch_sync":
            r = batch_results["__sync__"]
            seqs = pl["batch"]["__sync__"][1]
            if r[0] == "ok" or r == ("error", "no_consensus"):
                actions = r[1] if r[0] == "ok" else _merge_sequences_with_mode(seqs)
                action = {"action": "batch_sync", "params": {"actions": actions},
                          "reasoning": next((a.get("reasoning") for a in win["actions"] if a.get("reasoning") not in ("", None)), ""),
                          "wait": wait_rule(waits) if waits else False}
            else:
                # the reference destructures {:error, reason} as {action, accumulator} here (result.ex:30-47 over
                # :58-84), so the "action" it reports is the atom :error; kept visible rather than papered over
                action = {"action": "error", "error": r[1]}
        elif pl["schema"] is None:
            action = dict(rep)
        else:
            for p, r in batch_results.items():
                vals = pl["batch"][p][1]
                merged[p] = r[1] if r[0] == "ok" else mode_value(vals)
            ordered = {p: merged[p] for p in pl["schema"]["required"] + pl["schema"]["optional"] if p in merged}
            waits = [a.get("wait") for a in win["actions"] if a.get("wait") is not None]
            action = {"action": rep["action"], "params": ordered,
                      "reasoning": next((a.get("reasoning") for a in win["actions"] if a.get("reasoning") != ""), ""),
                      "wait": wait_rule(waits) if waits else False}
        out.append({"agent_id": pl["agent_id"], "kind": pl["kind"], "action": action,
                    "confidence": confidence(win["count"], pl["total"], pl["round"], pl["max_rounds"])})
    return out
