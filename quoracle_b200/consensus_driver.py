"""The consensus round loop for a whole agent tree at once.

Reference, one agent at a time inside its GenServer: `get_consensus_with_state/2` -> `execute_consensus_process/3`
-> on no majority `execute_refinement/3` re-queries the pool and recurses with round + 1, until a majority appears
or the round passes `max_refinement_rounds` (lib/quoracle/agent/consensus.ex:64-195,295-390; refinement answers:
lib/quoracle/agent/consensus/per_model_query.ex:405-443); the agent's decide step wraps that with a retry of at
most three attempts for `:all_models_failed` / `:all_responses_invalid` (lib/quoracle/agent/message_handler.ex
:125-148,353-389).

Here the loop is turned inside out: every agent of the tree is at some round; one SWEEP takes all agents whose
responses are in, clusters them on the host, and formats everyone who is done (majority, or past the last round, or
a dead refinement) with ONE `decide_tree` call -- one GPU flush for every semantic merge of the sweep -- while the
others are re-queried and join the next sweep.  Results per agent are what the sequential loop gives
(oracle/consensus_driver.py; tests/test_host_logic.py, tests/test_gpu_host.py).

LLM I/O stays behind the reference's own seam (`:model_query_fn`, per_model_query/helpers.ex:112-124), one level
up: `query_fn(state, model_pool, opts) -> ("ok", responses, state') | ("error", reason)` with parsed proposals
(None = unparsable); `validate_fn(response) -> response | None` stands in for the schema validator.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

from .collector import DecisionCollector
from .consensus import cluster, decide_tree, find_majority

MAX_CONSENSUS_ATTEMPTS = 3                                            # message_handler.ex:354
RETRYABLE = ("all_responses_invalid", "all_models_failed")            # message_handler.ex:353


def _validated(responses, validate_fn):
    valid, invalid = [], 0
    for r in responses:
        if r is None:
            continue
        v = validate_fn(r) if validate_fn is not None else r
        if v is None:
            invalid += 1
        else:
            valid.append(v)
    return valid, invalid


def run_consensus_rounds(collector: DecisionCollector, agents: Sequence[dict], query_fn: Callable,
                         validate_fn: Optional[Callable] = None) -> dict:
    """agents: [{"agent_id", "state" (with "model_histories"), "opts"?: {model_pool?, round?, force_reflection?}}].
    -> {agent_id: ("ok", {"kind", "action", "confidence", "round", "meta"}, state') | ("error", reason)}.
    `sweeps` (GPU flushes) is reported under the key "__sweeps__"."""
    out: dict = {}
    live = []
    for a in agents:                                                   # get_consensus_with_state/2 up to the first query
        state, opts = a["state"], dict(a.get("opts") or {})
        if state.get("model_histories") is None:
            out[a["agent_id"]] = ("error", "missing_model_histories")
            continue
        pool = opts.get("model_pool", list(state["model_histories"].keys()))
        opts.setdefault("max_refinement_rounds", state.get("max_refinement_rounds", 4))
        r = query_fn(state, pool, opts)
        if r[0] != "ok":
            out[a["agent_id"]] = ("error", r[1])
            continue
        valid, invalid = _validated(r[1], validate_fn)
        if not valid:
            out[a["agent_id"]] = ("error", "all_responses_invalid" if invalid > 0 else "all_models_failed")
            continue
        # `returned`: get_consensus_with_state/2 hands back the state of the FIRST query (consensus.ex:99,188); the states
        # of refinement queries only travel inside the context (consensus.ex:375)
        live.append({"agent_id": a["agent_id"], "pool": pool, "opts": opts, "state": r[2], "returned": r[2], "responses": valid,
                     "round": opts.get("round", 1), "max_rounds": opts["max_refinement_rounds"], "forced_meta": False,
                     "force_reflection": opts.get("force_reflection", state.get("force_reflection", False))})
    sweeps = 0
    while live:
        finishing, nxt = [], []
        for ag in live:
            clusters = cluster(ag["responses"])
            total = len(ag["pool"])                                    # consensus.ex:297: the POOL size, not the answers
            forced = bool(ag["force_reflection"]) and total == 1 and ag["round"] == 1
            majority = find_majority(clusters, total, ag["round"]) is not None
            if (majority and not forced) or (not majority and ag["round"] > ag["max_rounds"]):
                finishing.append(dict(ag, total=total))
                continue
            if forced:
                ag["forced_meta"] = True                                # v39.0: marked in the result's meta
            # execute_refinement/3: re-query with round + 1; a dead query or an empty answer formats what we have, with
            # total = the number of answers in hand (consensus.ex:347,386; per_model_query.ex:427-438)
            opts = dict(ag["opts"], round=ag["round"] + 1, refinement_prompt=("refine", ag["round"]))
            r = query_fn(ag["state"], ag["pool"], opts)
            if r[0] != "ok":
                finishing.append(dict(ag, total=len(ag["responses"]), force=True))
                continue
            valid, _ = _validated(r[1], validate_fn)
            if not valid:
                finishing.append(dict(ag, state=r[2], total=len(ag["responses"]), force=True))
                continue
            nxt.append(dict(ag, state=r[2], responses=valid, round=ag["round"] + 1))
        if finishing:
            sweeps += 1
            pending = [{"agent_id": f["agent_id"], "responses": f["responses"], "total": f["total"], "round": f["round"],
                        "max_refinement_rounds": f["max_rounds"], "force": f.get("force", False)} for f in finishing]
            for f, res in zip(finishing, decide_tree(collector, pending)):
                meta = {"per_model_queries": True}
                if f["forced_meta"]:
                    meta["forced_reflection_applied"] = True
                out[f["agent_id"]] = ("ok", {"kind": res["kind"], "action": res["action"], "confidence": res["confidence"],
                                             "round": f["round"], "meta": meta}, f["returned"])
        live = nxt
    out["__sweeps__"] = sweeps
    return out


def run_consensus_cycles(collector: DecisionCollector, agents: Sequence[dict], query_fn: Callable,
                         execute_action_fn: Callable, validate_fn: Optional[Callable] = None) -> dict:
    """The per-agent decide step (run_consensus_cycle/2) for a whole tree: every agent's cycle, retried up to
    MAX_CONSENSUS_ATTEMPTS on the two retryable errors, all attempts of one generation sharing the sweeps above.
    -> {agent_id: (state, ("executed", action) | ("stalled", reason, attempts) | ("failed", reason))}."""
    states = {a["agent_id"]: dict(a["state"]) for a in agents}
    opts = {a["agent_id"]: a.get("opts") for a in agents}
    todo = [a["agent_id"] for a in agents]
    done: dict = {}
    while todo:
        res = run_consensus_rounds(collector, [{"agent_id": i, "state": states[i], "opts": opts[i]} for i in todo], query_fn,
                                   validate_fn)
        again = []
        for i in todo:
            r = res[i]
            st = states[i]
            if r[0] == "ok":
                merged = dict(st, **r[2])
                merged["consensus_retry_count"] = 0
                merged["correction_feedback"] = {}
                done[i] = (execute_action_fn(merged, r[1]["action"]), ("executed", r[1]["action"]))
                continue
            reason, retry = r[1], st.get("consensus_retry_count", 0)
            if reason in RETRYABLE and retry + 1 < MAX_CONSENSUS_ATTEMPTS:
                st["consensus_retry_count"] = retry + 1
                st["correction_feedback"] = {m: reason for m in (st.get("model_histories") or {})}
                again.append(i)
            elif reason in RETRYABLE:
                st["correction_feedback"] = {}
                done[i] = (st, ("stalled", reason, retry + 1))
            else:
                done[i] = (st, ("failed", reason))
        todo = again
    return done
