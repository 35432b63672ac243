"""Restatement of the consensus round loop and the per-agent decide step -- TEST INFRASTRUCTURE (oracle/__init__.py).

  execute_consensus_process/3, execute_refinement/3      reference lib/quoracle/agent/consensus.ex:295-390
  PerModelQuery.handle_refinement_responses/5            reference lib/quoracle/agent/consensus/per_model_query.ex:405-443
  get_consensus_with_state/2 (the part after the query)  reference lib/quoracle/agent/consensus.ex:80-195
  MessageHandler.run_consensus_cycle/2 + retry <= 3      reference lib/quoracle/agent/message_handler.ex:125-148,353-389

One agent at a time, strictly sequential, as the reference runs it inside each agent's GenServer.  LLM I/O is
outside the hot path: it enters through the reference's own seam, an injectable query function
(`:model_query_fn`, lib/quoracle/agent/consensus/per_model_query/helpers.ex:112-124).  Here the seam is one level
up -- `query_fn(state, model_pool, opts) -> ("ok", responses, state') | ("error", reason)`, the contract of
query_models_with_per_model_histories/3 (consensus.ex:99,351-355) -- and `responses` are already parsed proposals
({"action", "params", "reasoning", "wait"?} or None for an unparsable answer), because JSON extraction and schema
validation (ActionParser, Validator) are not on the embed -> cosine -> verdict path; `validate_fn(response) ->
response | None` stands in for Validator.validate_params/3 (consensus.ex:268-293).
"""
from __future__ import annotations

from typing import Callable, Optional

from . import aggregator, result

MAX_CONSENSUS_ATTEMPTS = 3                                            # message_handler.ex:354
RETRYABLE = ("all_responses_invalid", "all_models_failed")            # message_handler.ex:353


def _cost_opts(context):
    """extract_cost_opts/1 -- consensus.ex:415-418: only these keys survive (an :embedding_fn does NOT; the drop-in
    therefore replaces Embeddings.get_embedding/2 itself, SURVEY.md section 0 fact 5 -- the oracle threads the
    embedding function explicitly instead)."""
    opts = dict(context.get("original_opts") or {})
    keep = {k: opts[k] for k in ("agent_id", "task_id", "pubsub", "cost_accumulator", "max_refinement_rounds") if k in opts}
    if "embedding_fn" in context:
        keep["embedding_fn"] = context["embedding_fn"]
    return keep


def filter_invalid_responses(responses, validate_fn: Optional[Callable] = None):
    """consensus.ex:268-293 -> (valid, invalid_count); validated (possibly coerced) responses are kept."""
    valid, invalid = [], 0
    for r in responses:
        v = validate_fn(r) if validate_fn is not None else r
        if v is None:
            invalid += 1
        else:
            valid.append(v)
    return valid, invalid


def execute_consensus_process(responses, context, round_, query_fn, validate_fn=None):
    """consensus.ex:295-330 -> ("ok", (kind, action, meta))."""
    clusters = aggregator.cluster_responses(responses)
    total = len(context["model_pool"])
    max_rounds = context.get("max_refinement_rounds", 4)
    cost_opts = dict(_cost_opts(context), max_refinement_rounds=max_rounds)
    forced = bool(context.get("force_reflection")) and len(context["model_pool"]) == 1 and round_ == 1
    maj = aggregator.find_majority_cluster(clusters, total, round_)
    if maj[0] == "majority" and not forced:
        return ("ok", result.format_result(clusters, total, round_, cost_opts))
    if maj[0] == "no_majority" and round_ > max_rounds:
        return ("ok", result.format_result(clusters, total, round_, cost_opts))
    if forced:
        r = execute_refinement(responses, context, round_, query_fn, validate_fn)
        if r[0] == "ok":
            kind, action, meta = r[1]
            return ("ok", (kind, action, dict(meta, forced_reflection_applied=True)))
        return r
    return execute_refinement(responses, context, round_, query_fn, validate_fn)


def execute_refinement(responses, context, round_, query_fn, validate_fn=None):
    """consensus.ex:332-390."""
    context = dict(context)
    context.setdefault("round_history", [])
    context["round_history"] = context["round_history"] + [(round_, responses)]     # Manager.update_context_with_round
    state = context.get("state")
    if state is None:                      # messages-based API: no per-model histories, forced decision on what we have
        clusters = aggregator.cluster_responses(responses)
        return ("ok", result.format_result(clusters, len(responses), round_, _cost_opts(context)))
    opts = dict(context.get("original_opts") or {}, round=round_ + 1, refinement_prompt=("refine", round_))
    r = query_fn(state, context["model_pool"], opts)
    if r[0] != "ok":
        clusters = aggregator.cluster_responses(responses)
        return ("ok", result.format_result(clusters, len(responses), round_, _cost_opts(context)))
    _, refined, new_state = r
    context["state"] = new_state
    return handle_refinement_responses(refined, responses, context, round_, query_fn, validate_fn)


def handle_refinement_responses(refined, original, context, round_, query_fn, validate_fn=None):
    """per_model_query.ex:405-443: nothing usable came back -> format the ORIGINAL responses at this round."""
    valid = [r for r in refined if r is not None]
    validated, _ = filter_invalid_responses(valid, validate_fn)
    if not validated:
        clusters = aggregator.cluster_responses(original)
        return ("ok", result.format_result(clusters, len(original), round_, _cost_opts(context)))
    return execute_consensus_process(validated, context, round_ + 1, query_fn, validate_fn)


def get_consensus_with_state(state, opts, query_fn, validate_fn=None, embedding_fn=None):
    """consensus.ex:64-195 -> ("ok", (kind, action, meta), state') | ("error", reason)."""
    if state.get("model_histories") is None:
        return ("error", "missing_model_histories")
    opts = dict(opts or {})
    model_pool = opts.get("model_pool", list(state["model_histories"].keys()))
    max_rounds = state.get("max_refinement_rounds", 4)
    opts.setdefault("max_refinement_rounds", max_rounds)
    r = query_fn(state, model_pool, opts)
    if r[0] != "ok":
        return ("error", r[1])
    _, responses, updated_state = r
    non_nil = [x for x in responses if x is not None]
    valid, invalid = filter_invalid_responses(non_nil, validate_fn)
    if not valid and invalid > 0:
        return ("error", "all_responses_invalid")
    if not valid:
        return ("error", "all_models_failed")
    context = {"model_pool": model_pool, "original_opts": opts, "max_refinement_rounds": opts["max_refinement_rounds"],
               "force_reflection": opts.get("force_reflection", state.get("force_reflection", False)),
               "state": updated_state}
    if embedding_fn is not None:
        context["embedding_fn"] = embedding_fn
    out = execute_consensus_process(valid, context, opts.get("round", 1), query_fn, validate_fn)
    if out[0] != "ok":
        return out
    kind, action, meta = out[1]
    return ("ok", (kind, action, dict(meta, per_model_queries=True)), updated_state)


def run_consensus_cycle(state, consensus_fn, execute_action_fn):
    """message_handler.ex:125-148 with handle_consensus_error/4 (:353-389), iterated: a retryable failure schedules
    another cycle (StateUtils.schedule_consensus_continuation) until MAX_CONSENSUS_ATTEMPTS is used up.

    consensus_fn(state) -> ("ok", action, state', acc) | ("error", reason, acc)   (the :consensus_fn seam,
    message_handler.ex:32-44).  -> (state, outcome) with outcome ("executed", action) | ("stalled", reason, attempts)
    | ("failed", reason)."""
    state = dict(state)
    while True:
        r = consensus_fn(state)
        if r[0] == "ok":
            _, action, updated, _acc = r
            merged = dict(state, **updated)
            merged["consensus_retry_count"] = 0
            merged["correction_feedback"] = {}
            return execute_action_fn(merged, action), ("executed", action)
        _, reason, _acc = r
        retry = state.get("consensus_retry_count", 0)
        if reason in RETRYABLE and retry + 1 < MAX_CONSENSUS_ATTEMPTS:
            state["consensus_retry_count"] = retry + 1
            state["correction_feedback"] = {m: reason for m in (state.get("model_histories") or {})}
            continue
        if reason in RETRYABLE:
            state["correction_feedback"] = {}
            return state, ("stalled", reason, retry + 1)
        return state, ("failed", reason)
