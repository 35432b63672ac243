"""Property tests of the oracle mirroring the reference's StreamData properties
(test/quoracle/actions/consensus_rules_test.exs:382-466) plus structural properties of the restated math."""
import math

import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import consensus_rules as rules
from oracle import numeric, result, terms

texts = st.text(alphabet=st.characters(min_codepoint=32, max_codepoint=126), max_size=12)


@given(st.lists(texts, min_size=1, max_size=8))
def test_exact_match_is_commutative(values):                      # consensus_rules_test.exs:383-389
    assert rules.apply_rule(("exact_match",), values) == rules.apply_rule(("exact_match",), values[::-1])


@given(st.lists(st.lists(st.sampled_from(list("abcdefgh")), max_size=10), min_size=1, max_size=5))
def test_union_merge_includes_all_unique_elements(lists):        # :391-405
    ok, merged = rules.apply_rule(("union_merge",), lists)
    flat = [x for l in lists for x in l]
    assert ok == "ok" and sorted(merged) == sorted(set(flat)) and merged == terms.uniq(flat)


@given(st.lists(st.integers(-1000, 1000), min_size=1, max_size=20))
def test_mode_selection_returns_an_input_value(values):          # :407-412
    ok, v = rules.apply_rule(("mode_selection",), values)
    assert ok == "ok" and v in values
    assert values.count(v) == max(values.count(x) for x in values)


@given(st.lists(st.integers(-10**6, 10**6), min_size=1, max_size=20), st.integers(0, 100))
def test_percentile_within_min_max(values, p):                   # :414-424
    ok, v = rules.apply_rule(("percentile", p), values)
    assert ok == "ok" and min(values) <= v <= max(values)


@given(st.lists(st.dictionaries(st.sampled_from(list("abcdef")), st.integers(), max_size=6), min_size=1, max_size=5))
def test_structural_merge_preserves_all_keys(maps):              # :450-465
    ok, merged = rules.apply_rule(("structural_merge",), maps)
    assert ok == "ok" and sorted(merged) == sorted({k for m in maps for k in m})


vec = st.lists(st.floats(-10, 10, allow_nan=False, allow_infinity=False, width=32), min_size=1, max_size=24)


@given(vec, st.floats(0.01, 100.0))
def test_cosine_is_scale_invariant_symmetric_and_bounded(a, scale):
    b = [x * 0.5 + 0.25 for x in a]
    c1 = numeric.cosine_similarity_rules(a, b)
    assert -1.0 - 1e-12 <= c1 <= 1.0 + 1e-12
    assert abs(c1 - numeric.cosine_similarity_rules(b, a)) < 1e-12
    assert abs(c1 - numeric.cosine_similarity_rules([x * scale for x in a], b)) < 1e-9
    assert abs(c1 - numeric.cosine_similarity_aggregator(a, b)) < 1e-12
    if any(x != 0.0 for x in a):
        assert abs(numeric.cosine_similarity_rules(a, a) - 1.0) < 1e-12
    assert numeric.cosine_similarity_rules(a, [0.0] * len(a)) == 0.0


@settings(max_examples=40, deadline=None)
@given(st.integers(2, 9), st.integers(1, 40), st.integers(0, 10**6))
def test_medoid_and_rule_are_permutation_equivariant(n, d, seed):
    rng = np.random.default_rng(seed)
    emb = rng.standard_normal((1, n, d))
    ok, med, cos, m = numeric.decide_batch(emb, np.array([0.2]))
    perm = rng.permutation(n)
    ok2, med2, cos2, m2 = numeric.decide_batch(emb[:, perm], np.array([0.2]))
    assert np.allclose(cos2[0], cos[0][perm][:, perm], atol=1e-12)
    if m["vote"][0] > 1e-9:
        assert perm[med2[0]] == med[0]
    if perm[0] == 0 and m["thr"][0] > 1e-9:
        assert ok2[0] == ok[0]


@given(st.integers(1, 12), st.integers(1, 12), st.integers(1, 12))
def test_confidence_bounds_and_monotonicity(count, extra, round_num):
    total = count + extra - 1
    c = result.calculate_confidence({"count": count}, total, round_num)
    assert 0.1 <= c <= 1.0
    if count < total:
        assert result.calculate_confidence({"count": count + 1}, total, round_num) >= c
    assert result.calculate_confidence({"count": count}, total, round_num + 1) <= c + 1e-15


@given(st.lists(st.one_of(st.booleans(), st.integers(0, 100)), min_size=1, max_size=7))
def test_wait_rule_total_and_typed(values):
    ok, v = rules.apply_rule(("wait_parameter",), values)
    assert ok == "ok" and isinstance(v, (bool, int))
    if all(isinstance(x, bool) for x in values) and len(set(values)) == 1:
        assert v is values[0]
    ws = [result.wait_score(x) for x in values]
    assert all(len(w) == 2 and w[0] in (0, 1) for w in ws)
    assert math.isfinite(sum(w[1] for w in ws))


# ---- the round loop / decide step restatement (oracle/consensus_driver.py) against facts the reference's tests hold ----
def _orient(text="Processing"):
    return {"action": "orient", "params": {"current_situation": text, "goal_clarity": "Clear",
                                           "available_resources": "Available", "key_challenges": "None",
                                           "delegation_consideration": "none"}, "reasoning": "Reconsidering"}


def test_driver_facts_held_by_the_reference_tests():
    from oracle import consensus_driver as drv
    state = {"agent_id": "a", "model_histories": {"model-a": [], "model-b": []}}

    def agree(st, pool, opts):
        return ("ok", [_orient() for _ in pool], dict(st, seen_round=opts.get("round", 1)))

    # test/quoracle/agent/consensus_per_model_test.exs:294 -- no per-model histories -> {:error, :missing_model_histories}
    assert drv.get_consensus_with_state({"agent_id": "a"}, {}, agree) == ("error", "missing_model_histories")
    # test/quoracle/agent/consensus_refinement_context_test.exs:88-101 -- a pool whose members return the same valid
    # orient answer completes: result type in [:consensus, :forced_decision], action not nil, state handed back
    r = drv.get_consensus_with_state(state, {}, agree)
    assert r[0] == "ok" and r[1][0] == "consensus" and r[1][1]["action"] == "orient" and r[1][2]["per_model_queries"] is True
    assert r[2]["seen_round"] == 1
    # test/quoracle/agent/consensus_test.exs:177-193 -- no majority through every refinement round -> :forced_decision
    # with confidence <= 0.5
    calls = []

    def never(st, pool, opts):
        calls.append(opts.get("round", 1))
        # two different ACTIONS (different fingerprints; rewording a semantic parameter would still cluster together)
        return ("ok", [_orient(), {"action": "send_message", "params": {"to": "parent", "content": "status"},
                                   "reasoning": "r"}][:len(pool)], st)

    r = drv.get_consensus_with_state(state, {}, never)
    assert r[0] == "ok" and r[1][0] == "forced_decision" and r[1][2]["confidence"] <= 0.5
    assert calls == [1, 2, 3, 4, 5]                       # default max_refinement_rounds = 4 (consensus.ex:90): rounds 1..5
    # test/quoracle/agent/consensus_test.exs:150-163 -- every model failing is {:error, :all_models_failed}
    assert drv.get_consensus_with_state(state, {}, lambda st, pool, o: ("ok", [None, None], st)) == ("error", "all_models_failed")
    # answers that do not validate: {:error, :all_responses_invalid} (consensus.ex:132-134)
    assert drv.get_consensus_with_state(state, {}, agree, validate_fn=lambda r_: None) == ("error", "all_responses_invalid")
    # lib/quoracle/agent/message_handler.ex:353-354 -- at most three attempts for the two retryable errors
    attempts = []

    def failing(st):
        attempts.append(st.get("consensus_retry_count", 0))
        return ("error", "all_models_failed", None)

    st, out = drv.run_consensus_cycle(state, failing, lambda s, a: s)
    assert attempts == [0, 1, 2] and out == ("stalled", "all_models_failed", 3) and st["correction_feedback"] == {}
    st, out = drv.run_consensus_cycle(state, lambda s: ("error", "rate_limit_exceeded", None), lambda s, a: s)
    assert out == ("failed", "rate_limit_exceeded")
