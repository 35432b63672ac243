"""Pins the oracle against the reference's own known-answer tests (tests/golden/reference_vectors.json,
transcribed by tests/golden/make_golden.py with reference file:line per vector).  CPU only."""
import json
import math
from pathlib import Path

import numpy as np
import pytest

from oracle import aggregator, consensus_rules, cport, embeddings, lesson_manager, numeric, result
from oracle.consensus_rules import Accumulator

GOLD = json.loads((Path(__file__).parent / "golden" / "reference_vectors.json").read_text())
R = GOLD["rules"]


def _cos(case, a, b):
    fn = numeric.cosine_similarity_aggregator if case["fn"] == "aggregator" else numeric.cosine_similarity_rules
    return fn(a, b)


@pytest.mark.parametrize("case", GOLD["cosine"], ids=lambda c: c["src"])
def test_cosine_known_answers(case):
    got = _cos(case, case["a"], case["b"])
    c_got = (cport.cosine_aggregator if case["fn"] == "aggregator" else cport.cosine_rules)(case["a"], case["b"])
    for v in (got, c_got):
        if "equals" in case:
            assert abs(v - case["equals"]) <= case["tol"], case["src"]
        if "gt" in case:
            assert v > case["gt"]
        if "lt" in case:
            assert v < case["lt"]
    assert got == c_got                      # Python and C restatements agree bit for bit (same f64 order)


def test_cosine_length_mismatch():
    with pytest.raises(ValueError):           # ConsensusRules raises ArgumentError (consensus_rules.ex:189-191)
        numeric.cosine_similarity_rules([0.5] * 3072, [0.5] * 100)
    assert numeric.cosine_similarity_aggregator([1.0, 2.0, 3.0], [1.0, 2.0]) == 0.0   # aggregator.ex:350


def test_random_cosine_in_range_and_two_versions_agree():
    rng = np.random.default_rng(0)
    for _ in range(20):
        a, b = rng.random(3072) - 0.5, rng.random(3072) - 0.5
        c1 = numeric.cosine_similarity_rules(a.tolist(), b.tolist())
        c2 = numeric.cosine_similarity_aggregator(a.tolist(), b.tolist())
        assert -1.0 <= c1 <= 1.0
        assert abs(c1 - c2) < 1e-12


@pytest.mark.parametrize("case", R["apply_rule"], ids=lambda c: c["src"] + str(c["values"])[:30])
def test_apply_rule_known_answers(case):
    rule = tuple(case["rule"])
    fail = lambda v, o: ("error", "not_configured")            # noqa: E731  (no provider in tests)
    if rule[0] == "semantic_similarity":
        got = consensus_rules.apply_rule(rule, case["values"], {"embedding_fn": fail})
    else:
        got = consensus_rules.apply_rule(rule, case["values"])
    assert list(got) == case["expect"], case["src"]


@pytest.mark.parametrize("case", R["merge_params"], ids=lambda c: c["src"])
def test_merge_params_known_answers(case):
    assert list(consensus_rules.merge_params(case["action"], case["param"], case["values"])) == case["expect"]


def test_identical_values_make_no_embedding_call():
    calls = []

    def fn(v, o):
        calls.append(v)
        return ("ok", {"embedding": [1.0, 0.0]})

    assert consensus_rules.apply_rule(("semantic_similarity", 0.9), ["same", "same", "same"],
                                      {"embedding_fn": fn}) == ("ok", "same")
    assert calls == []                                          # consensus_rules.ex:251-256


def test_constant_embedding_rule_and_accumulator_threading():
    """consensus_rules_test.exs:487-552 and result_test.exs:839-884."""
    case = GOLD["constant_embedding_rule"]
    seen = []

    def fn(text, opts):
        seen.append((text, "cost_accumulator" in opts))
        acc = opts.get("cost_accumulator", Accumulator())
        return ("ok", {"embedding": case["embedding"]}, acc.add({"cost_type": "llm_embedding"}))

    res, acc = consensus_rules.apply_rule(("semantic_similarity", {"threshold": case["threshold"], "embedding_fn": fn}),
                                          case["values"], {"cost_accumulator": Accumulator()})
    assert res == ("ok", case["expect"])
    assert [s[0] for s in seen] == case["values"]               # sequential, pool order
    assert all(s[1] for s in seen)
    assert acc.count() >= 2 and acc.count() == len(case["values"])
    # without an accumulator the plain 2-tuple comes back
    fn2 = lambda t, o: ("ok", {"embedding": case["embedding"]})   # noqa: E731
    assert consensus_rules.apply_rule(("semantic_similarity", 0.8), ["a b", "c d"], {"embedding_fn": fn2}) == ("ok", "a b")


def test_reference_rule_threshold_boundary():
    e0 = [1.0, 0.0]
    e_ok = [0.9, math.sqrt(1 - 0.81)]          # cos = 0.9 exactly-ish
    fn = lambda v, o: ("ok", {"embedding": {"a": e0, "b": e_ok, "c": [0.0, 1.0]}[v]})   # noqa: E731
    assert consensus_rules.apply_rule(("semantic_similarity", 0.85), ["a", "b"], {"embedding_fn": fn}) == ("ok", "a")
    assert consensus_rules.apply_rule(("semantic_similarity", 0.85), ["a", "b", "c"], {"embedding_fn": fn}) == \
        ("error", "no_consensus")
    # any failed embedding -> error (-> mode fallback upstream), consensus_rules.ex:294-297
    bad = lambda v, o: ("error", "network_error") if v == "b" else ("ok", {"embedding": e0})   # noqa: E731
    assert consensus_rules.apply_rule(("semantic_similarity", 0.85), ["a", "b"], {"embedding_fn": bad}) == \
        ("error", "no_consensus")


@pytest.mark.parametrize("case", R["wait_score"], ids=lambda c: str(c["value"]))
def test_wait_score(case):
    assert list(result.wait_score(case["value"])) == case["expect"]


@pytest.mark.parametrize("case", R["cluster_wait_score"], ids=lambda c: c["src"])
def test_cluster_wait_score(case):
    actions = [({"action": "x", "params": {}} if w == "__missing__" else {"action": "x", "params": {}, "wait": w})
               for w in case["waits"]]
    assert list(result.cluster_wait_score({"actions": actions})) == case["expect"]


@pytest.mark.parametrize("case", R["break_tie"], ids=lambda c: c["src"])
def test_break_tie(case):
    clusters = [{"representative": {"action": a, "params": {}, "reasoning": ""}, "count": 2, "tag": i}
                for i, a in enumerate(case["actions"])]
    w = result.break_tie(clusters)
    if "expect" in case:
        assert w["representative"]["action"] == case["expect"]
    else:
        assert w["tag"] == case["expect_index"]


@pytest.mark.parametrize("case", R["confidence"], ids=lambda c: c["src"])
def test_confidence(case):
    def conf(over=None):
        c = dict(case)
        c.update(over or {})
        opts = {"max_refinement_rounds": c["max_rounds"]} if "max_rounds" in c else {}
        return result.calculate_confidence({"count": c["count"]}, c["total"], c["round"], opts)
    v = conf()
    if "gt" in case:
        assert v > case["gt"]
    if "ge" in case:
        assert case["ge"] <= v <= case["le"]
    if "gt_other" in case:
        assert v > conf(case["gt_other"])
    if "lt_other" in case:
        assert v < conf(case["lt_other"])
    if "same_as" in case:
        assert v == conf(case["same_as"])
    if "minus_other" in case:
        assert abs((v - conf(case["minus_other"])) - case["equals"]) <= case["tol"]


@pytest.mark.parametrize("case", R["format_result"], ids=lambda c: c["src"])
def test_format_result_plurality(case):
    clusters = [{"count": n, "actions": [{"action": a, "params": {}, "reasoning": ""}] * 0,
                 "representative": {"action": a, "params": {}, "reasoning": ""}} for a, n in case["clusters"]]
    kind, action, opts = result.format_result(clusters, case["total"], case["round"])
    assert kind == case["expect_kind"] and action["action"] == case["expect_action"]
    assert 0.1 <= opts["confidence"] <= 1.0


def test_mode_fallback_when_embeddings_unavailable():
    """result_test.exs:895-919: 4 orient proposals, embeddings unavailable -> merged value is the MODE."""
    case = R["mode_fallback"][0]
    actions = [{"action": "orient", "params": {case["param"]: v}, "reasoning": "r", "wait": False} for v in case["values"]]
    cluster = {"count": 4, "actions": actions, "representative": actions[0]}
    fail = lambda v, o: ("error", "not_configured")            # noqa: E731
    merged, acc = result.merge_cluster_params(cluster, {"embedding_fn": fail})
    assert merged["params"][case["param"]] == case["expect"]
    assert merged["wait"] is False and acc is None


@pytest.mark.parametrize("case", R["majority"], ids=lambda c: str(c["counts"]) + str(c["round"]))
def test_majority_thresholds(case):
    clusters = [{"count": n} for n in case["counts"]]
    assert aggregator.find_majority_cluster(clusters, case["total"], case["round"])[0] == case["expect"]


def test_clustering_fingerprints():
    """aggregator_test.exs:893-905 and :14-172 (shape): identical semantic params -> one cluster."""
    rs = [{"action": "spawn_child", "params": {"task_description": "analyze data"}, "reasoning": r}
          for r in ("Need data", "Get data", "Load data")]
    cl = aggregator.cluster_responses(rs)
    assert len(cl) == 1 and cl[0]["count"] == 3
    # key-term normalisation (aggregator.ex:423-454): case, punctuation and short words are ignored below 0.95
    a = {"action": "send_message", "params": {"to": "parent", "content": "Deploy the NEW build, please!"}, "reasoning": ""}
    b = {"action": "send_message", "params": {"to": "parent", "content": "please deploy build"}, "reasoning": ""}
    c = {"action": "send_message", "params": {"to": "child", "content": "please deploy build"}, "reasoning": ""}
    cl = aggregator.cluster_responses([a, b, c])
    assert [x["count"] for x in cl] == [2, 1]
    assert aggregator.action_fingerprint({"action": "nope", "params": {}}) == ("nope", "invalid")
    # mode / percentile params never split clusters
    f1 = aggregator.action_fingerprint({"action": "fetch_web", "params": {"url": "u", "timeout": 5}})
    f2 = aggregator.action_fingerprint({"action": "fetch_web", "params": {"url": "u", "timeout": 50}})
    assert f1 == f2
    # batch fingerprints: sync keeps order, async sorts
    s1 = aggregator.action_fingerprint({"action": "batch_sync", "params": {"actions": [{"action": "todo"}, {"action": "orient"}]}})
    s2 = aggregator.action_fingerprint({"action": "batch_async", "params": {"actions": [{"action": "todo"}, {"action": "orient"}]}})
    assert s1 == ("batch_sync", ["todo", "orient"]) and s2 == ("batch_async", ["orient", "todo"])


def test_calculate_semantic_similarity_shapes():
    """aggregator_test.exs:984-1098: 1-arity fn [1,0,0] for both texts -> 1.0; with accumulator -> (float, acc)."""
    fn1 = lambda text: ("ok", [1.0, 0.0, 0.0])                 # noqa: E731
    assert aggregator.calculate_semantic_similarity("a", "b", {"embedding_fn": fn1}) == 1.0

    def fn2(text, m):
        acc = m.get("cost_accumulator", Accumulator())
        return ("ok", {"embedding": [1.0, 0.0, 0.0]}, acc.add({"t": text}))

    sim, acc = aggregator.calculate_semantic_similarity("a", "b", {"embedding_fn": fn2, "embedding_fn_arity": 2,
                                                                   "cost_accumulator": Accumulator()})
    assert sim == 1.0 and acc.count() == 2
    assert aggregator.get_cached_embedding("cache-test", {"embedding_fn": lambda t: ("ok", [0.1, 0.2, 0.3])}) == \
        ("ok", {"embedding": [0.1, 0.2, 0.3], "cached": False, "chunks": 1})


@pytest.mark.parametrize("case", GOLD["lesson_dedup"], ids=lambda c: c["src"])
def test_lesson_dedup_vectors(case):
    assert numeric.first_match(case["new"], case["existing"], case["threshold"]) == case["expect_index"]
    assert cport.first_match(case["new"], case["existing"], case["threshold"]) == case["expect_index"]


def test_lesson_manager_flow():
    """lesson_manager_test.exs:34-76 mock: similar lessons merge (confidence + 1, NEW content kept)."""
    table = {"API requires bearer auth": [1.0, 0.0, 0.0], "API needs bearer token": [0.95, 0.05, 0.0],
             "Be concise in responses": [0.5, 0.5, 0.0], "Stay brief": [0.22, 0.97, 0.0]}
    fn = lambda t: ("ok", {"embedding": table[t]})             # noqa: E731
    ex = [{"type": "factual", "content": "API requires bearer auth", "confidence": 2}]
    ok, out = lesson_manager.accumulate_lessons(ex, [{"type": "factual", "content": "API needs bearer token", "confidence": 1}],
                                                {"embedding_fn": fn})
    assert out == [{"type": "factual", "content": "API needs bearer token", "confidence": 3}]
    ex = [{"type": "behavioral", "content": "Be concise in responses", "confidence": 1}]
    ok, out = lesson_manager.accumulate_lessons(ex, [{"type": "behavioral", "content": "Stay brief", "confidence": 1}],
                                                {"embedding_fn": fn})
    assert len(out) == 2
    lessons = [{"type": "factual", "content": f"l{i}", "confidence": c} for i, c in enumerate([1, 5, 2])]
    assert [l["confidence"] for l in lesson_manager.prune_lessons(lessons, 2)] == [5, 2]


@pytest.mark.parametrize("case", GOLD["chunk_mean"], ids=lambda c: c["src"])
def test_chunk_mean_constant(case):
    chunks = [[case["value"]] * case["dim"] for _ in range(case["chunks"])]
    m = numeric.average_embeddings(chunks)
    assert max(abs(x - case["expect"]) for x in m) <= case["tol"]
    assert np.max(np.abs(cport.chunk_mean(chunks) - np.array(m))) == 0.0


def test_embeddings_cache_chunking_and_errors():
    """model_query_embedding_test.exs:47-148 (shape): chunks > 1 for long text, cached second call, ttl 0 bypass."""
    calls = []

    def provider(text):
        calls.append(text)
        return ("ok", [0.01] * 16)

    table = {}
    opts = {"embedding_fn": provider, "cache_table": table, "model_token_limit": 10}     # limit -> 9 tokens/chunk
    long_text = " ".join(f"w{i}" for i in range(40))
    r = embeddings.get_embedding(long_text, opts)
    assert r[0] == "ok" and r[1]["chunks"] > 2 and r[1]["cached"] is False
    assert all(abs(x - 0.01) < 1e-15 for x in r[1]["embedding"])
    n = len(calls)
    r2 = embeddings.get_embedding(long_text, opts)
    assert r2[1]["cached"] is True and r2[1]["embedding"] == r[1]["embedding"] and len(calls) == n
    r3 = embeddings.get_embedding(long_text, dict(opts, cache_ttl=0))
    assert r3[1]["cached"] is False and len(calls) > n
    assert embeddings.get_embedding("", opts) == ("error", "invalid_input")              # embeddings.ex:49-50
    assert embeddings.get_embedding("x", {"cache_table": {}}) == ("error", "not_configured")
    # non-binary values are embedded as their JSON encoding (embeddings.ex:46)
    embeddings.get_embedding([{"a": 1}], dict(opts, cache_table={}))
    assert calls[-1] == '[{"a":1}]'
    # accumulator: non-cached -> 3-tuple, cached -> 2-tuple (embeddings.ex:53-57,116-127)
    t2 = {}
    a = embeddings.get_embedding("short text", {"embedding_fn": provider, "cache_table": t2, "cost_accumulator": Accumulator()})
    assert len(a) == 3 and a[2].count() == 1
    b = embeddings.get_embedding("short text", {"embedding_fn": provider, "cache_table": t2, "cost_accumulator": Accumulator()})
    assert len(b) == 2 and b[1]["cached"] is True
    # eviction: at 1000 entries the 100 oldest go (embeddings.ex:402-426)
    t3 = {}
    clock = [0]
    for i in range(1001):
        clock[0] = i
        embeddings.get_embedding(f"text {i}", {"embedding_fn": provider, "cache_table": t3, "now_ms": lambda: clock[0]})
    assert len(t3) == 901
    # context_length_exceeded -> re-chunk at half the token count (embeddings.ex:159-164)
    state = {"n": 0}

    def picky(text):
        state["n"] += 1
        return ("error", "context_length_exceeded") if len(text.split()) > 4 else ("ok", [1.0, 2.0])

    r = embeddings.get_embedding("a b c d e f g h", {"embedding_fn": picky, "cache_table": {}})
    # 8 one-token words, limit div(8,2)=4, +1 per separator: "a b" = 3, "a b c" = 5 > 4 -> 4 chunks of 2 words
    assert r[0] == "ok" and r[1]["chunks"] == 4


def test_sha_test_embedding_bridge():
    """helpers.ex:69-76: sha256(text)/255 -> 32 dims; the reproducible bridge for BASELINE config 1."""
    info = GOLD["sha_embedding"]
    embs = [embeddings.sha_test_embedding(t)[1]["embedding"] for t in info["texts"]]
    assert all(len(e) == info["dim"] and all(0.0 <= x <= 1.0 for x in e) for e in embs)
    m = numeric.cosine_matrix(embs)
    assert m[0][0] == pytest.approx(1.0, abs=1e-12) and 0.0 < m[0][1] < 1.0
    ok, med, cos, _ = numeric.decide_batch(np.array([embs]), np.array([0.8]))
    assert cos.shape == (1, 3, 3) and np.allclose(cos[0], np.array(m), atol=1e-12)
    assert med[0] == numeric.medoid_vote(m)
    assert bool(ok[0]) == numeric.reference_rule(embs, 0.8)


def test_numpy_batch_matches_scalar_oracle_and_c_port():
    rng = np.random.default_rng(3)
    emb = rng.standard_normal((20, 5, 64)).astype(np.float32)
    emb[3, 2] = 0.0                                            # zero vector inside a pool
    thr = np.full(20, 0.1)
    ok, med, cos, _ = numeric.decide_batch(emb, thr)
    c_ok, c_med, c_cos = cport.similarity_batch(emb, thr, threads=2)
    assert np.array_equal(ok, c_ok) and np.array_equal(med, c_med)
    assert np.max(np.abs(cos - c_cos)) < 1e-12
    for b in range(20):
        e = emb[b].astype(np.float64).tolist()
        m = numeric.cosine_matrix(e)
        assert np.max(np.abs(np.array(m) - cos[b])) < 1e-12
        assert numeric.medoid_vote(m) == med[b]
        assert numeric.reference_rule(e, 0.1) == bool(ok[b])
    assert np.all(cos[3, 2] == 0.0)


@pytest.mark.parametrize("case", GOLD["batch_merge"], ids=lambda c: c["src"].split(":")[-1])
def test_batch_sequence_merge_known_answers(case):
    """The reference's own :batch_sequence_merge tests (consensus_rules_batch_test.exs) against the oracle AND the
    product's two-pass restatement (quoracle_b200.consensus: plan on the host, semantic merges through the collector)."""
    from oracle import consensus_rules as orules
    from quoracle_b200 import consensus
    from quoracle_b200.collector import DecisionCollector
    from quoracle_b200.embeddings import Embeddings
    from quoracle_b200 import synth
    from tests.fake_ctx import FakeContext

    def check(r):
        if case["expect"] == "ok":
            assert r[0] == "ok" and len(r[1]) == case["length"], r
            if "actions" in case:
                assert [a["action"] for a in r[1]] == case["actions"]
            if "merged" in case:
                assert r[1] == case["merged"]
        else:
            assert list(r) == case["expect"], r

    check(orules.apply_rule(("batch_sequence_merge",), case["sequences"]))
    w = synth.make_encoder(64, vocab=512, seed=1)
    ctx = FakeContext(w.table_bf16, w.w_bf16)
    col = DecisionCollector(Embeddings(ctx, vocab=512))
    state, payload = consensus._plan_batch_merge(col, "agent", "batch_sync", case["sequences"])
    results = col.flush()
    check(payload if state == "done" else consensus._finish_batch_merge(payload, results, "agent"))
    assert ctx.calls["embed"] == 0                       # identical semantic values never reach the encoder


def test_batch_sync_clusters_in_result_known_answers():
    """result_batch_sync_test.exs: a batch cluster ranks by the MAX priority of its sequence; structural merge errors
    propagate out of merge_cluster_params.  Oracle and product."""
    from quoracle_b200 import consensus
    for case in GOLD["batch_result"]["break_tie"]:
        for winner in (result.break_tie(case["clusters"]), consensus.pick_winner(case["clusters"], 4)[1]):
            if "expect_action" in case:
                assert winner["representative"]["action"] == case["expect_action"], case["src"]
            else:
                assert winner["representative"]["params"]["actions"][0]["params"]["path"] == case["expect_first_path"]
    case = GOLD["batch_result"]["merge_error"]
    assert list(result.merge_cluster_params(case["cluster"])) == case["expect"]
    seqs = [a["params"]["actions"] for a in case["cluster"]["actions"]]
    state, payload = consensus._plan_batch_merge(None, "agent", "batch_sync", seqs)   # decided before any collector use
    assert state == "done" and list(payload) == case["expect"]


def test_batch_fingerprints_and_priorities_known_answers():
    """batch_consensus_test.exs:26-137,209-255 and batch_async_consensus_test.exs:33-236,309-418: batch_sync
    fingerprints are the ORDERED list of sub-action types, batch_async the SORTED one, the two never cluster together;
    an empty or missing `actions` list ranks with the default priority 999.  Oracle and product."""
    from quoracle_b200 import consensus

    def batch(kind, types):
        return {"action": kind, "params": {"actions": [{"action": t, "params": {}} for t in types]}, "reasoning": ""}

    for fp, cluster_fn in ((aggregator.action_fingerprint, aggregator.cluster_responses),
                           (consensus.fingerprint, consensus.cluster)):
        norm = lambda x: (x[0], list(x[1]))                                       # noqa: E731
        assert norm(fp(batch("batch_sync", ["file_read", "todo"]))) == ("batch_sync", ["file_read", "todo"])
        assert norm(fp(batch("batch_async", ["todo", "file_read"]))) == ("batch_async", ["file_read", "todo"])
        assert norm(fp(batch("batch_async", ["todo", "file_read", "fetch_web"]))) == \
            ("batch_async", ["fetch_web", "file_read", "todo"])
        assert fp(batch("batch_sync", ["file_read", "todo"]))[0] != fp(batch("batch_async", ["file_read", "todo"]))[0]
        # same actions, different order: one cluster for batch_async, two for batch_sync
        assert len(cluster_fn([batch("batch_async", ["file_read", "todo"]), batch("batch_async", ["todo", "file_read"])])) == 1
        assert len(cluster_fn([batch("batch_sync", ["file_read", "todo"]), batch("batch_sync", ["todo", "file_read"])])) == 2
        assert len(cluster_fn([batch("batch_async", ["file_read", "todo"]), batch("batch_async", ["file_read", "orient"])])) == 2
        assert norm(fp({"action": "batch_sync", "params": {}, "reasoning": ""})) == ("batch_sync", [])
    for kind in ("batch_sync", "batch_async"):
        for rep in (batch(kind, []), {"action": kind, "params": {}, "reasoning": ""}):
            c = {"representative": rep, "actions": [rep], "count": 1}
            assert result.calculate_cluster_priority(c) == 999 and consensus._priority(c) == 999
        c = {"representative": batch(kind, ["orient", "file_read"]), "actions": [], "count": 1}
        assert result.calculate_cluster_priority(c) == consensus._priority(c) == 6          # max(orient 1, file_read 6)


def test_round_penalty_uses_max_refinement_rounds_known_answers():
    """max_refinement_rounds_test.exs:178-236: the -0.1 per round penalty counts rounds beyond `max_refinement_rounds`
    (default 4): 2-of-3 at round 5 -> 2/3 + 0.10 - 0.10 = 0.667; max 0 penalises from round 1; the option threads
    through format_result.  Oracle and product."""
    from quoracle_b200 import consensus
    action = {"action": "orient", "params": {"current_situation": "test", "goal_clarity": "test", "available_resources": "test",
                                            "key_challenges": "test", "delegation_consideration": "none"},
              "reasoning": "test reasoning"}
    cluster = {"count": 2, "actions": [action, action], "representative": action}

    def conf(rnd, mx=None):
        opts = {} if mx is None else {"max_refinement_rounds": mx}
        o = result.format_result([cluster], 3, rnd, opts)[2]["confidence"]
        p = consensus.confidence(2, 3, rnd, 4 if mx is None else mx)
        assert abs(o - p) < 1e-12
        return o

    assert abs(conf(5) - 0.667) < 0.001
    assert conf(1, 0) < conf(1, 4)
    assert conf(6, 2) < conf(6, 8)
