"""CPU tests of the host-side mirror (tokenizer, Embeddings, consensus rule, collector, lesson manager,
sharding maths) with a fake context whose arithmetic is the oracle's."""
import numpy as np
import pytest

from oracle import consensus_rules as orules
from oracle import embeddings as oemb
from quoracle_b200 import consensus_rules, lesson_manager, sharding, synth, tokenizer
from quoracle_b200.collector import DecisionCollector
from quoracle_b200.embeddings import Embeddings, chunk_text_by_tokens
from tests.fake_ctx import FakeContext

W = synth.make_encoder(384, vocab=4096, seed=7)


def make_emb(**kw):
    ctx = FakeContext(W.table_bf16, W.w_bf16)
    return Embeddings(ctx, vocab=4096, **kw), ctx


def test_tokenizer_is_fnv1a_and_stable():
    assert tokenizer.fnv1a32(b"") == 0x811C9DC5
    assert tokenizer.fnv1a32(b"a") == 0xE40C292C
    assert tokenizer.fnv1a32(b"foobar") == 0xBF9CF968
    ids, lens = tokenizer.pack(["deploy the build", "", "one"], vocab=4096)
    assert lens.tolist() == [3, 0, 1] and ids.shape == (3, 3)
    assert ids[0].tolist() == tokenizer.tokenize("deploy the build", 4096) and (ids[1] == -1).all()
    assert all(0 <= t < 4096 for t in ids[0])


def test_chunker_matches_the_oracle_restatement():
    text = " ".join(f"w{i}" for i in range(57))
    for limit in (1, 4, 9, 56, 57, 100):
        assert chunk_text_by_tokens(text, limit) == oemb.chunk_text_by_tokens(text, limit)


def test_embeddings_cache_dedup_chunks_and_shapes():
    emb, ctx = make_emb(model_token_limit=10)
    r = emb.get_embeddings(["alpha beta", "alpha beta", "", "gamma", {"k": [1, 2]}])
    assert ctx.calls["embed"] == 1                                  # one GPU call for the whole batch
    assert r[2] == ("error", "invalid_input")
    assert r[0][0] == "ok" and r[0][1]["cached"] is False and np.array_equal(r[0][1]["embedding"], r[1][1]["embedding"])
    again = emb.get_embeddings(["alpha beta"])
    assert again[0][1]["cached"] is True and ctx.calls["embed"] == 1
    long_text = " ".join(f"w{i}" for i in range(40))
    lr = emb.get_embedding(long_text)
    assert lr[0] == "ok" and lr[1]["chunks"] > 2 and ctx.calls["chunk_mean"] == 1
    # the chunk mean is unweighted and NOT re-normalised (embeddings.ex:371-383)
    assert np.linalg.norm(lr[1]["embedding"]) < 1.0 + 1e-6
    # accumulator shape: 3-tuple when not cached, 2-tuple when cached
    acc = orules.Accumulator()
    a = emb.get_embedding("fresh text", {"cost_accumulator": acc})
    b = emb.get_embedding("fresh text", {"cost_accumulator": acc})
    assert len(a) == 3 and a[2].count() == 1 and len(b) == 2 and b[1]["cached"] is True
    assert emb.get_embedding("bypass", {"cache_ttl": 0})[1]["cached"] is False
    assert emb.get_embedding("bypass", {"cache_ttl": 0})[1]["cached"] is False


def test_rule_mirror_agrees_with_the_oracle_rule_for_the_same_embedding_function():
    emb, ctx = make_emb()
    fn = emb.embedding_fn()
    ofn = lambda v, o: (lambda r: ("ok", {"embedding": r[1]["embedding"].astype(np.float64).tolist()}) if r[0] == "ok" else r)(fn(v, {}))  # noqa: E731
    cases = [(["same", "same", "same"], 0.9), (["deploy build now", "deploy build now please"], 0.5),
             (["red green blue", "quantum flux capacitor"], 0.999999), ([], 0.9), (["only"], 0.9),
             (["a b c d", "a b c e", "a b c f", "zzz yyy"], 0.9999)]
    for values, thr in cases:
        want = orules.apply_rule(("semantic_similarity", thr), values, {"embedding_fn": ofn})
        got = consensus_rules.apply_rule(emb, values, thr)
        if values == []:
            assert got == ("error", "no_values") and want == ("error", "no_consensus")   # private vs public mapping
        else:
            assert got == want, (values, thr)
    before = ctx.calls["embed"]
    assert consensus_rules.apply_rule(emb, ["x", "x"], 0.9) == ("ok", "x") and ctx.calls["embed"] == before


def test_collector_batches_a_whole_tree_into_one_call():
    emb, ctx = make_emb()
    col = DecisionCollector(emb)
    for a in range(20):
        col.submit(f"agent-{a}", "content", [f"send report {a}", f"send the report {a}", f"send report {a} now"], 0.5)
        col.submit(f"agent-{a}", "task_description", ["same", "same", "same"], 0.95)
    out = col.flush()
    assert ctx.calls["embed"] == 1 and ctx.calls["similarity"] == 1
    assert len(out) == 20 and out["agent-3"]["task_description"]["result"] == ("ok", "same")
    assert out["agent-3"]["content"]["result"][0] in ("ok", "error") and out["agent-3"]["content"]["medoid"] in (0, 1, 2)
    assert col.flush() == {}


def test_lesson_manager_mirror():
    emb, ctx = make_emb()
    ex = [{"type": "factual", "content": "API requires bearer auth", "confidence": 2}]
    ok, out = lesson_manager.accumulate_lessons(emb, ex, [{"type": "factual", "content": "API requires bearer auth", "confidence": 1}])
    assert out == [{"type": "factual", "content": "API requires bearer auth", "confidence": 3}]
    ok, out = lesson_manager.accumulate_lessons(emb, ex, [{"type": "factual", "content": "zebra xylophone quartz", "confidence": 1}],
                                                threshold=0.9999)
    assert len(out) == 2
    assert [l["confidence"] for l in lesson_manager.prune_lessons(
        [{"content": str(i), "confidence": c} for i, c in enumerate([1, 5, 2])], 2)] == [5, 2]


def test_sharding_maths():
    for n, w in [(256, 1), (4096, 8), (10, 3), (7, 8)]:
        blocks = [sharding.shard_range(n, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
        assert max(hi - lo for lo, hi in blocks) - min(hi - lo for lo, hi in blocks) <= 1
    order = sharding.tree_agent_order(4, 8)
    assert len(order) == 8 + 64 + 512 + 4096 and order[0] == (0,) and order[1] == (0, 0)
    assert sum(1 for p in order if len(p) == 4) == 4096                      # BASELINE config 3 leaves
    parts = [sharding.shard_tree(4, 8, r, 8) for r in range(8)]
    assert sum(len(p) for p in parts) == len(order)
    assert all(p[0][0] == r for r, p in enumerate(parts))                     # a top-level subtree per GPU


def test_schema_data_is_in_sync_with_the_oracle_tables():
    import json
    from pathlib import Path
    from oracle import schema_tables as st
    data = json.loads((Path(__file__).parent.parent / "quoracle_b200" / "data" / "action_schemas.json").read_text())
    assert data["priorities"] == st.ACTION_PRIORITIES
    assert set(data["schemas"]) == set(st.SCHEMAS)
    for a, s in st.SCHEMAS.items():
        assert data["schemas"][a]["required"] == s["required"] and data["schemas"][a]["optional"] == s["optional"]
        assert {p: tuple(r) for p, r in data["schemas"][a]["rules"].items()} == s["rules"]


def _random_proposals(rng, n):
    """N proposals over a few actions with overlapping / conflicting params, like the reference's mock scenarios."""
    texts = ["analyze the customer data for patterns", "Analyze customer data patterns!", "deploy release to production",
             "summarise quarterly revenue figures please", "analyze the customer data for patterns"]
    out = []
    for _ in range(n):
        kind = int(rng.integers(0, 4))
        wait = [True, False, None, 5, 0][int(rng.integers(0, 5))]
        if kind == 0:
            r = {"action": "send_message", "params": {"to": ["parent", "child"][int(rng.integers(0, 2))],
                                                      "content": texts[int(rng.integers(0, len(texts)))]}}
        elif kind == 1:
            r = {"action": "fetch_web", "params": {"url": "https://example.com", "timeout": int(rng.integers(1, 60)),
                                                   "follow_redirects": bool(rng.integers(0, 2))}}
        elif kind == 2:
            r = {"action": "orient", "params": {"current_situation": texts[int(rng.integers(0, len(texts)))],
                                                "goal_clarity": texts[int(rng.integers(0, 2))],
                                                "available_resources": "cpu gpu disk", "key_challenges": "latency",
                                                "delegation_consideration": "none needed"}}
        else:
            r = {"action": "spawn_child", "params": {"task_description": texts[int(rng.integers(0, len(texts)))],
                                                     "success_criteria": "tests pass", "immediate_context": "ctx",
                                                     "approach_guidance": "carefully", "profile": "default",
                                                     "skills": [["a", "b"], ["b", "c"], ["c"]][int(rng.integers(0, 3))]}}
        r["reasoning"] = ["", "because", "since"][int(rng.integers(0, 3))]
        if wait is not None:
            r["wait"] = wait
        out.append(r)
    return out


def _random_batch_proposals(rng, n):
    """N batch proposals (batch_sync / batch_async) whose sub-action sequences mostly line up: same types per
    position with parameters that agree, nearly agree (semantic rule) or conflict (exact-match rule -> the
    mode-merge fallback); a few with a different length or a different type at one position."""
    texts = ["analyze the customer data for patterns", "Analyze customer data patterns!", "deploy release to production"]
    kind = ["batch_sync", "batch_async"][int(rng.integers(0, 2))]
    out = []
    for _ in range(n):
        seq = [{"action": "send_message", "params": {"to": ["parent", "child"][int(rng.integers(0, 4) == 0)],
                                                     "content": texts[int(rng.integers(0, len(texts)))]}},
               {"action": "fetch_web", "params": {"url": ["https://example.com", "https://example.org"][int(rng.integers(0, 5) == 0)],
                                                  "timeout": int(rng.integers(1, 60))}}]
        roll = int(rng.integers(0, 12))
        if roll == 0:
            seq.append({"action": "orient", "params": {"current_situation": texts[0], "goal_clarity": texts[1],
                                                       "available_resources": "cpu", "key_challenges": "latency",
                                                       "delegation_consideration": "none"}})
        elif roll == 1:
            seq[1] = {"action": "file_read", "params": {"path": "/tmp/x"}}
        if kind == "batch_async" and int(rng.integers(0, 6)) == 0:
            seq = seq[::-1]                               # same cluster (sorted type list), positions no longer line up
        r = {"action": kind, "params": {"actions": seq}, "reasoning": ["", "because"][int(rng.integers(0, 2))]}
        w = [True, False, None, 3][int(rng.integers(0, 4))]
        if w is not None:
            r["wait"] = w
        out.append(r)
    return out


def test_batched_decide_step_matches_the_oracle_format_result():
    """quoracle_b200.consensus.decide_tree == oracle Aggregator.cluster_responses + Result.format_result, decision
    by decision, when both see the same embedding function -- and the whole tree costs one embed call."""
    from oracle import aggregator as oagg
    from oracle import result as ores
    from quoracle_b200 import consensus
    emb, ctx = make_emb()
    fn = emb.embedding_fn()
    ofn = lambda v, o: (lambda r: ("ok", {"embedding": r[1]["embedding"].astype(np.float64).tolist()}) if r[0] == "ok" else r)(fn(v, {}))  # noqa: E731
    rng = np.random.default_rng(11)
    pending = [{"agent_id": f"a{i}", "responses": _random_proposals(rng, int(rng.integers(2, 9)))} for i in range(400)]
    pending += [{"agent_id": f"b{i}", "responses": _random_batch_proposals(rng, int(rng.integers(2, 7)))} for i in range(200)]
    for round_num in (1, 2, 6):
        emb2, ctx2 = make_emb()
        got = consensus.decide_tree(DecisionCollector(emb2), pending, round_num=round_num)
        assert ctx2.calls["embed"] <= 1
        for d, g in zip(pending, got):
            clusters = oagg.cluster_responses(d["responses"])
            maj = oagg.find_majority_cluster(clusters, len(d["responses"]), round_num)
            if maj[0] == "no_majority" and round_num <= 4:
                assert g["kind"] == "no_majority"
                continue
            kind, action, opts = ores.format_result(clusters, len(d["responses"]), round_num, {"embedding_fn": ofn})
            assert g["kind"] == kind, (d, g)
            if action == "error":                        # the reference's {:error, reason} read as {action, acc}
                assert g["action"]["action"] == "error" and g["action"]["error"] == opts.get("accumulator")
                continue
            assert g["action"]["action"] == action["action"]
            assert g["action"]["params"] == action["params"], (g["action"], action)
            assert g["action"]["wait"] == action["wait"] and g["action"]["reasoning"] == action["reasoning"]
            assert abs(g["confidence"] - opts["confidence"]) < 1e-12


def test_rechunk_at_half_when_the_encoder_rejects_a_text_within_the_limit(monkeypatch):
    """embeddings.ex:159-164: a text within the token limit that the provider still rejects with
    :context_length_exceeded is re-chunked at half its own token count; chunks that still fail are dropped and
    nothing left means :all_chunks_failed (:170-192).  The device's context limit plays the provider's part; the
    oracle is driven with a provider that rejects the same texts."""
    import quoracle_b200.embeddings as pe
    monkeypatch.setattr(pe, "QC_MAX_TOKENS", 12)                  # a small "context window" keeps the test light
    emb, ctx = make_emb(model_token_limit=100)                    # effective limit 90 > 12: the estimate says "fits"

    def provider(text):                                           # what the oracle's embedding_fn sees
        if tokenizer.count_tokens(text) > 12:
            return ("error", "context_length_exceeded")
        ids, lens = tokenizer.pack([text], 4096)
        return ("ok", ctx.embed_batch(ids, lens)[0].astype(np.float64).tolist())   # 1-arity seam: {:ok, [float]}

    texts = [" ".join(f"w{i}" for i in range(n)) for n in (5, 12, 13, 20, 24, 25, 60)]
    got = emb.get_embeddings(texts, cache_ttl=0)
    for text, g in zip(texts, got):
        want = oemb.get_embedding(text, {"embedding_fn": provider, "model_token_limit": 100, "cache_ttl": 0,
                                         "token_counter": tokenizer.count_tokens, "table": {}})
        assert g[0] == want[0], (tokenizer.count_tokens(text), g, want)
        if g[0] == "ok":
            assert g[1]["chunks"] == want[1]["chunks"]
            assert np.max(np.abs(g[1]["embedding"] - np.array(want[1]["embedding"]))) < 1e-6
        else:
            assert g[1] == want[1]
    kinds = [g[0] if g[0] == "ok" else g[1] for g in got]
    assert kinds[:3] == ["ok", "ok", "ok"] and "all_chunks_failed" in kinds     # 60 tokens -> halves of 30 still too long


def test_cache_ttl_and_eviction_follow_the_reference():
    """embeddings.ex:65-95 (hit only while now - ts < ttl, expired entries deleted on lookup) and :402-426 (at >= 1000
    entries the 100 oldest go before the insert)."""
    from quoracle_b200.embeddings import EmbeddingCache, MAX_CACHE_ENTRIES
    c = EmbeddingCache()
    v = np.ones(4, np.float32)
    for i in range(MAX_CACHE_ENTRIES):
        c.insert(i.to_bytes(4, "big"), v, now_ms=1000 + i, ttl=50_000)
    assert len(c.table) == MAX_CACHE_ENTRIES
    c.insert(b"new!", v, now_ms=5000, ttl=50_000)
    assert len(c.table) == MAX_CACHE_ENTRIES - 100 + 1
    assert (0).to_bytes(4, "big") not in c.table and (99).to_bytes(4, "big") not in c.table
    assert (100).to_bytes(4, "big") in c.table
    assert c.lookup(b"new!", 5000 + 49_999) is not None
    assert c.lookup(b"new!", 5000 + 50_000) is None and b"new!" not in c.table      # expired and removed
    # the same sequence through the oracle's table gives the same survivors
    table = {}
    now = [0]
    for i in range(MAX_CACHE_ENTRIES):
        now[0] = 1000 + i
        oemb.get_embedding(f"t{i}", {"embedding_fn": lambda t: ("ok", [1.0]), "cache_table": table, "cache_ttl": 50_000,
                                     "now_ms": lambda: now[0], "token_counter": tokenizer.count_tokens})
    now[0] = 5000
    oemb.get_embedding("new", {"embedding_fn": lambda t: ("ok", [1.0]), "cache_table": table, "cache_ttl": 50_000,
                               "now_ms": lambda: now[0], "token_counter": tokenizer.count_tokens})
    assert len(table) == MAX_CACHE_ENTRIES - 100 + 1             # c.table lost b"new!" to the expiry check above


def test_host_rules_agree_with_the_oracle_on_random_terms():
    """Differential check of every rule that needs no embeddings (exact_match, mode_selection, union_merge,
    structural_merge, percentile) on random nested terms: the product's restatement of Erlang term order and of the
    rule arithmetic must give what the oracle's gives (ties between compound values are decided by term order)."""
    import json
    from quoracle_b200 import consensus
    rng = np.random.default_rng(0)

    def rand_val(depth=0):
        k = int(rng.integers(0, 8))
        if k == 0:
            return int(rng.integers(-5, 50))
        if k == 1:
            return float(rng.integers(-5, 50)) / 2
        if k == 2:
            return bool(rng.integers(0, 2))
        if k == 3:
            return ["a", "b", "deploy now", "x y"][int(rng.integers(0, 4))]
        if k == 4 and depth < 2:
            return [rand_val(depth + 1) for _ in range(int(rng.integers(0, 4)))]
        if k == 5 and depth < 2:
            return {["k", "j", "m"][int(rng.integers(0, 3))]: rand_val(depth + 1) for _ in range(int(rng.integers(0, 3)))}
        return int(rng.integers(0, 4))

    rules = [("exact_match",), ("mode_selection",), ("union_merge",), ("structural_merge",), ("percentile", 50),
             ("percentile", 75), ("percentile", 100), ("percentile", 10)]
    for _ in range(6000):
        rule = rules[int(rng.integers(0, len(rules)))]
        vals = [rand_val() for _ in range(int(rng.integers(1, 7)))]
        if rule[0] == "structural_merge":
            vals = [v if isinstance(v, dict) else {"k": v} for v in vals]
        want = orules.apply_rule(rule, vals)
        got = consensus.host_rule(list(rule), vals)
        assert want[0] == got[0], (rule, vals, want, got)
        if want[0] == "ok":
            assert type(want[1]) is type(got[1]) and json.dumps(want[1], sort_keys=True) == json.dumps(got[1], sort_keys=True), \
                (rule, vals, want, got)
        else:
            assert want == got


def test_wait_rule_confidence_and_key_terms_agree_with_the_oracle():
    """Differential checks of the three small pieces of Result / Aggregator arithmetic the product restates:
    the wait-parameter merge (consensus_rules.ex:114-160 over result.ex:240-248; waits are true | false | integer),
    calculate_confidence/4 (result.ex:261-286) and the key-term normalisation (aggregator.ex:423-454)."""
    from oracle import aggregator as oagg
    from oracle import result as ores
    from quoracle_b200 import consensus
    rng = np.random.default_rng(1)
    for _ in range(5000):
        vals = [[True, False, 0, 1, 5, 30, 60, 7, 2][int(rng.integers(0, 9))] for _ in range(int(rng.integers(1, 8)))]
        want, got = ores._merge_wait_parameter(list(vals)), consensus.wait_rule(list(vals))
        assert want == got and type(want) is type(got), (vals, want, got)
    for _ in range(2000):
        total = int(rng.integers(1, 12))
        count, rnd, mx = int(rng.integers(1, total + 1)), int(rng.integers(1, 9)), int(rng.integers(1, 6))
        want = ores.calculate_confidence({"count": count}, total, rnd, {"max_refinement_rounds": mx})
        assert abs(want - consensus.confidence(count, total, rnd, mx)) < 1e-12
    words = ["Deploy", "the", "build,", "to", "PRODUCTION!", "now", "please", "naïve", "x", "yyyy", "zzzzz", "a-b-c-d",
             "foo_bar", "12345", "tab\\tsep", "new\\nline"]
    for _ in range(5000):
        s = " ".join(words[int(rng.integers(0, len(words)))] for _ in range(int(rng.integers(0, 9))))
        thr = [0.8, 0.9, 0.95, 0.99][int(rng.integers(0, 4))]
        assert oagg.normalize_semantic_string(s, thr) == consensus._key_terms(s, thr), (s, thr)


def test_equal_maps_embed_the_same_text_whatever_their_key_order():
    """Jason.encode!/1 (embeddings.ex:46) walks a map in BEAM order, so equal maps give one text: same cache entry, one
    row on the GPU, and a pool of equal maps short-circuits without any embedding call (consensus_rules.ex:251-256)."""
    emb, ctx = make_emb()
    a, b = {"path": "/tmp/x", "mode": "r"}, {"mode": "r", "path": "/tmp/x"}
    r = emb.get_embeddings([a, b, {"path": "/tmp/y"}])
    assert ctx.calls["embed"] == 1 and np.array_equal(r[0][1]["embedding"], r[1][1]["embedding"])
    ofn = lambda t: ("ok", [float(len(t)), float(sum(t.encode()))])          # noqa: E731 - depends on the text only
    assert oemb.get_embedding(a, {"embedding_fn": ofn, "cache_ttl": 0})[1]["embedding"] == \
        oemb.get_embedding(b, {"embedding_fn": ofn, "cache_ttl": 0})[1]["embedding"]
    res, _ = consensus_rules.decide_many(emb, [([a, b, dict(a)], 0.9)])
    assert res == [("ok", a)] and ctx.calls["embed"] == 1                  # no further embedding call


def test_semantic_similarity_mirror_agrees_with_the_oracle():
    """Aggregator.calculate_semantic_similarity/3 (aggregator.ex:249-289): float for a pair, 0.0 when an embedding is
    unavailable, `(float, acc)` with an accumulator; many pairs cost one embedding call."""
    from oracle import aggregator as oagg
    from quoracle_b200 import aggregator as pagg
    emb, ctx = make_emb()
    fn = emb.embedding_fn()
    ofn = lambda t: (lambda r: ("ok", r[1]["embedding"].astype(np.float64).tolist()) if r[0] == "ok" else r)(fn(t, {}))  # noqa: E731
    pairs = [("deploy the build", "deploy the new build"), ("alpha", "alpha"), ("", "x"), ({"k": 1}, {"k": 1}),
             ("summarise revenue figures", "zebra quartz")]
    emb2, ctx2 = make_emb()
    got = pagg.semantic_similarity_many(emb2, pairs)
    assert ctx2.calls["embed"] == 1 and ctx2.calls["similarity"] == 1
    for (a, b), g in zip(pairs, got):
        sa, sb = (x if isinstance(x, str) else oemb.json.dumps(x, separators=(",", ":"), sort_keys=True) for x in (a, b))
        want = oagg.calculate_semantic_similarity(sa, sb, {"embedding_fn": ofn}) if sa and sb else 0.0
        assert abs(g - want) < 1e-6, (a, b, g, want)
    assert got[1] == pytest_approx_one(got[1]) and got[2] == 0.0
    acc = orules.Accumulator()
    emb3, _ = make_emb()
    sim, acc2 = pagg.calculate_semantic_similarity(emb3, "first text", "second text", acc)
    assert isinstance(sim, float) and acc2.count() == 2
    sim_b, acc3 = pagg.calculate_semantic_similarity(emb3, "first text", "second text", acc2)
    assert sim_b == sim and acc3.count() == 2                               # both cached: nothing accumulates
    assert pagg.semantic_similarity_with_embeddings(emb3, "first text", "first text", 0.999999) is True
    assert pagg.semantic_similarity_with_embeddings(emb3, "first text", "", 0.0) is True     # 0.0 >= 0.0, as in the reference
    assert pagg.semantic_similarity_with_embeddings(emb3, "first text", "", 0.1) is False


def pytest_approx_one(x):
    assert abs(x - 1.0) < 1e-6
    return x


def test_packed_stream_plan_invariants():
    """The host-side planner behind the packed token stream (qc_debug_pack_plan: same code the host entry points run,
    no device needed).  Checked from first principles for 32-, 16- and 8-token chunks: every valid token is covered
    exactly once and in order, chunks never mix candidates, a partial row never crosses a block, a candidate's rows
    are consecutive, padding chunks trail and point nowhere."""
    from quoracle_b200 import native
    rng = np.random.default_rng(4)
    for tb, ch in ((128, 32), (64, 32), (48, 16), (56, 8), (104, 8), (112, 16)):
        nch = tb // ch
        for trial in range(30):
            r, t = int(rng.integers(1, 40)), int(rng.integers(1, 400))
            lens = rng.integers(0, t + 1, size=r).astype(np.int32)
            if trial % 5 == 0:
                lens[rng.integers(0, r)] = 0
            if trial % 7 == 0:
                lens[:] = 0
            plan = native.debug_pack_plan(lens, t, tb)
            src, cnt, crow, row = plan["chunk_src"], plan["chunk_n"], plan["chunk_row"], plan["row_start"]
            n = src.shape[0]
            assert n % nch == 0 and n == -(-int(np.sum(-(-lens // ch))) // nch) * nch
            k = 0
            for cand in range(r):
                nchunks = -(-int(lens[cand]) // ch)
                assert row[cand + 1] - row[cand] == (0 if nchunks == 0 else len(set(crow[k:k + nchunks].tolist())))
                for j in range(nchunks):
                    assert src[k] == cand * t + j * ch and cnt[k] == min(ch, lens[cand] - j * ch)
                    assert row[cand] <= crow[k] < row[cand + 1]
                    if j > 0:                                   # same row iff same block, rows ascend by one
                        same_block = (k // nch) == ((k - 1) // nch)
                        assert crow[k] == crow[k - 1] + (0 if same_block else 1)
                    k += 1
            assert np.all(cnt[k:] == 0) and np.all(crow[k:] == -1)
            assert row[0] == 0 and row[r] == plan["nruns"] == (int(crow[:k].max()) + 1 if k else 0)
            assert np.all(np.diff(crow[:k]) >= 0) and np.all(np.diff(crow[:k]) <= 1)
    with pytest.raises(native.QcError):
        native.debug_pack_plan(np.array([5], np.int32), 4, 128)          # len > T


def test_lesson_store_equals_the_oracle_over_random_histories():
    """LessonStore (rows resident in a lesson cache, one query row per dedup) against the oracle's
    accumulate_lessons over random sequences of calls: merges (incl. duplicate contents), new lessons, lessons whose
    embedding fails (empty content), and pruning."""
    from oracle import lesson_manager as olm
    emb, ctx = make_emb()
    words = ["alpha", "beta", "gamma", "delta", "epsilon", "zeta", "eta", "theta", "iota", "kappa"]
    rng = np.random.default_rng(12)

    def ofn(text):
        r = emb.get_embedding(text)
        return ("ok", {"embedding": r[1]["embedding"].astype(np.float64).tolist()}) if r[0] == "ok" else r

    for trial in range(6):
        thr = float(rng.choice([0.9, 0.97, 0.999]))
        max_lessons = int(rng.choice([5, 8, 100]))
        store = lesson_manager.LessonStore(emb, threshold=thr, max_lessons=max_lessons)
        want = []
        for _ in range(12):
            news = []
            for _ in range(int(rng.integers(1, 4))):
                k = int(rng.integers(1, 4))
                content = " ".join(rng.choice(words, k)) if rng.random() > 0.08 else ""
                news.append({"type": "factual", "content": content, "confidence": int(rng.integers(1, 4))})
            _, want = olm.accumulate_lessons(want, news, {"embedding_fn": ofn, "similarity_threshold": thr,
                                                           "max_lessons": max_lessons})
            _, got = store.accumulate(news)
            assert got == want, (trial, news)
            assert len(store.cache) == sum(1 for r in store.row_of if r >= 0)
        store.close()


# ---- A10 / A11: the round loop and the decide step, batched over a tree vs the sequential oracle -------------------
def scripted_query_fn(seed: int, pool_size_of, p_agree=0.45, p_error=0.06, p_empty=0.05, p_invalid=0.05):
    """A deterministic `query_fn` (the :model_query_fn seam one level up): what the pool of agent `state["agent_id"]`
    answers at round `opts["round"]` depends only on (seed, agent, round, attempt), so the batched driver and the
    sequential oracle see the same LLM."""
    import zlib

    def fn(state, model_pool, opts):
        rnd = opts.get("round", 1)
        key = f"{seed}:{state['agent_id']}:{rnd}:{state.get('consensus_retry_count', 0)}".encode()
        rng = np.random.default_rng(zlib.crc32(key))
        roll = rng.random()
        new_state = dict(state, queries=state.get("queries", 0) + 1)
        n = pool_size_of(state["agent_id"])
        if roll < p_error:
            return ("error", "service_unavailable")
        if roll < p_error + p_empty:
            return ("ok", [None] * n, new_state)
        if roll < p_error + p_empty + p_invalid:
            return ("ok", [{"action": "invalid", "params": {}} for _ in range(n)], new_state)
        props = _random_proposals(rng, n)
        if rng.random() < p_agree:                                  # the pool converges: everyone proposes the same action
            props = [dict(props[0]) for _ in range(n)]
        if n > 2 and rng.random() < 0.2:
            props[int(rng.integers(0, n))] = None                    # one model's answer did not parse
        return ("ok", props, new_state)
    return fn


def _validate(r):
    return None if r["action"] == "invalid" else r


def _tree_agents(depth, fanout, limit=None):
    from quoracle_b200 import sharding
    paths = sharding.tree_agent_order(depth, fanout)
    if limit:
        paths = paths[:limit]
    agents = []
    for p in paths:
        aid = "/".join(map(str, p))
        agents.append({"agent_id": aid,
                       "state": {"agent_id": aid, "model_histories": {f"m{k}": [] for k in range(3 + (len(aid) % 6))},
                                 "max_refinement_rounds": [1, 2, 4][len(p) % 3],
                                 "force_reflection": False}})
    return agents


def _compare_driver(emb, agents, seed):
    from oracle import consensus_driver as odrv
    from quoracle_b200 import consensus_driver

    def ofn(value, opts):
        r = emb.get_embedding(value)
        return ("ok", {"embedding": r[1]["embedding"].astype(np.float64).tolist()}) if r[0] == "ok" else r

    pool = {a["agent_id"]: len(a["state"]["model_histories"]) for a in agents}
    qf = scripted_query_fn(seed, lambda i: pool[i])
    got = consensus_driver.run_consensus_rounds(DecisionCollector(emb), agents, qf, _validate)
    kinds = {}
    for a in agents:
        want = odrv.get_consensus_with_state(a["state"], {}, qf, _validate, embedding_fn=ofn)
        g = got[a["agent_id"]]
        assert g[0] == want[0], (a["agent_id"], g, want)
        if want[0] == "error":
            assert g[1] == want[1]
            kinds[want[1]] = kinds.get(want[1], 0) + 1
            continue
        kind, action, meta = want[1]
        assert g[1]["kind"] == kind, (a["agent_id"], g[1], want[1])
        assert g[2] == want[2]                                        # the state the last query left behind
        kinds[kind] = kinds.get(kind, 0) + 1
        if action == "error":
            assert g[1]["action"]["action"] == "error"
            continue
        assert g[1]["action"]["action"] == action["action"] and g[1]["action"]["params"] == action["params"]
        assert g[1]["action"]["wait"] == action["wait"] and g[1]["action"]["reasoning"] == action["reasoning"]
        assert abs(g[1]["confidence"] - meta["confidence"]) < 1e-12
        assert g[1]["meta"].get("forced_reflection_applied", False) == meta.get("forced_reflection_applied", False)
    return got, kinds


def test_round_loop_over_a_tree_equals_the_sequential_oracle():
    emb, ctx = make_emb()
    agents = _tree_agents(3, 6, limit=180)
    # a few single-model pools with forced reflection (consensus.ex:303-321)
    for a in agents[::17]:
        a["state"]["model_histories"] = {"solo": []}
        a["state"]["force_reflection"] = True
    got, kinds = _compare_driver(emb, agents, seed=7)
    assert kinds.get("consensus", 0) > 20 and kinds.get("forced_decision", 0) > 5
    assert kinds.get("all_models_failed", 0) + kinds.get("all_responses_invalid", 0) + kinds.get("service_unavailable", 0) > 3
    # every sweep is one batched flush: the rounds in flight bound the GPU calls, not the number of agents
    assert got["__sweeps__"] <= 6 and ctx.calls["embed"] <= got["__sweeps__"]


def test_decide_step_retries_at_most_three_times():
    """run_consensus_cycle/2 + handle_consensus_error/4 (message_handler.ex:125-148,353-389) for a whole tree vs the
    oracle's per-agent loop: retryable failures get a second and a third attempt, then stall; other errors fail."""
    from oracle import consensus_driver as odrv
    from quoracle_b200 import consensus_driver
    emb, ctx = make_emb()
    agents = _tree_agents(2, 8, limit=60)
    pool = {a["agent_id"]: len(a["state"]["model_histories"]) for a in agents}
    qf = scripted_query_fn(11, lambda i: pool[i], p_agree=0.5, p_error=0.05, p_empty=0.25, p_invalid=0.2)

    def ofn(value, opts):
        r = emb.get_embedding(value)
        return ("ok", {"embedding": r[1]["embedding"].astype(np.float64).tolist()}) if r[0] == "ok" else r

    def execute(state, action):
        return dict(state, executed=action["action"] if isinstance(action, dict) else action)

    got = consensus_driver.run_consensus_cycles(DecisionCollector(emb), agents, qf, execute, _validate)
    outcomes = {}
    for a in agents:
        def cfn(state):
            r = odrv.get_consensus_with_state(state, {}, qf, _validate, embedding_fn=ofn)
            return ("ok", r[1][1], r[2], None) if r[0] == "ok" else ("error", r[1], None)
        w_state, w_out = odrv.run_consensus_cycle(a["state"], cfn, execute)
        g_state, g_out = got[a["agent_id"]]
        assert g_out[0] == w_out[0], (a["agent_id"], g_out, w_out)
        outcomes[w_out[0]] = outcomes.get(w_out[0], 0) + 1
        if w_out[0] == "executed":
            assert g_out[1]["action"] == w_out[1]["action"] and g_out[1].get("params") == w_out[1].get("params")
            assert g_state["executed"] == w_state["executed"] and g_state["consensus_retry_count"] == 0
        else:
            assert g_out[1:] == w_out[1:]
            assert g_state.get("consensus_retry_count", 0) == w_state.get("consensus_retry_count", 0)
    assert outcomes.get("executed", 0) > 20 and outcomes.get("stalled", 0) >= 1 and outcomes.get("failed", 0) >= 1


def test_timed_collector_flushes_on_size_or_deadline():
    """A lone agent is answered after ~max_wait_ms (no batch ever fills); a burst of agents shares flushes."""
    import threading
    import time
    from quoracle_b200.collector import TimedCollector
    emb, ctx = make_emb()
    col = TimedCollector(emb, max_pending=8, max_wait_ms=20.0)
    try:
        t0 = time.perf_counter()
        r, med = col.decide(["send report now", "send the report now", "send report now please"], 0.5)
        lone = time.perf_counter() - t0
        assert r[0] in ("ok", "error") and col.flushes == 1
        assert 0.015 <= lone < 1.0                                   # waited for the deadline, not for 8 merges
        outs = [None] * 24
        def agent(i):
            outs[i] = col.decide([f"restart service {i % 3}", f"restart the service {i % 3}", "ship it"], 0.9)
        ts = [threading.Thread(target=agent, args=(i,)) for i in range(24)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        assert all(o is not None for o in outs)
        assert col.decided == 25 and 4 <= col.flushes <= 25          # bursts batch; size 8 forces at least 3 more flushes
        assert col.decide(["same", "same"], 0.9)[0] == ("ok", "same")   # identical values short-circuit, no embedding
    finally:
        col.close()
