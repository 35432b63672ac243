"""The host-side mirror on a real B200: the GPU rule / collector / lesson dedup agree with the oracle's
restatement of the reference rule evaluated on the same embedding vectors (what an `:embedding_fn` returns)."""
import numpy as np
import pytest

from oracle import consensus_rules as orules
from oracle import lesson_manager as olm
from quoracle_b200 import consensus_rules, lesson_manager, native, synth
from quoracle_b200.collector import DecisionCollector
from quoracle_b200.embeddings import Embeddings

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def emb():
    w = synth.make_encoder(768, vocab=32768, seed=synth.BASE_SEED + 9)
    ctx = native.Context(0, w.table_bf16, w.w_bf16)
    yield Embeddings(ctx, vocab=32768, model_token_limit=600)
    ctx.close()


def oracle_fn(emb):
    def fn(value, opts):
        r = emb.get_embedding(value)
        return ("ok", {"embedding": r[1]["embedding"].astype(np.float64).tolist()}) if r[0] == "ok" else r
    return fn


SENTENCES = ["deploy the new build to production tonight", "deploy new build to production this evening",
             "roll back the release and page the on-call engineer", "summarise the quarterly revenue figures",
             "deploy the new build to production tonight please", "ship it"]


def test_rule_on_gpu_matches_oracle_rule(emb):
    fn = oracle_fn(emb)
    rng = np.random.default_rng(0)
    for _ in range(40):
        n = int(rng.integers(2, 9))
        values = [SENTENCES[int(i)] for i in rng.integers(0, len(SENTENCES), n)]
        thr = float(rng.choice([0.8, 0.85, 0.9, 0.95, 0.99, 0.999]))
        want = orules.apply_rule(("semantic_similarity", thr), values, {"embedding_fn": fn})
        got = consensus_rules.apply_rule(emb, values, thr)
        # guard the f32/f64 threshold edge: recompute the margin with the oracle cosine
        if len(set(values)) > 1:
            e = [fn(v, {})[1]["embedding"] for v in values]
            margin = min(abs(orules.cosine_similarity(e[0], x) - thr) for x in e[1:])
            if margin < 2e-5:
                continue
        assert got == want, (values, thr)


def test_long_text_is_chunked_and_mean_is_unnormalised(emb):
    from oracle import embeddings as oemb
    long_text = " ".join(f"token{i % 97}" for i in range(1500))          # limit trunc(600 * 0.9) = 540
    want_chunks = len(oemb.chunk_text_by_tokens(long_text, 540))          # words cost 1 + 1 per separator -> 6
    r = emb.get_embedding(long_text, {"cache_ttl": 0})
    assert r[0] == "ok" and r[1]["chunks"] == want_chunks == 6
    assert np.linalg.norm(r[1]["embedding"]) <= 1.0 + 1e-5
    r2 = emb.get_embedding(long_text)
    r3 = emb.get_embedding(long_text)
    assert r3[1]["cached"] is True and np.array_equal(r2[1]["embedding"], r3[1]["embedding"])


def test_collector_one_gpu_call_for_the_tree(emb):
    col = DecisionCollector(emb)
    before = emb.ctx.launch_count()
    calls = emb.gpu_calls
    for a in range(64):
        col.submit(f"a{a}", "content", [f"{SENTENCES[a % 3]} run {a}", f"{SENTENCES[a % 3]} run {a} now", SENTENCES[3]], 0.9)
    out = col.flush()
    assert emb.gpu_calls == calls + 1                                     # one embed call for 192 values
    assert emb.ctx.launch_count() - before in (3, 4)                      # K1 + K2 + one K3 (+ the token pack kernel: texts are ragged)
    assert len(out) == 64 and all(v["content"]["medoid"] in (0, 1, 2) for v in out.values())


def test_lesson_dedup_matches_oracle(emb):
    fn1 = lambda t: oracle_fn(emb)(t, {})                                 # noqa: E731
    existing = [{"type": "factual", "content": s, "confidence": i + 1} for i, s in enumerate(SENTENCES[:4])]
    for new, thr in [(SENTENCES[4], 0.9), ("completely unrelated words about zebras", 0.9999), (SENTENCES[0], 0.9)]:
        lesson = {"type": "factual", "content": new, "confidence": 1}
        want = olm.deduplicate_lesson(lesson, existing, {"embedding_fn": fn1, "similarity_threshold": thr})
        got = lesson_manager.deduplicate_lesson(emb, lesson, existing, thr)
        assert got == want


def test_decide_tree_on_gpu_matches_oracle(emb):
    """The batched decide step over a small tree on the real GPU path vs oracle format_result.

    Scope: the oracle's `:embedding_fn` here returns the GPU's OWN embeddings (oracle_fn(emb)), so this checks the host
    logic and K3 (cosine + rule on those vectors) -- NOT K1.  K1's values are checked against the numpy f64 encoder in
    tests/test_gpu_encoder.py."""
    from oracle import aggregator as oagg
    from oracle import result as ores
    from quoracle_b200 import consensus
    from tests.test_host_logic import _random_batch_proposals, _random_proposals
    fn = oracle_fn(emb)
    rng = np.random.default_rng(5)
    pending = [{"agent_id": f"a{i}", "responses": _random_proposals(rng, int(rng.integers(3, 9)))} for i in range(40)]
    pending += [{"agent_id": f"b{i}", "responses": _random_batch_proposals(rng, int(rng.integers(2, 7)))} for i in range(30)]
    before = emb.ctx.launch_count()
    got = consensus.decide_tree(DecisionCollector(emb), pending, round_num=5)
    assert emb.ctx.launch_count() - before <= 4 + 7                      # pack + K1 + K2 + one K3 per pool size
    for d, g in zip(pending, got):
        clusters = oagg.cluster_responses(d["responses"])
        kind, action, opts = ores.format_result(clusters, len(d["responses"]), 5, {"embedding_fn": fn})
        assert g["kind"] == kind and g["action"]["action"] == action["action"]
        assert g["action"]["params"] == action["params"]


def test_http_facade_serves_the_library_embeddings(emb):
    """The OpenAI-compatible endpoint the unmodified reference can point its local-model route at
    (local_model_helper.ex:92-110): vectors are the library's, concurrent requests share GPU calls."""
    import json
    import threading
    import urllib.request

    from oracle import encoder as oenc
    from quoracle_b200 import tokenizer
    from quoracle_b200.http_facade import EmbeddingServer
    srv = EmbeddingServer(emb, max_wait_ms=250.0).start()   # generous window: the assertion is about coalescing, not latency
    try:
        texts = [f"agent {i}: restart service {i % 5} after the config change" for i in range(16)]
        out = [None] * len(texts)

        def worker(i):
            req = urllib.request.Request(srv.url + "/embeddings", data=json.dumps({"model": "vllm:x", "input": texts[i]}).encode(),
                                         headers={"Content-Type": "application/json"})
            with urllib.request.urlopen(req, timeout=60) as r:
                out[i] = json.loads(r.read())

        calls0 = emb.gpu_calls
        ts = [threading.Thread(target=worker, args=(i,)) for i in range(len(texts))]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        assert emb.gpu_calls - calls0 == srv.batches < len(texts)
        ids, lens = tokenizer.pack(texts, emb.vocab)
        w = synth.make_encoder(768, vocab=32768, seed=synth.BASE_SEED + 9)      # the fixture's weights
        want = oenc.encode(w.table_bf16, w.w_bf16, ids, lens)
        got = np.array([o["data"][0]["embedding"] for o in out], np.float64)
        assert np.max(np.abs(got - want)) < 2e-4
        assert np.allclose(np.linalg.norm(got, axis=1), 1.0, atol=1e-5)
    finally:
        srv.stop()


def test_semantic_similarity_pairs_on_gpu(emb):
    """Aggregator.calculate_semantic_similarity/3 mirror on the real GPU path vs the oracle's f64 cosine of the same
    (f32) embeddings: many pairs, one embedding call, one K3 launch."""
    from oracle import aggregator as oagg
    from quoracle_b200 import aggregator as pagg
    pairs = [(SENTENCES[0], SENTENCES[1]), (SENTENCES[2], SENTENCES[3]), (SENTENCES[4], SENTENCES[4]), ("", SENTENCES[0])]
    before = emb.ctx.launch_count()
    got = pagg.semantic_similarity_many(emb, pairs)
    assert emb.ctx.launch_count() - before <= 4                       # (pack +) K1 + K2 + one K3
    for (a, b), g in zip(pairs, got):
        if not a or not b:
            assert g == 0.0
            continue
        ea, eb = emb.get_embedding(a)[1]["embedding"], emb.get_embedding(b)[1]["embedding"]
        want = oagg.cosine_similarity(ea.astype(np.float64).tolist(), eb.astype(np.float64).tolist())
        assert abs(g - want) < 1e-6
    assert abs(got[2] - 1.0) < 1e-6


def test_lesson_store_on_gpu_matches_oracle(emb):
    """LessonStore: lesson rows resident in a device-side cache (qc_lesson_cache_*), one query row uploaded per dedup,
    against the oracle's accumulate_lessons fed the same embeddings; and against the uncached mirror."""
    fn1 = lambda t: oracle_fn(emb)(t, {})                                 # noqa: E731
    rng = np.random.default_rng(8)
    store = lesson_manager.LessonStore(emb, threshold=0.97, max_lessons=6)
    want, plain = [], []
    for step in range(10):
        news = [{"type": "factual", "content": " ".join(rng.choice(SENTENCES, int(rng.integers(1, 3)))) if rng.random() > 0.1 else "",
                 "confidence": int(rng.integers(1, 4))} for _ in range(int(rng.integers(1, 4)))]
        _, want = olm.accumulate_lessons(want, news, {"embedding_fn": fn1, "similarity_threshold": 0.97, "max_lessons": 6})
        _, plain = lesson_manager.accumulate_lessons(emb, plain, news, threshold=0.97, max_lessons=6)
        _, got = store.accumulate(news)
        assert got == want == plain, step
    store.close()


def test_lesson_cache_equals_first_match_batch(emb):
    ctx = emb.ctx
    rng = np.random.default_rng(4)
    rows = rng.standard_normal((100, ctx.dim)).astype(np.float32)
    cache = ctx.lesson_cache(128)
    cache.append(rows[:60])
    cache.append(rows[60:])
    assert len(cache) == 100
    for k in (0, 17, 99):
        q = (2.0 * rows[k] + 0.02 * rows[(k + 1) % 100]).astype(np.float32)
        idx, cos = cache.match(q, 0.95, want_cos=True)
        w_idx, w_cos = ctx.first_match_batch(q[None], rows[None], np.array([0.95], np.float32))
        assert idx == int(w_idx[0]) == k and np.array_equal(cos, w_cos[0])
    perm = rng.permutation(100)[:40]
    cache.gather(perm)
    idx, _ = cache.match(rows[perm[5]], 0.99)
    assert idx == 5 and len(cache) == 40
    cache.close()


def test_config3_tree_through_the_round_loop_matches_the_oracle(emb):
    """BASELINE config 3's shape -- a depth-4 / fan-out-8 agent tree (8 + 64 + 512 + 4096 agents) -- driven through the
    consensus round loop (rows A10 / A11) on the real GPU path: scripted pools (the :model_query_fn seam), majority /
    refinement / forced decision per agent, every sweep ONE batched flush, and the result of every agent equal to
    the sequential oracle's (execute_consensus_process/3 + execute_refinement/3 restated, consensus.ex:295-390).
    The oracle embeds through the GPU's own vectors (see oracle_fn): host logic + K3 are under test here."""
    from tests.test_host_logic import _compare_driver, _tree_agents
    agents = _tree_agents(4, 8)
    assert len(agents) == 8 + 64 + 512 + 4096
    before = emb.ctx.launch_count()
    got, kinds = _compare_driver(emb, agents, seed=3)
    launches = emb.ctx.launch_count() - before
    print(f"config-3 tree: {len(agents)} agents, outcomes {kinds}, {got['__sweeps__']} sweeps, {launches} kernel launches")
    assert kinds.get("consensus", 0) > 1000 and kinds.get("forced_decision", 0) > 100
    assert got["__sweeps__"] <= 6
    assert launches <= got["__sweeps__"] * 12                        # per sweep: pack + K1 + K2 + one K3 per pool size


def test_single_agent_latency_through_the_timed_collector(emb):
    """A11's latency bound: one agent alone, asking for one 8-value merge, is answered within the collector's
    max_wait_ms + the p95 of a single-decision call (profiles/latency_r1.md: 76 us on the C ABI; the Python mirror adds
    its own tokenising) + 2 ms of slack -- it never waits for a batch that does not fill."""
    import time
    from quoracle_b200.collector import TimedCollector
    col = TimedCollector(emb, max_pending=4096, max_wait_ms=2.0)
    try:
        vals = [f"{SENTENCES[i % len(SENTENCES)]} variant {i}" for i in range(8)]
        col.decide(vals, 0.9)                                           # warm (first call sizes workspaces)
        lat = []
        for k in range(30):
            vs = [v + f" #{k}" for v in vals]                           # new texts: no embedding-cache hits
            t0 = time.perf_counter()
            r, med = col.decide(vs, 0.9)
            lat.append(time.perf_counter() - t0)
            assert r[0] in ("ok", "error") and med is not None
        p95 = float(np.percentile(lat, 95)) * 1e3
        print(f"single-agent latency through TimedCollector(max_wait_ms=2): p50 {np.percentile(lat, 50) * 1e3:.2f} ms, p95 {p95:.2f} ms")
        assert p95 <= 2.0 + 0.076 + 2.0
    finally:
        col.close()
