"""OpenAI-compatible /embeddings endpoint: the reference's own local-model route
(lib/quoracle/models/local_model_helper.ex:92-110) posts {"model", "input"} and reads body["data"][0]["embedding"]
(lib/quoracle/models/embeddings.ex:340-344); error statuses must map back to its atoms (embeddings.ex:300-331).
Host logic on CPU against the fake context; the same checks run against the real library in test_gpu_host.py."""
import base64
import json
import threading
import urllib.error
import urllib.request

import numpy as np
import pytest

from quoracle_b200 import synth
from quoracle_b200.embeddings import Embeddings
from quoracle_b200.http_facade import EmbeddingServer
from tests.fake_ctx import FakeContext


def post(url, obj, path="/embeddings"):
    req = urllib.request.Request(url + path, data=json.dumps(obj).encode(), headers={"Content-Type": "application/json",
                                                                                     "Authorization": "Bearer not-needed"})
    try:
        with urllib.request.urlopen(req, timeout=30) as r:
            return r.status, json.loads(r.read())
    except urllib.error.HTTPError as e:
        return e.code, json.loads(e.read())


def reference_extract(body):
    """extract_embedding/1 of the reference (embeddings.ex:340-344)."""
    data = body.get("data")
    if isinstance(data, list) and data and isinstance(data[0], dict) and "embedding" in data[0]:
        return ("ok", data[0]["embedding"])
    return ("error", "invalid_response")


def reference_status_atom(status, body):
    """Status -> atom mapping of the reference's HTTP path (embeddings.ex:300-331)."""
    if 200 <= status <= 299:
        return None
    if status == 401:
        return "authentication_failed"
    if status == 429:
        return "rate_limit_exceeded"
    if status == 400:
        s = json.dumps(body)
        return "context_length_exceeded" if ("context_length_exceeded" in s or "maximum context length" in s) else "bad_request"
    if status >= 500:
        return "service_unavailable"
    return "network_error"


@pytest.fixture()
def server():
    w = synth.make_encoder(64, vocab=512, seed=3)
    emb = Embeddings(FakeContext(w.table_bf16, w.w_bf16), vocab=512)
    srv = EmbeddingServer(emb, max_wait_ms=200.0).start()   # generous window: coalescing must not depend on machine load
    yield srv, emb
    srv.stop()


def test_single_text_has_the_shape_the_reference_parses(server):
    srv, emb = server
    status, body = post(srv.url, {"model": "vllm:quoracle-b200-encoder", "input": "deploy the staging cluster"})
    assert status == 200 and body["object"] == "list" and body["model"] == "vllm:quoracle-b200-encoder"
    kind, vec = reference_extract(body)
    assert kind == "ok" and len(vec) == 64 and all(isinstance(x, float) for x in vec)
    direct = emb.get_embeddings(["deploy the staging cluster"])[0][1]["embedding"]
    assert np.array_equal(np.array(vec, np.float32), direct)
    assert body["usage"]["prompt_tokens"] == body["usage"]["total_tokens"] > 0
    assert post(srv.url, {"input": "deploy the staging cluster"}, path="/v1/embeddings")[0] == 200


def test_list_input_and_base64(server):
    srv, emb = server
    texts = ["alpha beta", "gamma delta epsilon", "alpha beta"]
    status, body = post(srv.url, {"input": texts, "encoding_format": "base64"})
    assert status == 200 and [d["index"] for d in body["data"]] == [0, 1, 2]
    got = [np.frombuffer(base64.b64decode(d["embedding"]), "<f4") for d in body["data"]]
    want = [r[1]["embedding"] for r in emb.get_embeddings(texts)]
    for g, w_ in zip(got, want):
        assert np.array_equal(g, w_)
    assert np.array_equal(got[0], got[2])


def test_error_statuses_map_back_to_reference_atoms(server):
    srv, _ = server
    status, body = post(srv.url, {"input": ""})
    assert status == 400 and reference_status_atom(status, body) == "bad_request" and body["error"]["code"] == "invalid_input"
    status, body = post(srv.url, {"inputs": "typo"})
    assert status == 400 and reference_status_atom(status, body) == "bad_request"
    status, body = post(srv.url, {"input": [1, 2]})
    assert status == 400
    status, body = post(srv.url, {"input": "x"}, path="/nope")
    assert status == 404
    assert reference_extract({"data": []}) == ("error", "invalid_response")


def test_context_length_maps_to_the_reference_atom():
    from quoracle_b200.native import QcError

    class TooLong(FakeContext):
        def embed_ragged(self, tokens, offsets):
            raise QcError(-4, "too many tokens", "context_length_exceeded")

    w = synth.make_encoder(64, vocab=512, seed=3)
    srv = EmbeddingServer(Embeddings(TooLong(w.table_bf16, w.w_bf16), vocab=512)).start()
    try:
        status, body = post(srv.url, {"input": "some text"})
        assert status == 400 and reference_status_atom(status, body) == "context_length_exceeded"
    finally:
        srv.stop()


def test_device_failure_is_a_503_not_a_fallback():
    class Broken(FakeContext):
        def embed_ragged(self, tokens, offsets):
            raise RuntimeError("libqconsensus.so is missing")

    w = synth.make_encoder(64, vocab=512, seed=3)
    srv = EmbeddingServer(Embeddings(Broken(w.table_bf16, w.w_bf16), vocab=512)).start()
    try:
        status, body = post(srv.url, {"input": "some text"})
        assert status == 503 and reference_status_atom(status, body) == "service_unavailable"
    finally:
        srv.stop()


def test_concurrent_agents_are_coalesced_into_few_gpu_calls(server):
    srv, emb = server
    texts = [f"agent {i} proposes action number {i % 7} with some words" for i in range(24)]
    results = [None] * len(texts)

    def worker(i):
        results[i] = post(srv.url, {"input": texts[i]})

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(len(texts))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert all(r[0] == 200 for r in results)
    assert srv.requests == 24 and srv.batches < 24            # the reference would have made 24 serial calls
    assert emb.ctx.calls["embed"] == srv.batches
    direct = emb.get_embeddings(texts)                         # now cached: no further GPU call
    for r, d in zip(results, direct):
        assert np.array_equal(np.array(r[1]["data"][0]["embedding"], np.float32), d[1]["embedding"])
    assert emb.ctx.calls["embed"] == srv.batches


def test_health_and_models(server):
    srv, _ = server
    with urllib.request.urlopen(srv.url + "/health", timeout=10) as r:
        assert json.loads(r.read())["status"] == "ok"
    with urllib.request.urlopen(srv.url + "/v1/models", timeout=10) as r:
        assert json.loads(r.read())["data"][0]["id"] == srv.model
