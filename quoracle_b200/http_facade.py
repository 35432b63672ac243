"""OpenAI-compatible `/embeddings` endpoint over the C ABI.

The unmodified reference already knows how to talk to a local embedding server: an embedding model id with a
`vllm:` / `ollama:` / `lmstudio:` / `llamacpp:` / `tgi:` prefix plus an `endpoint_url` credential makes
`Embeddings` POST `{"model": ..., "input": text}` to `<endpoint_url>/embeddings` with `api_key: "not-needed"`
(reference lib/quoracle/models/local_model_helper.ex:13-19,92-110) and read `body["data"][0]["embedding"]`
(lib/quoracle/models/embeddings.ex:340-344).  This module is that server.  The reference sends one text per request,
serially per agent, so the server coalesces the requests that are in flight from *different* agents into one GPU
call (`Embeddings.get_embeddings`: cache, in-batch dedup, chunking, one `qc_embed_batch`).

HTTP status codes are chosen so that the reference maps them back to the atoms of include/qconsensus.h
(embeddings.ex:300-331): 400 + "context_length_exceeded" in the body -> :context_length_exceeded, other 400 ->
:bad_request, >= 500 -> :service_unavailable.  There is no CPU fallback: without the library the server answers 503.
"""
from __future__ import annotations

import base64
import json
import queue
import threading
import time
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer
from typing import Optional

import numpy as np

from . import tokenizer

_HTTP_FOR_ATOM = {
    "invalid_input": 400, "bad_request": 400, "context_length_exceeded": 400, "not_configured": 503,
    "service_unavailable": 503, "invalid_response": 502, "all_chunks_failed": 502,
}


class _Pending:
    __slots__ = ("texts", "done", "results")

    def __init__(self, texts):
        self.texts = texts
        self.done = threading.Event()
        self.results = None


class EmbeddingServer:
    """`EmbeddingServer(embeddings).start()`; `.url` is the value for the credential's `endpoint_url`."""

    def __init__(self, embeddings, host: str = "127.0.0.1", port: int = 0, model: str = "quoracle-b200-encoder",
                 max_batch: int = 512, max_wait_ms: float = 2.0):
        self.embeddings = embeddings
        self.model = model
        self.max_batch = max_batch
        self.max_wait_ms = max_wait_ms
        self.batches = 0                     # GPU flushes
        self.requests = 0
        self._q: "queue.Queue[Optional[_Pending]]" = queue.Queue()
        outer = self

        class Handler(BaseHTTPRequestHandler):
            protocol_version = "HTTP/1.1"

            def log_message(self, *a):     # quiet
                pass

            def _send(self, status: int, obj) -> None:
                body = json.dumps(obj).encode("utf-8")
                self.send_response(status)
                self.send_header("Content-Type", "application/json")
                self.send_header("Content-Length", str(len(body)))
                self.end_headers()
                self.wfile.write(body)

            def do_GET(self):
                if self.path.rstrip("/") in ("/health", "/v1/health"):
                    self._send(200, {"status": "ok", "model": outer.model, "dim": outer.dim})
                elif self.path.rstrip("/") in ("/models", "/v1/models"):
                    self._send(200, {"object": "list", "data": [{"id": outer.model, "object": "model"}]})
                else:
                    self._send(404, {"error": {"message": "not found", "code": "not_found"}})

            def do_POST(self):
                if self.path.rstrip("/") not in ("/embeddings", "/v1/embeddings"):
                    return self._send(404, {"error": {"message": "not found", "code": "not_found"}})
                try:
                    n = int(self.headers.get("Content-Length", "0"))
                    req = json.loads(self.rfile.read(n) or b"{}")
                    inp = req["input"]
                    texts = [inp] if isinstance(inp, str) else list(inp)
                    if not texts or not all(isinstance(t, str) for t in texts):
                        raise ValueError("input must be a string or a list of strings")
                    fmt = req.get("encoding_format", "float")
                    if fmt not in ("float", "base64"):
                        raise ValueError("encoding_format must be float or base64")
                except (KeyError, ValueError, json.JSONDecodeError) as e:
                    return self._send(400, {"error": {"message": f"bad request: {e}", "code": "bad_request"}})
                status, obj = outer._serve(texts, fmt, req.get("model") or outer.model)
                self._send(status, obj)

        self._httpd = ThreadingHTTPServer((host, port), Handler)
        self._httpd.daemon_threads = True
        self._threads: list[threading.Thread] = []

    # -- lifecycle -----------------------------------------------------------------------------
    @property
    def url(self) -> str:
        host, port = self._httpd.server_address[:2]
        return f"http://{host}:{port}"

    @property
    def dim(self) -> int:
        return int(getattr(self.embeddings.ctx, "dim", 0) or 0)

    def start(self) -> "EmbeddingServer":
        self._threads = [threading.Thread(target=self._httpd.serve_forever, daemon=True),
                         threading.Thread(target=self._batcher, daemon=True)]
        for t in self._threads:
            t.start()
        return self

    def stop(self) -> None:
        self._httpd.shutdown()
        self._httpd.server_close()
        self._q.put(None)
        for t in self._threads:
            t.join(timeout=5)

    # -- request side --------------------------------------------------------------------------
    def _serve(self, texts, fmt: str, model: str):
        p = _Pending(texts)
        self._q.put(p)
        if not p.done.wait(timeout=120):
            return 503, {"error": {"message": "embedding batcher timed out", "code": "service_unavailable"}}
        data, tokens = [], 0
        for i, (text, r) in enumerate(zip(texts, p.results)):
            if r[0] != "ok":
                atom = r[1]
                return _HTTP_FOR_ATOM.get(atom, 503), {"error": {"message": f"embedding failed: {atom}", "code": atom,
                                                                  "index": i}}
            vec = np.asarray(r[1]["embedding"], dtype="<f4")
            data.append({"object": "embedding", "index": i,
                         "embedding": base64.b64encode(vec.tobytes()).decode("ascii") if fmt == "base64"
                         else [float(x) for x in vec]})
            tokens += tokenizer.count_tokens(text)
        return 200, {"object": "list", "data": data, "model": model,
                     "usage": {"prompt_tokens": tokens, "total_tokens": tokens}}

    # -- GPU side: one flush for everything that is in flight ------------------------------------
    def _batcher(self) -> None:
        while True:
            first = self._q.get()
            if first is None:
                return
            group, n = [first], len(first.texts)
            deadline = time.monotonic() + self.max_wait_ms / 1000.0
            while n < self.max_batch:
                left = deadline - time.monotonic()
                if left <= 0:
                    break
                try:
                    nxt = self._q.get(timeout=left)
                except queue.Empty:
                    break
                if nxt is None:
                    self._q.put(None)
                    break
                group.append(nxt)
                n += len(nxt.texts)
            flat = [t for p in group for t in p.texts]
            try:
                res = self.embeddings.get_embeddings(flat)
            except Exception:                       # library missing / device lost: fail loudly to every caller
                res = [("error", "service_unavailable")] * len(flat)
            self.batches += 1
            self.requests += len(group)
            pos = 0
            for p in group:
                p.results = res[pos:pos + len(p.texts)]
                pos += len(p.texts)
                p.done.set()


def main():  # pragma: no cover - manual entry point: python -m quoracle_b200.http_facade --dim 768 --port 8089
    import argparse

    from . import native, synth
    from .embeddings import Embeddings
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--port", type=int, default=8089)
    args = ap.parse_args()
    w = synth.make_encoder(args.dim, seed=synth.BASE_SEED + 2)
    ctx = native.Context(args.device, w.table_bf16, w.w_bf16)
    srv = EmbeddingServer(Embeddings(ctx), port=args.port).start()
    print(f"serving {srv.url}/embeddings (dim {args.dim}, synthetic weights)")
    try:
        while True:
            time.sleep(3600)
    except KeyboardInterrupt:
        srv.stop()


if __name__ == "__main__":  # pragma: no cover
    main()
