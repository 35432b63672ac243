"""Cross-agent batching: gather every pending decision of the agent tree into one GPU call.

The reference runs each agent's consensus synchronously inside its own GenServer
(lib/quoracle/agent/core.ex:421-422 -> message_handler.ex:125-148) and embeds the N values of each semantic
parameter one by one (consensus_rules.ex:302-318).  The collector is the new piece the north star asks
for: agents enqueue their parameter merges, ``flush`` evaluates all of them with one embedding call and one
similarity launch per pool size, and hands each agent its results in submission order.  (The production
form is the GenServer in elixir/quoracle_native.ex; this class is its executable specification.)
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Sequence

from . import consensus_rules
from .embeddings import Embeddings


@dataclass
class PendingDecision:
    agent_id: str
    param: str
    values: Sequence[Any]
    threshold: float


@dataclass
class DecisionCollector:
    emb: Embeddings
    max_pending: int = 4096
    _queue: list = field(default_factory=list)

    def submit(self, agent_id: str, param: str, values: Sequence[Any], threshold: float) -> int:
        """Enqueue one `{:semantic_similarity, threshold: t}` merge; returns its ticket."""
        self._queue.append(PendingDecision(agent_id, param, list(values), float(threshold)))
        return len(self._queue) - 1

    def should_flush(self) -> bool:
        return len(self._queue) >= self.max_pending

    def flush(self) -> dict:
        """-> {agent_id: {param: {"result": ("ok", v) | ("error", atom), "medoid": int | None}}}"""
        q, self._queue = self._queue, []
        results, medoids = consensus_rules.decide_many(self.emb, [(d.values, d.threshold) for d in q])
        out: dict = {}
        for d, r, m in zip(q, results, medoids):
            out.setdefault(d.agent_id, {})[d.param] = {"result": r, "medoid": m}
        return out


class TimedCollector:
    """The collector as agents see it at run time: `decide()` blocks the calling agent until its merge has been
    evaluated; a flush happens when `max_pending` merges are queued or `max_wait_ms` after the FIRST one of the batch
    was queued, whichever comes first ("Latency vs batching": a lone agent must not wait for a
    batch that never fills).  One worker thread owns the GPU calls, so the context is used from one thread at a time.
    Executable specification of the `Quoracle.Native.Collector` GenServer (elixir/quoracle_native.ex); the per-agent
    caller it serves is ConsensusHandler.get_action_consensus/1 (lib/quoracle/agent/consensus_handler.ex:64-246)."""

    def __init__(self, emb: Embeddings, max_pending: int = 4096, max_wait_ms: float = 2.0):
        import queue
        import threading
        self.emb = emb
        self.max_pending = int(max_pending)
        self.max_wait_ms = float(max_wait_ms)
        self.flushes = 0
        self.decided = 0
        self._q: "queue.Queue" = queue.Queue()
        self._empty = queue.Empty
        self._thread = threading.Thread(target=self._worker, daemon=True)
        self._event = threading.Event
        self._thread.start()

    def decide(self, values: Sequence[Any], threshold: float, timeout: float = 30.0):
        """Blocking: -> (("ok", value_0) | ("error", atom), medoid | None) for one `{:semantic_similarity, thr}` merge."""
        item = {"values": list(values), "thr": float(threshold), "done": self._event(), "out": None}
        self._q.put(item)
        if not item["done"].wait(timeout):
            raise TimeoutError("collector did not answer")
        return item["out"]

    def _worker(self):
        import time
        while True:
            first = self._q.get()
            if first is None:
                return
            batch = [first]
            deadline = time.monotonic() + self.max_wait_ms / 1000.0
            while len(batch) < self.max_pending:
                left = deadline - time.monotonic()
                if left <= 0:
                    break
                try:
                    nxt = self._q.get(timeout=left)
                except self._empty:
                    break
                if nxt is None:
                    self._q.put(None)
                    break
                batch.append(nxt)
            try:
                results, medoids = consensus_rules.decide_many(self.emb, [(d["values"], d["thr"]) for d in batch])
            except Exception:                  # library / device failure: every caller falls back to the mode, loudly
                results, medoids = [("error", "service_unavailable")] * len(batch), [None] * len(batch)
            self.flushes += 1
            self.decided += len(batch)
            for d, r, m in zip(batch, results, medoids):
                d["out"] = (r, m)
                d["done"].set()

    def close(self):
        self._q.put(None)
        self._thread.join(timeout=5)
