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
