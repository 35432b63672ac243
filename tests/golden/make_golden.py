"""Writes tests/golden/reference_vectors.json -- the reference's own known-answer tests for the hot path.

The reference is Elixir and cannot be executed in this image (no BEAM), so these vectors are
TRANSCRIBED from the reference's ExUnit files: each entry carries the file:line it was read from
(relative to the reference checkout, shelvick/quoracle @ 6c0f0eb7) and the assertion it made.  Long
constant vectors are generated here exactly as the test generates them (``List.duplicate``).
Re-run with ``python tests/golden/make_golden.py`` -- it needs nothing but the standard library.
"""
from __future__ import annotations

import json
import math
from pathlib import Path

AGG = "test/quoracle/consensus/aggregator_test.exs"
RULES = "test/quoracle/actions/consensus_rules_test.exs"
LESSON = "test/quoracle/agent/lesson_manager_test.exs"
RESULT = "test/quoracle/consensus/result_test.exs"
EMBED = "test/quoracle/models/model_query_embedding_test.exs"


def one_hot(n, i):
    v = [0.0] * n
    v[i] = 1.0
    return v


cosine = [
    # Aggregator.cosine_similarity/2
    {"src": f"{AGG}:884-892", "fn": "aggregator", "a": [0.1, 0.2, 0.3, 0.4, 0.5], "b": [0.11, 0.19, 0.31, 0.39, 0.51],
     "gt": 0.99},
    {"src": f"{AGG}:922-928", "fn": "aggregator", "a": [0.1, 0.2, 0.3], "b": [0.1, 0.2, 0.3], "equals": 1.0, "tol": 0.0,
     "f32_tol": 1e-6},
    {"src": f"{AGG}:924-931", "fn": "aggregator", "a": [0.1, 0.2, 0.3], "b": [-0.1, -0.2, -0.3], "equals": -1.0,
     "tol": 0.01},
    {"src": f"{AGG}:936-943", "fn": "aggregator", "a": [0.1, 0.2, 0.3], "b": [0.11, 0.21, 0.29], "gt": 0.95},
    {"src": f"{AGG}:936-946", "fn": "aggregator", "a": [0.1, 0.2, 0.3], "b": [-0.1, -0.2, -0.3], "lt": -0.9},
    # ConsensusRules.cosine_similarity/2, 3072-dim
    {"src": f"{RULES}:241-248", "fn": "rules", "a": [0.5] * 3072, "b": [0.5] * 3072, "equals": 1.0, "tol": 1e-4},
    {"src": f"{RULES}:250-258", "fn": "rules", "a": one_hot(3072, 0), "b": one_hot(3072, 1), "equals": 0.0, "tol": 1e-4},
    {"src": f"{RULES}:260-267", "fn": "rules", "a": [1.0] * 3072, "b": [-1.0] * 3072, "equals": -1.0, "tol": 1e-4},
    {"src": f"{RULES}:279-290", "fn": "rules", "a": [0.0] * 3072, "b": [0.0] * 3072, "equals": 0.0, "tol": 0.0},
    # LessonManager R12
    {"src": f"{LESSON}:405-416", "fn": "aggregator", "a": [1.0, 0.0, 0.0], "b": [0.95, 0.05, 0.0], "gt": 0.90,
     "lt": 1.0},
]
for c in cosine:
    c.setdefault("tol", 1e-6)

cosine_raises = [
    {"src": f"{RULES}:269-277", "fn": "rules", "len_a": 3072, "len_b": 100, "raises": True},
    {"src": "lib/quoracle/consensus/aggregator.ex:350", "fn": "aggregator", "len_a": 3, "len_b": 2, "equals": 0.0},
]

# Lesson dedup (first match >= threshold); vectors from the mock embedding fn, lesson_manager_test.exs:34-76.
lesson_dedup = [
    {"src": f"{LESSON}:34-41,418-430", "new": [0.95, 0.05, 0.0], "existing": [[1.0, 0.0, 0.0]], "threshold": 0.90,
     "expect_index": 0},
    {"src": f"{LESSON}:53-60", "new": [0.22, 0.97, 0.0], "existing": [[0.5, 0.5, 0.0]], "threshold": 0.90,
     "expect_index": -1, "note": "cos ~0.84 < 0.90: not merged"},
    {"src": f"{LESSON}:42-49", "new": [0.05, 0.95, 0.0], "existing": [[1.0, 0.0, 0.0], [0.0, 1.0, 0.0]],
     "threshold": 0.90, "expect_index": 1},
    {"src": f"{LESSON}:65-72", "new": one_hot(200, 7), "existing": [one_hot(200, i) for i in range(7)] +
     [one_hot(200, i) for i in range(8, 30)], "threshold": 0.90, "expect_index": -1,
     "note": "200-dim one-hots never merge"},
    {"src": f"{LESSON}:62-63", "new": [-1.0, 0.0, 0.0], "existing": [[1.0, 0.0, 0.0]], "threshold": 0.90,
     "expect_index": -1},
]

# Chunk mean, model_query_embedding_test.exs:47-148: constant 0.01 vectors -> mean is 0.01 everywhere.
chunk_mean = [
    {"src": f"{EMBED}:47-72", "chunks": 3, "dim": 3072, "value": 0.01, "expect": 0.01, "tol": 1e-12},
    {"src": f"{EMBED}:136-148", "chunks": 5, "dim": 3072, "value": 0.01, "expect": 0.01, "tol": 1e-12},
]

# sha256 deterministic test embedding, lib/quoracle/agent/consensus/per_model_query/helpers.ex:69-76
# (sha256(text) bytes / 255 -> 32-dim): reproducible with hashlib, the bridge for BASELINE config 1.
sha_embedding = {"src": "lib/quoracle/agent/consensus/per_model_query/helpers.ex:69-76", "dim": 32,
                 "texts": ["analyze the data", "analyse the data", "deploy to production"]}

# Constant embedding sin(1..10) for every value => cosine 1 => the rule returns value_0.
constant_embedding_rule = {"src": f"{RULES}:487-552; {RESULT}:839-884",
                           "embedding": [math.sin(i) for i in range(1, 11)],
                           "values": ["first phrasing", "second phrasing", "third phrasing"],
                           "threshold": 0.8, "expect": "first phrasing"}

# Rule-level known answers (filled in by the rule oracle's tests): see RULE_VECTORS below.
rule_vectors = {
    "exact_match": [
        {"src": f"{RULES}:16-49", "values": ["a", "a", "a"], "expect": ["ok", "a"]},
        {"src": f"{RULES}:16-49", "values": ["a", "b", "a"], "expect": ["error", "no_consensus"]},
        {"src": "lib/quoracle/actions/consensus_rules.ex:49", "values": [], "expect": ["error", "no_values"]},
        {"src": "lib/quoracle/actions/consensus_rules.ex:50", "values": [42], "expect": ["ok", 42]},
    ],
    "semantic_identical": [
        {"src": f"{RULES}:52-72", "values": ["same", "same", "same"], "threshold": 0.9, "expect": ["ok", "same"],
         "embedding_calls": 0},
        {"src": f"{RULES}:428-458", "values": ["", "", ""], "threshold": 0.9, "expect": ["ok", ""], "embedding_calls": 0},
        {"src": f"{RULES}:555-575", "values": ["alpha text", "beta text"], "threshold": 0.9, "embedding": "fail",
         "expect": ["error", "no_consensus"]},
    ],
    "mode_selection": [
        {"src": f"{RULES}:76-99", "values": ["get", "post", "get", "post"], "expect": ["ok", "get"],
         "note": "atoms :get/:post; tie -> first in term order"},
        {"src": f"{RULES}:76-99", "values": ["a", "b", "a", "a"], "expect": ["ok", "a"]},
        {"src": f"{RESULT}:895-919", "values": ["aa", "b", "b", "ccccc"], "expect": ["ok", "b"],
         "note": "orient fallback is the MODE, not median length"},
    ],
    "percentile": [
        {"src": "lib/quoracle/actions/consensus_rules.ex:355", "p": 75, "values": [100, 200, 300, 400], "expect": 325},
        {"src": f"{RULES}:164-188", "p": 50, "values": [10, 20, 30], "expect": 20},
        {"src": f"{RULES}:164-188", "p": 50, "values": [10, 20, 30, 40], "expect": 25},
        {"src": f"{RULES}:164-188", "p": 100, "values": [5, 50, 20], "expect": 50},
    ],
    "wait_parameter": [
        {"src": f"{RULES}:303-367", "values": [False, False, False], "expect": False},
        {"src": f"{RULES}:303-367", "values": [True, True, True], "expect": True},
        {"src": f"{RULES}:303-367", "values": [True, False, False], "expect": True,
         "note": ">= 3 booleans with any true -> true"},
        {"src": f"{RULES}:303-367", "values": [10, 20, 30], "expect": 20},
        {"src": f"{RULES}:303-367", "values": [10, 20], "expect": 15},
        {"src": f"{RULES}:303-367", "values": [True, False], "expect": 15, "note": "true -> 30, false -> 0, median 15"},
        {"src": f"{RULES}:303-367", "values": [False, 10, 20], "expect": 10},
        {"src": f"{RULES}:303-367", "values": [True, 10, 20], "expect": 20, "note": "true -> max int"},
    ],
    "wait_score": [
        {"src": f"{RESULT}:12-36", "value": True, "expect": [0, 0]},
        {"src": f"{RESULT}:12-36", "value": None, "expect": [0, 1]},
        {"src": f"{RESULT}:12-36", "value": 5, "expect": [0, 6]},
        {"src": f"{RESULT}:12-36", "value": 0, "expect": [1, 0]},
        {"src": f"{RESULT}:12-36", "value": False, "expect": [1, 0]},
    ],
    "confidence": [
        {"src": f"{RESULT}:711-778", "count": 3, "total": 3, "round": 1, "gt": 0.9},
        {"src": f"{RESULT}:711-778", "count": 3, "total": 3, "round": 1, "equals": 1.0},
        {"src": f"{RESULT}:711-778", "count": 2, "total": 3, "round": 1, "equals": 2 / 3 + 0.10},
        {"src": f"{RESULT}:711-778", "count": 3, "total": 3, "round": 6, "equals": 0.8,
         "note": "two rounds beyond max 4: 1.0 + .15 - .2 -> clamp(0.95)?  computed: 1.15 - 0.2 = 0.95 -> 0.95",
         "skip": True},
        {"src": f"{RESULT}:711-778", "count": 1, "total": 10, "round": 20, "equals": 0.1, "note": "clamped at 0.1"},
    ],
    "majority": [
        {"src": f"{AGG}:174-275", "counts": [3], "total": 3, "round": 1, "expect": "majority"},
        {"src": f"{AGG}:174-275", "counts": [2, 1], "total": 3, "round": 1, "expect": "no_majority",
         "note": "round 1 needs unanimity"},
        {"src": f"{AGG}:174-275", "counts": [2, 1], "total": 3, "round": 2, "expect": "majority"},
        {"src": f"{AGG}:174-275", "counts": [2, 2], "total": 4, "round": 2, "expect": "no_majority",
         "note": "exactly 50% is not a majority"},
    ],
}

out = {
    "reference": "shelvick/quoracle @ 6c0f0eb7 (app 0.2.4)",
    "how": "transcribed from the reference's ExUnit tests by tests/golden/make_golden.py; the reference "
           "cannot run in this image (no BEAM)",
    "cosine": cosine,
    "cosine_raises": cosine_raises,
    "lesson_dedup": lesson_dedup,
    "chunk_mean": chunk_mean,
    "sha_embedding": sha_embedding,
    "constant_embedding_rule": constant_embedding_rule,
    "rules": rule_vectors,
}

if __name__ == "__main__":
    path = Path(__file__).parent / "reference_vectors.json"
    path.write_text(json.dumps(out, indent=1) + "\n")
    print(f"wrote {path} ({path.stat().st_size} bytes)")
