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

# Rule-level known answers, transcribed test by test.  Atoms are written as strings.
rule_vectors = {
    "apply_rule": [
        {"src": f"{RULES}:22-25", "rule": ["exact_match"], "values": ["command1", "command1", "command1"], "expect": ["ok", "command1"]},
        {"src": f"{RULES}:27-30", "rule": ["exact_match"], "values": ["command1", "command2"], "expect": ["error", "no_consensus"]},
        {"src": f"{RULES}:32-34", "rule": ["exact_match"], "values": [], "expect": ["error", "no_values"]},
        {"src": f"{RULES}:36-38", "rule": ["exact_match"], "values": ["value"], "expect": ["ok", "value"]},
        {"src": f"{RULES}:52-56", "rule": ["semantic_similarity", 0.95], "values": ["exact", "exact", "exact"], "expect": ["ok", "exact"]},
        {"src": f"{RULES}:58-62", "rule": ["semantic_similarity", 0.95], "values": ["", "", ""], "expect": ["ok", ""]},
        {"src": f"{RULES}:64-72", "rule": ["semantic_similarity", 0.99], "values": ["Process data", "Beautiful sunset", "Pizza recipe"],
         "expect": ["error", "no_consensus"], "note": "embeddings unavailable"},
        {"src": f"{RULES}:76-79", "rule": ["mode_selection"], "values": ["get", "get", "post", "get"], "expect": ["ok", "get"]},
        {"src": f"{RULES}:81-84", "rule": ["mode_selection"], "values": ["get", "post", "get", "post"], "expect": ["ok", "get"],
         "note": "tie: first in map (term) order"},
        {"src": f"{RULES}:86-89", "rule": ["mode_selection"], "values": ["atom1", "atom1", "atom2"], "expect": ["ok", "atom1"]},
        {"src": f"{RULES}:91-94", "rule": ["mode_selection"], "values": ["a", "b", "a", "a"], "expect": ["ok", "a"]},
        {"src": f"{RULES}:96-98", "rule": ["mode_selection"], "values": ["single"], "expect": ["ok", "single"]},
        {"src": f"{RULES}:114-117", "rule": ["union_merge"], "values": [["a", "b"], ["c"], ["d"]], "expect": ["ok", ["a", "b", "c", "d"]]},
        {"src": f"{RULES}:119-122", "rule": ["union_merge"], "values": [[], ["a"], []], "expect": ["ok", ["a"]]},
        {"src": f"{RULES}:108-112", "rule": ["union_merge"], "values": [["a", "a"], ["b", "b"], ["a", "b"]], "expect": ["ok", ["a", "b"]]},
        {"src": f"{RULES}:131-134", "rule": ["structural_merge"], "values": [{"a": 1}, {"b": 2}, {"c": 3}], "expect": ["ok", {"a": 1, "b": 2, "c": 3}]},
        {"src": f"{RULES}:136-139", "rule": ["structural_merge"], "values": [{"a": 1}, {"a": 2}, {"a": 3}], "expect": ["ok", {"a": 3}]},
        {"src": f"{RULES}:141-149", "rule": ["structural_merge"], "values": [{"outer": {"inner": 1}}, {"outer": {"other": 2}}],
         "expect": ["ok", {"outer": {"inner": 1, "other": 2}}]},
        {"src": f"{RULES}:151-154", "rule": ["structural_merge"], "values": [{}, {"a": 1}, {}], "expect": ["ok", {"a": 1}]},
        {"src": f"{RULES}:164-167", "rule": ["percentile", 50], "values": [100, 200, 300, 400, 500], "expect": ["ok", 300]},
        {"src": f"{RULES}:169-172", "rule": ["percentile", 75], "values": [100, 200, 300, 400], "expect": ["ok", 325]},
        {"src": f"{RULES}:174-176", "rule": ["percentile", 50], "values": [42], "expect": ["ok", 42]},
        {"src": f"{RULES}:178-181", "rule": ["percentile", 50], "values": [10, 20, 30, 40], "expect": ["ok", 25]},
        {"src": f"{RULES}:183-187", "rule": ["percentile", 50], "values": [10, 20, 30, 40, 50], "expect": ["ok", 30]},
        {"src": f"{RULES}:304-307", "rule": ["wait_parameter"], "values": [False, False, False], "expect": ["ok", False]},
        {"src": f"{RULES}:309-312", "rule": ["wait_parameter"], "values": [True, True, True], "expect": ["ok", True]},
        {"src": f"{RULES}:314-321", "rule": ["wait_parameter"], "values": [True, False, True], "expect": ["ok", True]},
        {"src": f"{RULES}:314-321", "rule": ["wait_parameter"], "values": [False, True, True, True, True], "expect": ["ok", True]},
        {"src": f"{RULES}:314-321", "rule": ["wait_parameter"], "values": [False, False, True], "expect": ["ok", True]},
        {"src": f"{RULES}:323-328", "rule": ["wait_parameter"], "values": [10, 20, 30], "expect": ["ok", 20]},
        {"src": f"{RULES}:323-328", "rule": ["wait_parameter"], "values": [5, 10, 15, 20, 25], "expect": ["ok", 15]},
        {"src": f"{RULES}:330-332", "rule": ["wait_parameter"], "values": [42], "expect": ["ok", 42]},
        {"src": f"{RULES}:334-337", "rule": ["wait_parameter"], "values": [False, 10, 20], "expect": ["ok", 10]},
        {"src": f"{RULES}:339-342", "rule": ["wait_parameter"], "values": [True, 10, 20], "expect": ["ok", 20]},
        {"src": f"{RULES}:344-347", "rule": ["wait_parameter"], "values": [True, False], "expect": ["ok", 15]},
        {"src": f"{RULES}:349-352", "rule": ["wait_parameter"], "values": [10, 20, 30, 40], "expect": ["ok", 25]},
        {"src": f"{RULES}:354-357", "rule": ["wait_parameter"], "values": [0, False, 0], "expect": ["ok", 0]},
    ],
    "merge_params": [
        {"src": f"{RULES}:191-196", "action": "spawn_child", "param": "task_description",
         "values": ["Process data", "Process data", "Process data"], "expect": ["ok", "Process data"]},
        {"src": f"{RULES}:198-201", "action": "wait", "param": "wait", "values": [1000, 2000, 3000, 4000, 5000], "expect": ["ok", 3000]},
        {"src": f"{RULES}:203-206", "action": "send_message", "param": "to", "values": ["parent", "parent", "parent"], "expect": ["ok", "parent"]},
        {"src": f"{RULES}:208-213", "action": "spawn_child", "param": "unknown", "values": ["value1", "value2"], "expect": ["error", "unknown_param"]},
        {"src": f"{RULES}:215-220", "action": "spawn_child", "param": "models", "values": [["model1"], ["model2"]], "expect": ["error", "unknown_param"]},
        {"src": f"{RULES}:222-225", "action": "execute_shell", "param": "command", "values": ["ls -la", "ls -la", "ls -la"], "expect": ["ok", "ls -la"]},
        {"src": f"{RULES}:227-233", "action": "execute_shell", "param": "command", "values": ["ls -la", "rm -rf"], "expect": ["error", "no_consensus"]},
        {"src": f"{RULES}:361-364", "action": "orient", "param": "wait", "values": [True, False, True], "expect": ["ok", True]},
    ],
    "wait_score": [
        {"src": f"{RESULT}:6-8", "value": True, "expect": [0, 0]},
        {"src": f"{RESULT}:10-12", "value": None, "expect": [0, 1]},
        {"src": f"{RESULT}:14-19", "value": 1, "expect": [0, 2]},
        {"src": f"{RESULT}:14-19", "value": 5, "expect": [0, 6]},
        {"src": f"{RESULT}:14-19", "value": 100, "expect": [0, 101]},
        {"src": f"{RESULT}:14-19", "value": 10000, "expect": [0, 10001]},
        {"src": f"{RESULT}:21-24", "value": False, "expect": [1, 0]},
        {"src": f"{RESULT}:21-24", "value": 0, "expect": [1, 0]},
    ],
    "cluster_wait_score": [
        {"src": f"{RESULT}:28-38", "waits": [False, 5, True], "expect": [1, 6]},
        {"src": f"{RESULT}:40-49", "waits": [False, "__missing__"], "expect": [1, 1]},
        {"src": f"{RESULT}:51-60", "waits": [False, False], "expect": [2, 0]},
        {"src": f"{RESULT}:62-71", "waits": [True, True], "expect": [0, 0]},
    ],
    "break_tie": [
        {"src": f"{RESULT}:585-594", "actions": ["spawn_child", "orient"], "expect": "orient"},
        {"src": f"{RESULT}:596-606", "actions": ["execute_shell", "spawn_child", "send_message"], "expect": "send_message"},
        {"src": f"{RESULT}:608-615", "actions": ["orient", "orient"], "expect_index": 0, "note": "returns first when all scores tie"},
    ],
    "confidence": [
        {"src": f"{RESULT}:712-716", "count": 5, "total": 5, "round": 1, "gt": 0.9},
        {"src": f"{RESULT}:718-723", "count": 3, "total": 5, "round": 2, "max_rounds": 2, "gt_other": {"round": 3, "max_rounds": 2}},
        {"src": f"{RESULT}:725-730", "count": 3, "total": 5, "round": 4, "max_rounds": 2, "minus_other": {"round": 5, "max_rounds": 2}, "equals": 0.1, "tol": 1e-6},
        {"src": f"{RESULT}:732-737", "count": 3, "total": 5, "round": 1, "same_as": {"round": 4}},
        {"src": f"{RESULT}:739-744", "count": 3, "total": 5, "round": 1, "max_rounds": 0, "lt_other": {"round": 1, "max_rounds": 4}},
        {"src": f"{RESULT}:746-755", "count": 1, "total": 10, "round": 5, "ge": 0.1, "le": 1.0},
        {"src": f"{RESULT}:757-763", "count": 4, "total": 5, "round": 1, "gt_other": {"count": 3}},
    ],
    "format_result": [
        {"src": f"{RESULT}:766-779", "clusters": [["spawn_child", 2], ["wait", 2], ["orient", 1]], "total": 5, "round": 1,
         "expect_kind": "forced_decision", "expect_action": "wait"},
        {"src": f"{RESULT}:781-801", "clusters": [["spawn_child", 1], ["execute_shell", 1], ["call_api", 1]], "total": 3, "round": 1,
         "expect_kind": "forced_decision", "expect_action": "call_api"},
    ],
    "mode_fallback": [
        {"src": f"{RESULT}:895-919", "action": "orient", "param": "current_situation", "values": ["aa", "b", "b", "ccccc"],
         "embedding": "fail", "expect": "b", "note": "embeddings unavailable -> MODE, not median length"},
    ],
    "majority": [
        {"src": f"{AGG}:174-275", "counts": [3], "total": 3, "round": 1, "expect": "majority"},
        {"src": f"{AGG}:174-275", "counts": [2, 1], "total": 3, "round": 1, "expect": "no_majority", "note": "round 1 needs unanimity"},
        {"src": f"{AGG}:174-275", "counts": [2, 1], "total": 3, "round": 2, "expect": "majority"},
        {"src": f"{AGG}:174-275", "counts": [2, 2], "total": 4, "round": 2, "expect": "no_majority", "note": "exactly 50% is not a majority"},
    ],
}

# :batch_sequence_merge known answers (2-arity apply_rule: no embeddings needed, semantic values are identical).
BATCH = "test/quoracle/actions/consensus_rules_batch_test.exs"
_fr = lambda path, **kw: {"action": "file_read", "params": dict({"path": path}, **kw)}       # noqa: E731
_todo = lambda op: {"action": "todo", "params": {"operation": op}}                           # noqa: E731
_orient = lambda s: {"action": "orient", "params": {"current_situation": s}}                 # noqa: E731
_spawn = {"action": "spawn_child", "params": {"task_description": "Process data", "success_criteria": "Complete",
                                              "immediate_context": "Test", "approach_guidance": "Standard",
                                              "profile": "test"}}
batch_merge = [
    {"src": f"{BATCH}:15-31", "sequences": [[_fr("/a.txt"), _todo("list")], [_fr("/a.txt"), _todo("list")]],
     "expect": "ok", "length": 2},
    {"src": f"{BATCH}:34-49", "sequences": [[_fr("/a.txt"), _todo("list")], [_fr("/b.txt")]],
     "expect": ["error", "sequence_length_mismatch"]},
    {"src": f"{BATCH}:52-68", "sequences": [[_fr("/a.txt"), _todo("list")], [_todo("add"), _fr("/b.txt")]],
     "expect": ["error", "sequence_mismatch"]},
    {"src": f"{BATCH}:71-98", "sequences": [[_fr("/a.txt", limit=100), _orient("Testing")],
                                            [_fr("/a.txt", limit=200), _orient("Testing")]],
     "expect": "ok", "length": 2,
     "merged": [{"action": "file_read", "params": {"path": "/a.txt", "limit": 150}},
                {"action": "orient", "params": {"current_situation": "Testing"}}],
     "note": "path exact_match, limit percentile 50 of 100, 200 = 150, current_situation identical -> no embedding call"},
    {"src": f"{BATCH}:101-116", "sequences": [[_fr("/a.txt")], [_fr("/b.txt")]], "expect": ["error", "no_consensus"],
     "note": "exact_match on differing paths fails the whole merge"},
    {"src": f"{BATCH}:119-137", "sequences": [[_todo("list"), _orient("Same")], [_todo("list"), _orient("Same")]],
     "expect": "ok", "length": 2, "actions": ["todo", "orient"]},
    {"src": f"{BATCH}:140-142", "sequences": [], "expect": "ok", "length": 0},
    {"src": f"{BATCH}:145-156", "sequences": [[_fr("/a.txt"), _todo("list")]], "expect": "ok", "length": 2,
     "actions": ["file_read", "todo"]},
    {"src": f"{BATCH}:160-170", "sequences": [[_todo("list")], [_todo("list")], [_todo("list")]], "expect": "ok",
     "length": 1, "actions": ["todo"]},
    {"src": f"{BATCH}:172-203", "sequences": [[_spawn], [_spawn]], "expect": "ok", "length": 1, "actions": ["spawn_child"]},
]

# batch_sync clusters in Result (tie-break by the MAX priority of the sequence, structural errors propagate)
RBATCH = "test/quoracle/consensus/result_batch_sync_test.exs"


def _batch_cluster(actions, count=2, wait=False):
    rep = {"action": "batch_sync", "params": {"actions": actions}, "reasoning": ""}
    return {"representative": rep, "actions": [{"action": "batch_sync", "params": {"actions": actions}, "wait": wait}],
            "count": count}


batch_result = {
    "break_tie": [
        {"src": f"{RBATCH}:18-63",
         "clusters": [_batch_cluster([{"action": "spawn_child", "params": {"task_description": "test"}},
                                      {"action": "file_read", "params": {"path": "/tmp/a.txt"}}]),
                      {"representative": {"action": "todo", "params": {"items": []}, "reasoning": ""},
                       "actions": [{"action": "todo", "wait": False, "params": {"items": []}}], "count": 2}],
         "expect_action": "todo", "note": "a batch ranks by the max priority of its members, not by batch_sync's own"},
        {"src": f"{RBATCH}:65-127",
         "clusters": [_batch_cluster([{"action": "file_read", "params": {"path": "/tmp/a.txt"}},
                                      {"action": "todo", "params": {"items": []}}]),
                      _batch_cluster([{"action": "file_read", "params": {"path": "/tmp/b.txt"}},
                                      {"action": "spawn_child", "params": {"task_description": "test"}}])],
         "expect_first_path": "/tmp/a.txt", "note": "lower max priority wins the tie"},
    ],
    "merge_error": {
        "src": f"{RBATCH}:288-331",
        "cluster": {"actions": [{"action": "batch_sync", "params": {"actions": [{"action": "file_read", "params": {"path": "/tmp/a.txt"}}]},
                                 "reasoning": "One file", "wait": False},
                                {"action": "batch_sync", "params": {"actions": [{"action": "file_read", "params": {"path": "/tmp/a.txt"}},
                                                                                {"action": "file_read", "params": {"path": "/tmp/b.txt"}}]},
                                 "reasoning": "Two files", "wait": False}],
                    "representative": {"action": "batch_sync", "params": {"actions": [{"action": "file_read", "params": {"path": "/tmp/a.txt"}}]},
                                       "reasoning": "One file"}},
        "expect": ["error", "sequence_length_mismatch"]},
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
    "batch_merge": batch_merge,
    "batch_result": batch_result,
}

if __name__ == "__main__":
    path = Path(__file__).parent / "reference_vectors.json"
    path.write_text(json.dumps(out, indent=1) + "\n")
    print(f"wrote {path} ({path.stat().st_size} bytes)")
