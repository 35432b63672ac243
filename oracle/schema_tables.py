"""Action schema DATA the consensus math depends on, transcribed from the reference.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Only what the hot path reads is kept: parameter
order (``required ++ optional`` is the merge/fingerprint iteration order, result.ex:318,
aggregator.ex:368), the per-parameter consensus rule, and the tie-break priorities.

Sources (reference checkout):
  lib/quoracle/actions/schema/agent_schemas.ex:7-247   spawn_child, wait, send_message, orient, todo,
                                                        dismiss_child, adjust_budget
  lib/quoracle/actions/schema/api_schemas.ex:10-403    answer_engine ... batch_async
  lib/quoracle/actions/schema/metadata.ex:55-78        @action_priorities

Rules are tuples: ("semantic_similarity", threshold) | ("percentile", n) | ("exact_match",) |
("mode_selection",) | ("union_merge",) | ("structural_merge",) | ("batch_sequence_merge",) |
("first_non_nil",) | ("merge_maps",).  The last two appear in schemas but have NO apply_rule clause
(consensus_rules.ex:162-163 -> {:error, :unknown_rule}), so merging falls back to the mode.
"""
from __future__ import annotations


def _sem(t):
    return ("semantic_similarity", t)


EXACT = ("exact_match",)
MODE = ("mode_selection",)
UNION = ("union_merge",)
STRUCT = ("structural_merge",)
BATCH = ("batch_sequence_merge",)
FIRST = ("first_non_nil",)
MERGE_MAPS = ("merge_maps",)


def _pct(n):
    return ("percentile", n)


SCHEMAS = {
    # agent_schemas.ex:7-91
    "spawn_child": {
        "required": ["task_description", "success_criteria", "immediate_context", "approach_guidance", "profile"],
        "optional": ["role", "cognitive_style", "output_style", "delegation_strategy", "sibling_context",
                     "downstream_constraints", "skills", "budget", "grove_vars"],
        "rules": {"task_description": _sem(0.95), "success_criteria": _sem(0.85), "immediate_context": _sem(0.85),
                  "approach_guidance": _sem(0.85), "profile": EXACT, "role": _sem(0.85), "cognitive_style": MODE,
                  "output_style": MODE, "delegation_strategy": EXACT, "sibling_context": STRUCT,
                  "downstream_constraints": _sem(0.90), "skills": UNION, "budget": EXACT, "grove_vars": EXACT},
    },
    # agent_schemas.ex:93-105
    "wait": {"required": [], "optional": ["wait"], "rules": {"wait": _pct(50)}},
    # agent_schemas.ex:107-122
    "send_message": {"required": ["to", "content"], "optional": [], "rules": {"to": EXACT, "content": _sem(0.85)}},
    # agent_schemas.ex:124-193
    "orient": {
        "required": ["current_situation", "goal_clarity", "available_resources", "key_challenges",
                     "delegation_consideration"],
        "optional": ["assumptions", "unknowns", "approach_options", "parallelization_opportunities", "risk_factors",
                     "success_criteria", "next_steps", "constraints_impact"],
        "rules": {k: _sem(0.8) for k in [
            "current_situation", "goal_clarity", "available_resources", "key_challenges", "assumptions", "unknowns",
            "approach_options", "parallelization_opportunities", "risk_factors", "success_criteria", "next_steps",
            "constraints_impact", "delegation_consideration"]},
    },
    # agent_schemas.ex:195-213
    "todo": {"required": ["items"], "optional": [], "rules": {"items": _sem(0.85)}},
    # agent_schemas.ex:215-229
    "dismiss_child": {"required": ["child_id"], "optional": ["reason"], "rules": {"child_id": EXACT, "reason": FIRST}},
    # agent_schemas.ex:231-245
    "adjust_budget": {"required": ["child_id", "new_budget"], "optional": [],
                      "rules": {"child_id": EXACT, "new_budget": EXACT}},
    # api_schemas.ex:10-22
    "answer_engine": {"required": ["prompt"], "optional": [], "rules": {"prompt": _sem(0.95)}},
    # api_schemas.ex:24-47
    "execute_shell": {"required": [], "optional": ["command", "check_id", "working_dir", "terminate"],
                      "rules": {"command": EXACT, "check_id": EXACT, "working_dir": EXACT, "terminate": EXACT}},
    # api_schemas.ex:49-78
    "fetch_web": {"required": ["url"], "optional": ["security_check", "timeout", "user_agent", "follow_redirects"],
                  "rules": {"url": EXACT, "security_check": MODE, "timeout": _pct(50), "user_agent": EXACT,
                            "follow_redirects": MODE}},
    # api_schemas.ex:80-143
    "call_api": {
        "required": ["api_type", "url"],
        "optional": ["method", "query_params", "body", "headers", "auth", "query", "variables", "rpc_method",
                     "rpc_params", "rpc_id", "timeout", "max_body_size"],
        "rules": {"api_type": EXACT, "url": EXACT, "method": EXACT, "timeout": _pct(100), "auth": EXACT,
                  "query_params": EXACT, "body": EXACT, "headers": EXACT, "query": EXACT, "variables": EXACT,
                  "rpc_method": EXACT, "rpc_params": EXACT, "rpc_id": EXACT, "max_body_size": _pct(100)},
    },
    # api_schemas.ex:145-192
    "call_mcp": {
        "required": [],
        "optional": ["transport", "command", "url", "cwd", "connection_id", "tool", "arguments", "terminate", "timeout"],
        "rules": {"transport": EXACT, "command": EXACT, "url": EXACT, "cwd": EXACT, "connection_id": EXACT,
                  "tool": EXACT, "arguments": EXACT, "terminate": EXACT, "timeout": _pct(50)},
    },
    # api_schemas.ex:194-224
    "generate_secret": {"required": ["name"], "optional": ["length", "include_symbols", "include_numbers", "description"],
                        "rules": {"name": EXACT, "length": _pct(50), "include_symbols": MODE, "include_numbers": MODE,
                                  "description": _sem(0.8)}},
    # api_schemas.ex:226-238
    "search_secrets": {"required": ["search_terms"], "optional": [], "rules": {"search_terms": UNION}},
    # api_schemas.ex:240-256
    "generate_images": {"required": ["prompt"], "optional": ["source_image"],
                        "rules": {"prompt": _sem(0.95), "source_image": FIRST}},
    # api_schemas.ex:258-279
    "record_cost": {"required": ["amount"], "optional": ["description", "category", "metadata"],
                    "rules": {"amount": EXACT, "description": FIRST, "category": FIRST, "metadata": MERGE_MAPS}},
    # api_schemas.ex:281-298
    "file_read": {"required": ["path"], "optional": ["offset", "limit"],
                  "rules": {"path": EXACT, "offset": _pct(50), "limit": _pct(50)}},
    # api_schemas.ex:300-328
    "file_write": {"required": ["path", "mode"], "optional": ["content", "old_string", "new_string", "replace_all"],
                   "rules": {"path": EXACT, "mode": EXACT, "content": _sem(0.95), "old_string": EXACT,
                             "new_string": EXACT, "replace_all": MODE}},
    # api_schemas.ex:330-345
    "learn_skills": {"required": ["skills"], "optional": ["permanent"], "rules": {"skills": UNION, "permanent": MODE}},
    # api_schemas.ex:347-373
    "create_skill": {"required": ["name", "description", "content"], "optional": ["metadata", "attachments"],
                     "rules": {"name": EXACT, "description": _sem(0.90), "content": _sem(0.95), "metadata": MERGE_MAPS,
                               "attachments": UNION}},
    # api_schemas.ex:375-387
    "batch_sync": {"required": ["actions"], "optional": [], "rules": {"actions": BATCH}},
    # api_schemas.ex:389-401
    "batch_async": {"required": ["actions"], "optional": [], "rules": {"actions": BATCH}},
}

# metadata.ex:55-78 -- lower = more conservative = wins ties
ACTION_PRIORITIES = {
    "orient": 1, "send_message": 2, "batch_sync": 3, "batch_async": 4, "fetch_web": 5, "file_read": 6,
    "search_secrets": 7, "learn_skills": 8, "answer_engine": 9, "todo": 10, "adjust_budget": 11, "wait": 12,
    "generate_secret": 13, "generate_images": 14, "record_cost": 15, "call_mcp": 16, "call_api": 17,
    "execute_shell": 18, "file_write": 19, "dismiss_child": 20, "create_skill": 21, "spawn_child": 22,
}

DEFAULT_SEMANTIC_THRESHOLD = 0.9   # consensus_rules.ex:19


def get_schema(action):
    """Schema.get_schema/1 (lib/quoracle/actions/schema.ex:30): dict or None for an unknown action."""
    return SCHEMAS.get(action)


def get_action_priority(action):
    """Schema.get_action_priority/1 (schema.ex:76-84): int, or None for {:error, :unknown_action}."""
    return ACTION_PRIORITIES.get(action)


def semantic_params():
    """[(action, param, threshold)] for every {:semantic_similarity, threshold: t} rule."""
    out = []
    for action, sc in SCHEMAS.items():
        for p in sc["required"] + sc["optional"]:
            r = sc["rules"].get(p)
            if r and r[0] == "semantic_similarity":
                out.append((action, p, r[1]))
    return out
