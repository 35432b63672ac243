"""BEAM term semantics the restatement needs: strict equality, term order, integer arithmetic.

TEST INFRASTRUCTURE (see oracle/__init__.py).

Representation: atoms and binaries are both Python ``str`` (the reference mixes atom and string
keys/values freely and normalises them, e.g. result.ex:321-328); ``nil`` is ``None``; ``true``/``false``
are ``bool``; integers/floats are ``int``/``float``; lists are ``list``; maps are ``dict``; tuples ``tuple``.
"""
from __future__ import annotations

import math
from typing import Any, Hashable, Iterable


def term_key(v: Any) -> Hashable:
    """Canonical hashable key implementing ``===`` (1 and 1.0 differ; True is not 1)."""
    if v is None:
        return ("nil",)
    if isinstance(v, bool):
        return ("bool", v)
    if isinstance(v, int):
        return ("int", v)
    if isinstance(v, float):
        return ("float", v)
    if isinstance(v, str):
        return ("str", v)
    if isinstance(v, (list,)):
        return ("list", tuple(term_key(x) for x in v))
    if isinstance(v, tuple):
        return ("tuple", tuple(term_key(x) for x in v))
    if isinstance(v, dict):
        return ("map", tuple(sorted(((term_key(k), term_key(x)) for k, x in v.items()), key=repr)))
    raise TypeError(f"unsupported term {type(v)!r}")


def term_equal(a: Any, b: Any) -> bool:
    return term_key(a) == term_key(b)


def uniq(values: Iterable[Any]) -> list:
    """Enum.uniq/1: first occurrence wins, strict equality."""
    seen, out = set(), []
    for v in values:
        k = term_key(v)
        if k not in seen:
            seen.add(k)
            out.append(v)
    return out


_TYPE_RANK = {"int": 0, "float": 0, "nil": 1, "bool": 1, "tuple": 2, "map": 3, "list": 4, "str": 5}


def term_order(v: Any):
    """Sort key approximating Erlang term order: number < atom < tuple < map < list < bitstring.

    nil/true/false are atoms (ordered by name: false < nil < true).  Python ``str`` ranks as a binary; a
    pool whose values mix real atoms and binaries is outside what this restatement can order."""
    k = term_key(v)
    tag = k[0]
    rank = _TYPE_RANK[tag]
    if tag in ("int", "float"):
        return (rank, float(v), 0 if tag == "int" else 1)
    if tag == "nil":
        return (rank, "nil")
    if tag == "bool":
        return (rank, "true" if v else "false")
    if tag == "str":
        return (rank, v.encode("utf-8"))
    if tag in ("list", "tuple"):
        seq = [term_order(x) for x in v]
        return (rank, (len(seq),) if tag == "tuple" else (), tuple(seq))
    if tag == "map":
        # Erlang: maps compare by size, then by their keys in ascending term order, then by the values in key order
        items = sorted(((term_order(kk), term_order(vv)) for kk, vv in v.items()), key=lambda kv: kv[0])
        return (rank, len(items), tuple(k for k, _ in items), tuple(x for _, x in items))
    raise TypeError(tag)


MAP_ORDER_LIMIT = 32   # above this many keys BEAM maps iterate in hash order: tie order is unpinned


def frequencies_max(values: list) -> Any:
    """Enum.frequencies/1 |> Enum.max_by(count) |> elem(0)  (result.ex:367-372, consensus_rules.ex:72-76).

    Map iteration is sorted term order for <= 32 distinct keys; max_by keeps the FIRST maximum, i.e. the
    smallest value in term order among the most frequent (pinned: [:get,:post,:get,:post] -> :get)."""
    counts: dict = {}
    first: dict = {}
    for v in values:
        k = term_key(v)
        counts[k] = counts.get(k, 0) + 1
        first.setdefault(k, v)
    keys = sorted(counts, key=lambda k: term_order(first[k]))
    best = None
    for k in keys:
        if best is None or counts[k] > counts[best]:
            best = k
    return first[best]


def erl_round(x: float) -> int:
    """Kernel.round/1: half away from zero."""
    return int(math.floor(x + 0.5)) if x >= 0 else -int(math.floor(-x + 0.5))


def erl_div(a: int, b: int) -> int:
    """Kernel.div/2: truncates toward zero."""
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def is_number(v: Any) -> bool:
    return isinstance(v, (int, float)) and not isinstance(v, bool)
