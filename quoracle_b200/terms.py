"""Erlang term semantics the host mirrors need: strict equality keys (`=:=`, so 1 and 1.0 differ and maps compare by
content, not by insertion order), Erlang term order (what `Enum.sort/1`, `Enum.max_by/2` over a frequency map and the
iteration order of small maps follow), `Enum.uniq/1` and the mode of a list (result.ex:367-372).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Any, Sequence


def key(v):
    """Hashable key with Erlang's exact equality: 1 and 1.0 differ, true is not 1, maps compare by content."""
    if isinstance(v, bool):
        return ("a", "true" if v else "false")
    if v is None:
        return ("a", "nil")
    if isinstance(v, int):
        return ("n", float(v), 0)
    if isinstance(v, float):
        return ("n", v, 1)
    if isinstance(v, str):
        return ("s", v.encode())
    if isinstance(v, (list, tuple)):
        return ("l", tuple(key(x) for x in v))
    if isinstance(v, dict):
        return ("m", tuple(sorted((key(k), key(x)) for k, x in v.items())))
    raise TypeError(type(v))


def order(v):
    """Erlang term order, restated: number < atom < tuple < map < list < binary.  nil / true / false are atoms and
    compare by name; lists compare element by element (a proper prefix is smaller); maps compare by size, then by
    their keys in ascending term order, then by the values in key order."""
    if isinstance(v, bool):
        return (1, "true" if v else "false")
    if v is None:
        return (1, "nil")
    if isinstance(v, int):
        return (0, float(v), 0)
    if isinstance(v, float):
        return (0, v, 1)
    if isinstance(v, str):
        return (5, v.encode("utf-8"))
    if isinstance(v, tuple):
        return (2, len(v), tuple(order(x) for x in v))
    if isinstance(v, list):
        return (4, tuple(order(x) for x in v))
    if isinstance(v, dict):
        items = sorted(((order(k), order(x)) for k, x in v.items()), key=lambda kv: kv[0])
        return (3, len(items), tuple(k for k, _ in items), tuple(x for _, x in items))
    raise TypeError(type(v))


def mode_value(values: Sequence[Any]):
    """Most frequent value; ties -> first in term order (Enum.frequencies + Enum.max_by, result.ex:367-372)."""
    counts: "OrderedDict" = OrderedDict()
    for v in values:
        k = key(v)
        c = counts.setdefault(k, [0, v])
        c[0] += 1
    best = None
    for k in sorted(counts, key=lambda kk: order(counts[kk][1])):
        if best is None or counts[k][0] > counts[best][0]:
            best = k
    return counts[best][1]


def uniq(values):
    seen, out = set(), []
    for v in values:
        k = key(v)
        if k not in seen:
            seen.add(k)
            out.append(v)
    return out
