"""Iteration order of `java.util.HashMap<String,…>.keySet()` (TEST INFRASTRUCTURE).

`LPSolver.restoreInitialLP` sums the objective contributions while iterating
`initial.coefficients.keySet()` (LPSolver.java:213-233), so the rounding order of the rebuilt
`c` and `v` is HashMap iteration order.  The JDK's HashMap is deterministic: bucket index =
`(h ^ (h >>> 16)) & (cap - 1)` with `h = String.hashCode()`, buckets visited in index order,
entries inside a bucket in insertion order (resize splits preserve relative order).  The
table starts at 16 and doubles whenever `size > 0.75 * cap` — which is the history of the
maps that reach that loop (`new HashMap<>()` filled by `LPInputReader.processObjective`,
LPInputReader.java:45-46,142-143, or by `addDefaultVariables`, LPSolver.java:391-397).
Tree bins (>= 8 colliding keys in one bucket) are not modelled.
"""
from __future__ import annotations

from typing import List


def string_hash_code(s: str) -> int:
    """java.lang.String.hashCode: s[0]*31^(n-1) + … + s[n-1] in int32 arithmetic."""
    h = 0
    for ch in s:
        h = (31 * h + ord(ch)) & 0xFFFFFFFF
    return h


def _spread(h: int) -> int:
    return (h ^ (h >> 16)) & 0xFFFFFFFF


def table_capacity(size: int) -> int:
    cap = 16
    while size > 0.75 * cap:
        cap *= 2
    return cap


def hashmap_key_order(keys_in_insertion_order: List[str]) -> List[str]:
    cap = table_capacity(len(keys_in_insertion_order))
    buckets = {}
    for k in keys_in_insertion_order:
        buckets.setdefault(_spread(string_hash_code(k)) & (cap - 1), []).append(k)
    out: List[str] = []
    for idx in sorted(buckets):
        out.extend(buckets[idx])
    return out
