"""Region sharding across GPUs: independent regions, no data-path collective.

Longest-processing-time greedy: regions sorted by candidate count, each assigned to the
currently lightest rank.  Deterministic, so every rank computes the same assignment."""
from __future__ import annotations

from typing import List, Sequence


def lpt_assign(costs: Sequence[int], n_ranks: int) -> List[List[int]]:
    """Returns, for each rank, the (ascending) list of region indices it owns."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0] * n_ranks
    owned: List[List[int]] = [[] for _ in range(n_ranks)]
    for i in order:
        r = min(range(n_ranks), key=lambda k: (load[k], k))
        owned[r].append(i)
        load[r] += costs[i]
    return [sorted(o) for o in owned]
