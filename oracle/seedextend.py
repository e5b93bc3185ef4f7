"""seedextend (unranked mode), oracle restatement (test infrastructure only).

Follows /root/reference/src/commands/seedextend.rs:94-149 statement by statement, and
:167-176 for the flattening of the selected ranges.  The ranked (`-r`) mode (:151-164) is
out of scope.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple


def seed_ranges(ids: Sequence[int], min_seed_size: int = 2, max_gap_size: int = 0
                ) -> Tuple[List[int], List[Tuple[int, int]]]:
    t = list(ids) + [0]                      # :99 sentinel
    seeds: List[Tuple[int, int]] = []
    start = 0                                # :102
    end = 1
    last_tid = t[start]
    same_tid = 1
    same_max = 1
    while end < len(t):                      # :107
        if last_tid == t[end]:               # :109-113 same tid as last
            same_tid += 1
            end += 1
            continue
        if last_tid == 0 and same_tid > max_gap_size:   # :116-127 gap became too big
            if same_max >= min_seed_size:
                seeds.append((start, end - same_tid))
            start = end
            last_tid = t[end]
            same_tid = 1
            same_max = 1
            end += 1
            continue
        if last_tid == 0 and (end - start) == same_tid:  # :130-134 don't start with a gap
            end += 1
            start = end
            continue
        if last_tid != 0:                    # :137-139 another taxon
            same_max = max(same_max, same_tid)
        last_tid = t[end]                    # :140-142
        same_tid = 1
        end += 1
    if same_max >= min_seed_size:            # :144-149
        if last_tid == 0:
            end -= same_tid
        seeds.append((start, end))
    return t, seeds


def seedextend(ids: Sequence[int], min_seed_size: int = 2, max_gap_size: int = 0) -> List[int]:
    t, seeds = seed_ranges(ids, min_seed_size, max_gap_size)
    out: List[int] = []
    for a, b in seeds:                       # :170-173
        out.extend(t[a:b])
    return out
