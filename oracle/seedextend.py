"""seedextend (unranked mode), oracle restatement (test infrastructure only).

Follows /root/reference/src/commands/seedextend.rs:94-149 statement by statement, :167-176 for the
flattening of the selected ranges and :151-164 for the ranked (`-r`) mode.
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


def rank_score(rank_index: int):
    """Rank::score, rank.rs:86-99, as written: a ladder of `self < X` tests on the enum order (PartialOrd is None,
    i.e. every test false, when either side is NoRank).  The first rung `self < Species` holds for every rank above
    species, so the other rungs never fire: ranks above species score 12, species and below and "no rank" None."""
    from .taxonomy import RANKS
    idx = {r: i for i, r in enumerate(RANKS)}

    def lt(other: str) -> bool:
        return rank_index != 0 and rank_index < idx[other]
    ladder = [("species", 12), ("species group", 11), ("genus", 10), ("tribe", 9), ("superfamily", 8), ("superorder", 7),
              ("superclass", 6), ("superphylum", 5), ("realm", 4), ("domain", 3), ("superkingdom", 2)]
    for name, sc in ladder:
        if lt(name):
            return sc
    return None


def taxon_score(tax, tid: int):
    """TaxonList::score, taxon.rs:181-191, on the list `new_with_unknown(taxa, true)` (seedextend.rs:84-90): taxon 0
    is the unknown taxon (no rank, its own parent) unless the file defines it."""
    current = tid
    steps = 0
    while True:
        if current == 0 and (len(tax.by_id) == 0 or tax.by_id[0] is None):
            t = (0, "unknown", 0, 0, False)
        elif current < len(tax.by_id) and tax.by_id[current] is not None:
            t = tax.by_id[current]
        else:
            return None
        if t[3] == current or t[2] != 0:
            return rank_score(t[2])
        current = t[3]
        steps += 1
        if steps > len(tax.by_id) + 1:
            raise ValueError("cycle in taxonomy")


def seedextend_ranked(ids: Sequence[int], tax, min_seed_size: int = 2, max_gap_size: int = 0, penalty: int = 5) -> List[int]:
    """seedextend.rs:151-164: only the extended seed with the highest summed score survives (max_by_key returns the
    LAST of equal maxima)."""
    t, seeds = seed_ranges(ids, min_seed_size, max_gap_size)
    if not seeds:
        return []
    best = None
    for a, b in seeds:
        sc = 0
        for x in t[a:b]:
            v = taxon_score(tax, x)
            sc += penalty if v is None else v
        if best is None or sc >= best[0]:
            best = (sc, a, b)
    return list(t[best[1]:best[2]])
