"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's index construction chain
`umgap splitkmers | sort | umgap joinkmers` (the input of `umgap buildindex`).

Follows src/commands/splitkmers.rs:44-66 and src/commands/joinkmers.rs:53-105 line by line.  The hybrid
aggregator breaks ties by HashSet order (tree/mix.rs:52-55), so `joinkmers` returns the SET of taxa the
reference may print for a k-mer.  Pinned by the reference's aggregator vectors through oracle.agg
(tests/test_oracle_golden.py); the joinkmers doc example (joinkmers.rs:37-45) needs the NCBI taxonomy and cannot
be replayed here.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Sequence, Set, Tuple

from . import agg
from .taxonomy import Taxonomy


def splitkmers(rows: Iterable[Tuple[int, str]], k: int = 9, prefix: str = "") -> List[Tuple[str, int]]:
    """splitkmers.rs:44-66: (taxon id, sequence) rows -> (k-mer, taxon id) rows, in input order."""
    out = []
    byte = prefix[:1]
    for tid, seq in rows:
        if len(seq) < k:                      # :52-54
            continue
        for i in range(len(seq) - k + 1):     # :55 windows(k)
            kmer = seq[i:i + k]
            if byte:
                if kmer[0] == byte:           # :56-59 (k-1)-mer suffixes of the k-mers starting with the prefix
                    out.append((kmer[1:], tid))
            else:
                out.append((kmer, tid))
    return out


def joinkmers(rows: Sequence[Tuple[str, int]], tax: Taxonomy) -> Dict[str, Set[int]]:
    """joinkmers.rs:53-105 on rows sorted by k-mer: k-mer -> admissible consensus taxa (ranked-snapped)."""
    ranksnapping = tax.snapping(True)         # :62
    validsnapping = tax.snapping(False)       # :63
    out: Dict[str, Set[int]] = {}

    def emit(kmer: str, tids: List[int]) -> None:          # :66-75
        counts = agg.count(tids)
        if not counts:                                      # MixCalculator::aggregate: EmptyInput -> nothing printed
            return
        res = agg.hybrid(tax, counts, 0.95)
        out[kmer] = {ranksnapping[a] for a in res}

    current = None
    tids: List[int] = []
    for kmer, tid in rows:                                  # :80-100
        if current is not None and current != kmer:
            emit(current, tids)
            tids = []
        current = kmer
        v = validsnapping[tid] if tid < len(validsnapping) else None   # out of range: index panic in the reference
        if v is not None:
            tids.append(v)
    if current is not None:
        emit(current, tids)
    return out


def build(rows: Iterable[Tuple[int, str]], tax: Taxonomy, k: int = 9) -> Dict[str, Set[int]]:
    """splitkmers | sort | joinkmers."""
    return joinkmers(sorted(splitkmers(rows, k)), tax)
