"""Reporting stages behind taxa2agg, restated line by line (oracle; test infrastructure only):
`snaptaxon` (src/commands/snaptaxon.rs:66-108), `taxa2freq` (src/commands/taxa2freq.rs:86-169), `bestof`
(src/commands/bestof.rs:50-79), over TaxonTree::filter_ancestors (src/taxon.rs:251-286).  Text in, text out."""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

from . import fasta
from .taxonomy import RANKS, Taxonomy, _parse_usize


def filter_ancestors(tax: Taxonomy, passes: Callable[[int], bool]) -> List[Optional[int]]:
    """taxon.rs:266-286: by id, the nearest ancestor-or-self passing the filter; the walk starts at the root with
    Some(root) as the inherited answer; ids the walk never reaches stay None."""
    out: List[Optional[int]] = [None] * (tax.max_id + 1)
    stack = [(tax.root, tax.root)]
    while stack:
        cur, anc = stack.pop()
        mine = cur if passes(cur) else anc
        out[cur] = mine
        for c in tax.children.get(cur, []):
            stack.append((c, mine))
    return out


def snaptaxon_text(text: str, tax: Taxonomy, rank: Optional[str] = None, taxons: Sequence[int] = (), invalid: bool = False) -> str:
    if rank == "no rank":
        raise ValueError("Snap to an actual rank.")
    rank_idx = RANKS.index(rank) if rank is not None else None

    def passes(tid: int) -> bool:
        if tid in taxons:
            return True
        t = tax.by_id[tid]
        return t is not None and (invalid or t[4]) and rank_idx is not None and t[2] == rank_idx

    snapping = filter_ancestors(tax, passes)
    out = []
    for line in text.split("\n")[:-1] if text.endswith("\n") else text.split("\n"):
        if line.endswith("\r"):
            line = line[:-1]
        if line.startswith(">"):
            out.append(line)
        else:
            s = snapping[_parse_usize(line)]
            out.append(str(0 if s is None else s))
    return "".join(x + "\n" for x in out)


def taxa2freq_rows(inputs: Sequence[str], tax: Taxonomy, rank: str = "species", min_frequency: int = 1
                   ) -> Tuple[List[Tuple[int, str, List[int]]], Dict[int, List[int]]]:
    """Rows (taxon id, name, counts per input) the command prints, ordered by descending total; rows with equal totals
    come in HashMap order in the reference, so callers compare those as sets."""
    if rank == "no rank":
        raise ValueError("Snap to an actual rank.")
    rank_idx = RANKS.index(rank)
    snapping = filter_ancestors(tax, lambda tid: tax.by_id[tid] is not None and tax.by_id[tid][2] == rank_idx)
    counts: Dict[int, List[int]] = {}
    for i, text in enumerate(inputs):
        for line in text.split("\n"):
            if line.endswith("\r"):
                line = line[:-1]
            try:
                t = _parse_usize(line)
            except ValueError:
                continue            # `if let Ok(taxon) = line?.parse()` (taxa2freq.rs:160)
            s = snapping[t]
            counts.setdefault(0 if s is None else s, [0] * len(inputs))[i] += 1
    rows = []
    for tid, row in sorted(counts.items(), key=lambda kv: -sum(kv[1])):
        if tax.by_id[tid] is None:
            raise KeyError("LCA taxon id not in taxon list. Check compatibility with index.")
        if sum(row) > min_frequency:     # strictly greater, as written (taxa2freq.rs:141)
            rows.append((tid, tax.by_id[tid][1], row))
    return rows, counts


def bestof_text(text: str, frames: int = 6) -> str:
    """bestof.rs:55-78 as written: the record that completes a group of `frames` is read but never a candidate."""
    out, chunk = [], []
    for header, seq in fasta.read_records(text, unwrap=False):
        if len(chunk) < frames - 1:
            chunk.append((header, seq))
            continue

        def score(rec):
            n = 0
            for tid in rec[1]:
                try:
                    v = _parse_usize(tid)
                except ValueError:
                    v = 0
                n += v not in (0, 1)
            return n

        best = None
        for rec in chunk:                # Iterator::max_by_key: the last of equal maxima
            if best is None or score(rec) >= score(best):
                best = rec
        if best is None:
            raise ValueError("called `Option::unwrap()` on a `None` value")
        out.append(fasta.write_record(best[0], best[1], "\n", False))
        chunk = []
    return "".join(out)
