"""Taxonomy model of the reference, oracle restatement (test infrastructure only).

Follows /root/reference/src/taxon.rs:89-128 (TSV parse), :135-163 (TaxonList, ancestry),
:224-247 (TaxonTree::new, root detection), :251-301 (snapping) and src/rank.rs:9-44 (ranks).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

RANKS = [
    "no rank", "superkingdom", "domain", "realm", "kingdom", "subkingdom", "superphylum",
    "phylum", "subphylum", "superclass", "class", "subclass", "infraclass", "superorder",
    "order", "suborder", "infraorder", "parvorder", "superfamily", "family", "subfamily",
    "tribe", "subtribe", "genus", "subgenus", "species group", "species subgroup", "species",
    "subspecies", "varietas", "forma", "strain",
]  # rank.rs:9-44 (32 entries, index 0 = NoRank)
_RANK_INDEX = {r: i for i, r in enumerate(RANKS)}


class TaxonError(Exception):
    pass


class UnknownTaxon(TaxonError):
    def __init__(self, tid: int):
        super().__init__(f"Unknown Taxon ID: {tid}")
        self.tid = tid


Taxon = Tuple[int, str, int, int, bool]  # id, name, rank index, parent, valid


def parse_taxon(line: str) -> Taxon:
    """taxon.rs:89-113."""
    split = line.rstrip().split("\t")  # trim_end() strips trailing whitespace
    if len(split) != 5:
        raise TaxonError("Taxon requires five fields")
    try:
        tid = _parse_usize(split[0])
    except ValueError as e:
        raise TaxonError(str(e))
    if split[2] not in _RANK_INDEX:
        raise TaxonError("Matching variant not found")
    try:
        parent = _parse_usize(split[3])
    except ValueError as e:
        raise TaxonError(str(e))
    if split[4] == "\x01":
        valid = True
    elif split[4] == "\x00":
        valid = False
    else:
        raise TaxonError("Couldn't parse the valid byte")
    return tid, split[1], _RANK_INDEX[split[2]], parent, valid


def _parse_usize(s: str) -> int:
    # Rust usize::from_str: optional '+', ASCII digits only, no whitespace.
    t = s[1:] if s.startswith("+") else s
    if t == "" or not all("0" <= c <= "9" for c in t):
        raise ValueError("invalid digit found in string")
    v = int(t)
    if v >= 1 << 64:
        raise ValueError("number too large to fit in target type")
    return v


def format_taxon(t: Taxon) -> str:
    return f"{t[0]}\t{t[1]}\t{RANKS[t[2]]}\t{t[3]}\t" + ("\x01" if t[4] else "\x00")


def read_taxa(text: str) -> List[Taxon]:
    """taxon.rs:119-128."""
    lines = text.split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    return [parse_taxon(l[:-1] if l.endswith("\r") else l) for l in lines]


class Taxonomy:
    """TaxonList + TaxonTree + snapping in one object."""

    def __init__(self, taxa: Sequence[Taxon]):
        if not taxa:
            raise TaxonError("empty taxonomy")
        self.max_id = max(t[0] for t in taxa)
        self.by_id: List[Optional[Taxon]] = [None] * (self.max_id + 1)  # taxon.rs:135-145
        for t in taxa:
            self.by_id[t[0]] = t
        # TaxonTree::new, taxon.rs:224-247
        self.children: Dict[int, List[int]] = {}
        roots = {t[0] for t in taxa}
        for t in taxa:
            if t[0] == t[3]:
                continue
            self.children.setdefault(t[3], []).append(t[0])
            roots.discard(t[0])
        if len(roots) > 1:
            raise TaxonError("More than one root!")
        if not roots:
            raise TaxonError("There's no root!")
        self.root = next(iter(roots))
        # TaxonList::ancestry, taxon.rs:158-163
        self.parents: List[Optional[int]] = [t[3] if t is not None else None for t in self.by_id]

    def parent(self, tid: int) -> int:
        """parents[id].ok_or(UnknownTaxon(id)) (tree/mod.rs:37); out of range panics in the
        reference, treated as unknown here."""
        if tid >= len(self.parents) or self.parents[tid] is None:
            raise UnknownTaxon(tid)
        return self.parents[tid]

    def snapping(self, ranked_only: bool) -> List[Optional[int]]:
        """taxon.rs:251-301 (iterative form of the recursion)."""
        def ok(i: int) -> bool:
            t = self.by_id[i] if i < len(self.by_id) else None
            return t is not None and t[4] and (not ranked_only or t[2] != 0)

        out: List[Optional[int]] = [None] * (self.max_id + 1)
        stack = [(self.root, self.root)]
        while stack:
            cur, anc = stack.pop()
            a = cur if ok(cur) else anc
            if cur < len(out):
                out[cur] = a
            for c in self.children.get(cur, []):
                stack.append((c, a))
        return out

    def root_path(self, tid: int) -> List[int]:
        """[root, ..., tid] by parent walking; UnknownTaxon on a broken chain."""
        path = [tid]
        seen = 0
        while True:
            p = self.parent(path[-1])
            if p == path[-1]:
                break
            path.append(p)
            seen += 1
            if seen > len(self.parents):
                raise TaxonError("cycle in taxonomy")
        path.reverse()
        return path
