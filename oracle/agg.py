"""Per-record taxon aggregation: LCA*, hybrid and MRTL (oracle; test infrastructure only).

Literal restatement of /root/reference/src/agg/mod.rs:27-44 (count, filter),
src/tree/mod.rs:29-101 (Tree::new/create/collapse/aggregate), src/tree/lca.rs:34-40 (LCA*),
src/tree/mix.rs:43-64 (hybrid), src/rmq/rtl.rs:28-57 (MRTL) and the record loop of
src/commands/taxa2agg.rs:159-181.

Where the reference picks among equal maxima by HashMap/HashSet iteration order
(tree/mix.rs:52-55, rmq/rtl.rs:52-55; acknowledged by its own tests at tree/mix.rs:79 and
rmq/rtl.rs:89-91) the functions here return the SET of every answer the reference can give.
f32 arithmetic is done with numpy.float32 exactly where the reference uses f32.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence, Set, Tuple

import numpy as np

from .taxonomy import Taxonomy, UnknownTaxon

f32 = np.float32


class EmptyInput(Exception):
    pass


def count(ids: Iterable[int]) -> Dict[int, np.float32]:
    """agg/mod.rs:27-36 with the unscored parser (taxa2agg.rs:150-152): 1.0 per occurrence."""
    c: Dict[int, np.float32] = {}
    for t in ids:
        c[t] = f32(c.get(t, f32(0.0)) + f32(1.0))
    return c


def count_scored(pairs: Iterable[Tuple[int, float]]) -> Dict[int, np.float32]:
    """agg/mod.rs:27-36 with the scored parser (taxa2agg.rs:141-148): a taxon's f32 scores added in input order."""
    c: Dict[int, np.float32] = {}
    for t, sc in pairs:
        c[t] = f32(c.get(t, f32(0.0)) + f32(sc))
    return c


def filter_counts(c: Dict[int, np.float32], lower_bound: float) -> Dict[int, np.float32]:
    """agg/mod.rs:39-44."""
    lb = f32(lower_bound)
    return {t: v for t, v in c.items() if v >= lb}


class _Tree:
    __slots__ = ("root", "value", "children")

    def __init__(self, root: int, value, children: List["_Tree"]):
        self.root = root
        self.value = value
        self.children = children


def _tree_new(tax: Taxonomy, taxons: Dict[int, np.float32], order=sorted) -> _Tree:
    """tree/mod.rs:29-67.  `order` arranges a node's children (a HashSet in the reference: any order can occur;
    it matters only to the rounding of f32 sums of scored input)."""
    tree: Dict[int, Set[int]] = {}
    queue = list(taxons.keys())
    qi = 0
    while qi < len(queue):
        tid = queue[qi]
        qi += 1
        parent = tax.parent(tid)             # :37 UnknownTaxon
        if tid == parent:
            continue
        if parent not in tree:
            queue.append(parent)
        tree.setdefault(parent, set()).add(tid)

    def create(root: int) -> _Tree:
        return _Tree(root, taxons.get(root, f32(0.0)),
                     [create(c) for c in order(tree.get(root, ()))])

    import sys
    sys.setrecursionlimit(max(10000, sys.getrecursionlimit()))
    return create(tax.root)


def _collapse(t: _Tree) -> _Tree:
    """tree/mod.rs:71-86 with combine = Add::add."""
    value = t.value
    new = t
    while len(new.children) == 1:
        new = new.children[0]
        value = f32(value + new.value)
    return _Tree(new.root, value, [_collapse(c) for c in new.children])


def _aggregate(t: _Tree) -> _Tree:
    """tree/mod.rs:90-101."""
    children = [_aggregate(c) for c in t.children]
    value = t.value
    for c in children:
        value = f32(value + c.value)
    return _Tree(t.root, value, children)


def lca_star(tax: Taxonomy, taxons: Dict[int, np.float32], order=sorted) -> int:
    """tree/lca.rs:34-40.  Deterministic."""
    if not taxons:
        raise EmptyInput()
    return _collapse(_tree_new(tax, taxons, order)).root


def hybrid(tax: Taxonomy, taxons: Dict[int, np.float32], factor: float, order=sorted) -> Set[int]:
    """tree/mix.rs:43-64.  Returns every answer reachable through tied maxima."""
    if not taxons:
        raise EmptyInput()
    fac = f32(factor)
    subtree = _aggregate(_collapse(_tree_new(tax, taxons, order)))
    results: Set[int] = set()

    def descend(base: _Tree):
        while True:
            if not base.children:
                results.add(base.root)
                return
            m = max(c.value for c in base.children)
            if f32(m / base.value) < fac:     # :57 f32 division then compare
                results.add(base.root)
                return
            tied = [c for c in base.children if c.value == m]
            if len(tied) == 1:
                base = tied[0]
                continue
            for c in tied:
                descend(c)
            return

    descend(subtree)
    return results


def mrtl(tax: Taxonomy, taxons: Dict[int, np.float32]) -> Set[int]:
    """rmq/rtl.rs:39-57.  Returns the set of arg-maxima."""
    rtl: Dict[int, np.float32] = {}
    for taxon, cnt in taxons.items():
        c = cnt
        nxt = taxon
        while True:
            # ancestors[next]; ancestors[root] = None (:31-33)
            if nxt == tax.root:
                break
            if nxt >= len(tax.parents) or tax.parents[nxt] is None:
                break
            anc = tax.parents[nxt]
            c = f32(c + taxons.get(anc, f32(0.0)))
            nxt = anc
        if nxt != tax.root:
            raise UnknownTaxon(nxt)           # :47-49
        rtl[taxon] = c
    if not rtl:
        raise EmptyInput()
    m = max(rtl.values())
    return {t for t, v in rtl.items() if v == m}


def _lca(tax: Taxonomy, a: int, b: int) -> int:
    """The plain LCA of two taxa (what rmq/lca.rs:42-47 computes through the Euler tour: the shallowest node between
    their first occurrences is their lowest common ancestor whichever of equal minima the RMQ returns)."""
    pa, pb = tax.root_path(a), tax.root_path(b)
    j = 0
    while j < min(len(pa), len(pb)) and pa[j] == pb[j]:
        j += 1
    return pa[j - 1]


def rmq_mix(tax: Taxonomy, taxons: Dict[int, np.float32], factor: float) -> Set[int]:
    """rmq/mix.rs:56-93 (`-m rmq -a hybrid`).  Returns the set of arg-maxima (max_by_key over a HashMap)."""
    fac = f32(factor)
    weights: Dict[int, List[np.float32]] = {}
    queue = list(taxons.keys())
    qi = 0
    while qi < len(queue):
        left = queue[qi]
        qi += 1
        if left in weights:
            continue
        for right, cnt in taxons.items():
            lca = _lca(tax, left, right)      # :71 (UnknownTaxon from first_occurence)
            if lca == left or lca == right:
                w = weights.setdefault(left, [f32(0.0), f32(0.0)])
                if lca == left:
                    w[0] = f32(w[0] + cnt)    # weight.lca
                if lca == right:
                    w[1] = f32(w[1] + cnt)    # weight.rtl
            else:
                queue.append(lca)
    if not weights:
        raise EmptyInput()
    val = {t: f32(f32(w[0] * fac) + f32(w[1] * f32(f32(1.0) - fac))) for t, w in weights.items()}   # :49-51
    m = max(val.values())
    return {t for t, v in val.items() if v == m}


LCA_STAR, HYBRID, MRTL, RMQ_HYBRID = 0, 1, 2, 3


def taxa2agg_record(tax: Taxonomy, snapping: Sequence[Optional[int]], ids: Sequence[int],
                    strategy: int, factor: float = 0.25, lower_bound: float = 0.0) -> Set[int]:
    """taxa2agg.rs:159-181 for one record: the set of ids the reference may print."""
    counts = filter_counts(count(t for t in ids if t != 0), lower_bound)   # :169-170
    if not counts:
        return {1}                                                        # :174-175 literal "1"
    if strategy == LCA_STAR:
        res = {lca_star(tax, counts)}
    elif strategy == HYBRID:
        res = hybrid(tax, counts, factor)
    elif strategy == MRTL:
        res = mrtl(tax, counts)
    else:
        raise ValueError("unknown strategy")
    out = set()
    for a in res:
        s = snapping[a]
        if s is None:
            raise UnknownTaxon(a)   # `.unwrap()` panic in the reference (:178)
        out.add(s)
    return out


def taxa2agg_record_scored(tax: Taxonomy, snapping: Sequence[Optional[int]], pairs: Sequence[Tuple[int, float]],
                           strategy: int, factor: float = 0.25, lower_bound: float = 0.0, orders=(sorted,)) -> Set[int]:
    """taxa2agg.rs:159-181 with -s for one record of (taxon, score) pairs.  `orders`: the child orders under which the
    induced tree is summed (the reference iterates a HashSet; `sorted` = ascending taxon id is the order of the device
    kernel); the result is the union over them and over tied maxima."""
    counts = filter_counts(count_scored((t, sc) for t, sc in pairs if t != 0), lower_bound)
    if not counts:
        return {1}
    res: Set[int] = set()
    if strategy == LCA_STAR:
        res = {lca_star(tax, counts)}
    elif strategy == HYBRID:
        for o in orders:
            res |= hybrid(tax, counts, factor, o)
    elif strategy == MRTL:
        res = mrtl(tax, counts)
    elif strategy == RMQ_HYBRID:
        res = rmq_mix(tax, counts, factor)
    else:
        raise ValueError("unknown strategy")
    out = set()
    for a in res:
        sn = snapping[a]
        if sn is None:
            raise UnknownTaxon(a)
        out.add(sn)
    return out
