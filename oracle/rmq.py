"""`taxa2agg -m rmq -a lca*`: the Euler tour, the range-minimum structure and the fold (oracle; test infrastructure only).

Literal restatement of /root/reference/src/taxon.rs:316-383 (EulerIterator), src/rmq/mod.rs:27-170 (RMQ: blocks of
usize::BITS = 64 positions, in-block labels, sparse table over the block minima, with the tie-breaking of every
comparison as written) and src/rmq/lca.rs:22-90 (LCACalculator: first occurrences, the fold of `aggregate`).

The fold runs over `taxons.keys()` of a HashMap, i.e. in an arbitrary order; `lca_star_rmq_orders` folds a record in
every given order.  tests/test_oracle_golden.py checks the reference's own vectors and, exhaustively over all orders
of small records on random trees, that the fold's answer does not depend on the order and equals the tree form's LCA*
(tree/lca.rs:34-40) -- which is why the product answers `-m rmq -a lca*` with the kernel of `-m tree -a lca*`.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence, Set, Tuple

from .taxonomy import Taxonomy, UnknownTaxon

BITS = 64  # size_of::<usize>() * 8 on the platforms the reference is built for


def euler_tour(tax: Taxonomy) -> List[Tuple[int, int]]:
    """taxon.rs:316-383: (taxon, depth) at every step of the tour; children in the order of the taxa file."""
    out: List[Tuple[int, int]] = []
    stack = [(tax.root, 0, 0)]  # node, depth, next child
    while stack:
        node, depth, k = stack.pop()
        out.append((node, depth))
        kids = tax.children.get(node, [])
        if k < len(kids):
            stack.append((node, depth, k + 1))
            stack.append((kids[k], depth + 1, 0))
    return out


def _clearbits(n: int, x: int) -> int:
    return (n >> x) << x


def _intlog2(n: int) -> int:
    return n.bit_length() - 1


class RMQ:
    """rmq/mod.rs:13-170."""

    def __init__(self, array: Sequence[int]):
        self.array = list(array)
        a = self.array
        # block_min (:53-66): Iterator::min_by_key returns the FIRST minimum
        self.block_min = []
        for i in range(0, len(a), BITS):
            c = a[i:i + BITS]
            self.block_min.append(i + min(range(len(c)), key=lambda j: (c[j], j)))
        # sparse (:77-86)
        def aggregate_minima(shift, minima):   # :68-74: `if array[l] < array[r] { l } else { r }`
            return [l if a[l] < a[r] else r for l, r in zip(minima, minima[shift:])]
        length = _intlog2(len(self.block_min)) if self.block_min else 0
        self.sparse = []
        if length >= 1 or True:
            self.sparse.append(aggregate_minima(1, self.block_min))
            for i in range(1, length):
                self.sparse.append(aggregate_minima(1 << i, self.sparse[i - 1]))
        # labels (:89-107)
        self.labels = []
        gstack: List[int] = []
        for i in range(len(a)):
            if i % BITS == 0:
                gstack = []
            self.labels.append(0)
            while gstack and a[i] < a[gstack[-1]]:
                gstack.pop()
            if gstack:
                g = gstack[-1]
                self.labels[i] = self.labels[g] | (1 << (g % BITS))
            gstack.append(i)

    def _min_in_block(self, left: int, right: int) -> int:   # :110-117
        v = _clearbits(self.labels[right], left % BITS)
        if v == 0:
            return right
        return _clearbits(left, _intlog2(BITS)) + ((v & -v).bit_length() - 1)

    def query(self, start: int, end: int) -> int:             # :121-169
        if start == end:
            return start
        left, right = (start, end) if start < end else (end, start)
        log2, size = _intlog2(BITS), BITS
        a = self.array
        block_diff = (right >> log2) - (left >> log2)
        if block_diff == 0:
            return self._min_in_block(left, right)
        l = self._min_in_block(left, _clearbits(left, log2) + size - 1)
        r = self._min_in_block(_clearbits(right, log2), right)
        if block_diff == 1:
            return l if a[l] <= a[r] else r
        if block_diff == 2:
            m = self.block_min[(left >> log2) + 1]
        else:
            k = _intlog2(block_diff - 1) - 1
            t1 = self.sparse[k][(left >> log2) + 1]
            t2 = self.sparse[k][(right >> log2) - (1 << (k + 1))]
            m = t1 if a[t1] <= a[t2] else t2
        ex = l if a[l] <= a[m] else m
        return ex if a[ex] <= a[r] else r


class LCACalculator:
    """rmq/lca.rs:10-56."""

    def __init__(self, tax: Taxonomy):
        self.euler_tour: List[int] = []
        depths: List[int] = []
        self.first_occurences: Dict[int, int] = {}
        for i, (tid, depth) in enumerate(euler_tour(tax)):
            self.euler_tour.append(tid)
            depths.append(depth)
            self.first_occurences.setdefault(tid, i)
        self.rmq = RMQ(depths)

    def first_occurence(self, tid: int) -> int:
        if tid not in self.first_occurences:
            raise UnknownTaxon(tid)
        return self.first_occurences[tid]

    def lca(self, left: int, right: int) -> int:              # :42-47
        return self.euler_tour[self.rmq.query(self.first_occurence(left), self.first_occurence(right))]

    def aggregate(self, keys: Sequence[int]) -> int:          # :60-90, `keys` = taxons.keys() in the order iterated
        if not keys:
            raise ValueError("Aggregration called on an empty list")
        arr = self.rmq.array
        consensus = self.first_occurence(keys[0])
        join_level: Optional[int] = None
        for t in keys[1:]:
            nxt = self.first_occurence(t)
            if consensus == nxt:
                continue
            rmq = self.rmq.query(consensus, nxt)
            a, b = rmq == consensus, rmq == nxt
            if not a and not b:
                lca, level = rmq, arr[rmq]
            elif a and not b:
                lca, level = nxt, join_level
            elif not a and b:
                lca, level = consensus, join_level
            else:
                raise AssertionError("Impossibru!")
            if join_level is not None and arr[lca] > join_level:
                lca = rmq   # join is below join level, we can't lower it
            consensus = lca
            join_level = level
        return self.euler_tour[consensus]


def lca_star_rmq_orders(calc: LCACalculator, keys: Iterable[int], orders: Iterable[Sequence[int]]) -> Set[int]:
    """The answers of the fold over the given orders of the record's distinct taxa."""
    keys = list(keys)
    return {calc.aggregate([keys[i] for i in o]) for o in orders}
