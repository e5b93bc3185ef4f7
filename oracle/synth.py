"""numpy mirror of the device-side synthetic workload generator (umgap_b200/csrc/synth.cu).

TEST / BENCHMARK INFRASTRUCTURE ONLY.  The synthetic proteome, index values and reads are defined
by a counter-based hash, so any slice of the workload the GPU generated can be re-derived on the
host: the parity tests compare the device output with this mirror bit for bit, and bench.py's
CPU legs build their host-resident `fst` image from it.  Not part of the reference.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np

U64 = np.uint64
AAS = "ACDEFGHIKLMNPQRSTVWY"
CUM = np.array([5407, 6305, 9877, 14300, 16830, 21463, 22951, 26832, 30638, 36969,
                38548, 41209, 44309, 46884, 50510, 54855, 58361, 62858, 63572, 65535], dtype=np.uint32)
TABLE1 = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"
K_TAXON, K_VALUE, K_MUT, K_CODON = U64(0x7461786F6E), U64(0x76616C7565), U64(0x6D7574), U64(0x636F646F6E)

# codons of each residue in ascending TCAG-order index (what synth.cu's c_codon table lists)
CODONS = [[i for i, a in enumerate(TABLE1) if a == aa] for aa in AAS]
NCODON = np.array([len(c) for c in CODONS], dtype=np.uint64)
CODON_TAB = np.zeros((20, 6), dtype=np.uint32)
for _i, _c in enumerate(CODONS):
    CODON_TAB[_i, :len(_c)] = _c


def sm64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + U64(0x9E3779B97F4A7C15)).astype(U64)
        x = ((x ^ (x >> U64(30))) * U64(0xBF58476D1CE4E5B9)).astype(U64)
        x = ((x ^ (x >> U64(27))) * U64(0x94D049BB133111EB)).astype(U64)
        return (x ^ (x >> U64(31))).astype(U64)


def rnd3(seed, a, b) -> np.ndarray:
    a = np.asarray(a, dtype=U64)
    b = np.asarray(b, dtype=U64)
    with np.errstate(over="ignore"):
        return sm64(sm64(U64(seed) ^ (a * U64(0xD6E8FEB86659FD93))) ^ b)


def residue_index(seed: int, j, p) -> np.ndarray:
    u = (rnd3(seed, j, p) & U64(0xFFFF)).astype(np.uint32)
    return (u[..., None] > CUM[:19]).sum(axis=-1).astype(np.uint32)


class Preorder:
    """The library's preorder numbering: DFS from the root, children in ascending id order."""

    def __init__(self, taxa: Sequence[Tuple[int, str, int, int, bool]]):
        parent_of = {t[0]: t[3] for t in taxa}
        children: Dict[int, List[int]] = {}
        for tid, par in parent_of.items():
            if tid != par:
                children.setdefault(par, []).append(tid)
        all_children = {c for cs in children.values() for c in cs}
        root = [t for t in parent_of if t not in all_children][0]
        self.id_of: List[int] = []
        self.depth: List[int] = []
        self.parent_dense: List[int] = []
        stack = [(root, -1, 0)]
        while stack:
            tid, pd, d = stack.pop()
            me = len(self.id_of)
            self.id_of.append(tid)
            self.depth.append(d)
            self.parent_dense.append(me if pd < 0 else pd)
            for c in sorted(children.get(tid, []), reverse=True):
                stack.append((c, me, d + 1))
        self.n = len(self.id_of)
        self.dense_of = {t: i for i, t in enumerate(self.id_of)}
        self.id_arr = np.array(self.id_of, dtype=np.uint64)
        self.depth_arr = np.array(self.depth, dtype=np.uint64)

    def ancestor_at(self, dense: int, d: int) -> int:
        x = dense
        while self.depth[x] > d:
            x = self.parent_dense[x]
        return x

    def lca(self, a: int, b: int) -> int:
        while a != b:
            if self.depth[a] >= self.depth[b]:
                a = self.parent_dense[a]
            else:
                b = self.parent_dense[b]
        return a


def windows(seed: int, n_proteins: int, protein_len: int, home_pct: int, ancestor_pct: int, pre: Preorder,
            j0: int = 0, j1: int | None = None):
    """All 9-mer windows of proteins [j0, j1): (keys uint8 [n,9] as ASCII, values uint64 taxon ids)."""
    j1 = n_proteins if j1 is None else j1
    wpp = protein_len - 8
    js = np.repeat(np.arange(j0, j1, dtype=U64), wpp)
    ps = np.tile(np.arange(wpp, dtype=U64), j1 - j0)
    res = residue_index(seed, np.repeat(np.arange(j0, j1, dtype=U64), protein_len),
                        np.tile(np.arange(protein_len, dtype=U64), j1 - j0)).reshape(j1 - j0, protein_len)
    aa = np.frombuffer(AAS.encode(), dtype=np.uint8)[res]
    keys = np.lib.stride_tricks.sliding_window_view(aa, 9, axis=1).reshape(-1, 9)
    home = (rnd3(seed ^ int(K_TAXON), js, 0) % U64(pre.n)).astype(np.int64)
    rv = rnd3(seed ^ int(K_VALUE), js, ps)
    u = (rv % U64(100)).astype(np.int64)
    hi = (rv >> U64(32))
    dense = home.copy()
    anc_sel = (u >= home_pct) & (u < home_pct + ancestor_pct)
    other = u >= home_pct + ancestor_pct
    if anc_sel.any():
        idx = np.nonzero(anc_sel)[0]
        d = (hi[idx] % (pre.depth_arr[home[idx]] + U64(1))).astype(np.int64)
        dense[idx] = [pre.ancestor_at(int(h), int(dd)) for h, dd in zip(home[idx], d)]
    dense[other] = (hi[other] % U64(pre.n)).astype(np.int64)
    return np.ascontiguousarray(keys), pre.id_arr[dense]


def build_index(seed: int, n_proteins: int, protein_len: int, home_pct: int, ancestor_pct: int, pre: Preorder):
    """Sorted unique keys ([n,9] uint8) and values (uint64); duplicates merged by LCA."""
    keys, vals = windows(seed, n_proteins, protein_len, home_pct, ancestor_pct, pre)
    packed = np.zeros(len(keys), dtype=U64)
    for c in range(9):
        packed = (packed << U64(8)) | keys[:, c].astype(U64) if c < 8 else packed
    # 9 bytes do not fit 64 bits: sort by (first 8 bytes, last byte)
    order = np.lexsort((keys[:, 8], packed))
    keys, vals, packed = keys[order], vals[order], packed[order]
    same = np.zeros(len(keys), dtype=bool)
    same[1:] = (packed[1:] == packed[:-1]) & (keys[1:, 8] == keys[:-1, 8])
    if same.any():
        vals = vals.copy()
        starts = np.nonzero(~same)[0]
        group = np.cumsum(~same) - 1
        for i in np.nonzero(same)[0]:
            s = starts[group[i]]
            vals[s] = pre.id_of[pre.lca(pre.dense_of[int(vals[s])], pre.dense_of[int(vals[i])])]
        keys, vals = keys[~same], vals[~same]
    return np.ascontiguousarray(keys), np.ascontiguousarray(vals)


def reads(seed: int, n_proteins: int, protein_len: int, read_seed: int, first_pair: int, npairs: int,
          read_len: int, hit_pct: int) -> np.ndarray:
    """uint8 [npairs*2, read_len] nucleotides, mate 2 reverse-complemented (synth_reads_kernel)."""
    ncod = read_len // 3
    nreads = npairs * 2
    read = np.arange(nreads, dtype=U64)
    pair = U64(first_pair) + read // U64(2)
    mate = read & U64(1)
    rp = rnd3(read_seed, pair, 0)
    hit = ((rp % U64(100)) < U64(hit_pct)) & (protein_len >= ncod)
    j = (rp >> U64(8)) % U64(n_proteins)
    rm = rnd3(read_seed, pair, U64(1) + mate)
    o = rm % U64(protein_len - ncod + 1)
    shift = (rm >> U64(32)) % U64(3)
    x = np.arange(read_len, dtype=U64)[None, :]
    rid = (pair * U64(2) + mate)[:, None]
    rx = rnd3(read_seed ^ int(K_MUT), rid, x)
    rand_base = ((rx >> U64(16)) & U64(3)).astype(np.int64)
    sh = shift[:, None]
    coding = hit[:, None] & (x >= sh) & (((x - sh) // U64(3)) < ((U64(read_len) - sh) // U64(3)))
    q = np.where(coding, (x - sh) // U64(3), U64(0))
    ph = np.where(coding, (x - sh) % U64(3), U64(0)).astype(np.int64)
    aa = residue_index(seed, np.broadcast_to(j[:, None], q.shape), o[:, None] + q).astype(np.int64)
    rc = rnd3(read_seed ^ int(K_CODON), np.broadcast_to(rid, q.shape), q)
    codon = CODON_TAB[aa, (rc % NCODON[aa]).astype(np.int64)].astype(np.int64)
    tcag = (codon >> (2 * (2 - ph))) & 3
    base_c = np.array([3, 1, 0, 2], dtype=np.int64)[tcag]
    sub = ((rx & U64(0xFFFF)) % U64(100)) == U64(0)
    base_c = np.where(sub, rand_base, base_c)
    base = np.where(coding, base_c, rand_base)
    is_n = ((rx >> U64(32)) % U64(1000)) == U64(0)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    fwd = np.where(is_n, np.uint8(ord("N")), acgt[base])
    rev = np.where(is_n, np.uint8(ord("N")), acgt[3 - base])[:, ::-1]
    return np.where((mate == U64(1))[:, None], rev, fwd).astype(np.uint8)
