"""Seeded synthetic inputs for the parity tests (SURVEY 8(d) recipe, small sizes).

Plain numpy / Python; independent of both the oracle and the CUDA library.
"""
from __future__ import annotations

import random
from typing import Dict, List, Tuple

AAS = "ACDEFGHIKLMNPQRSTVWY"
TABLE1 = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"
TCAG = "TCAG"
CODONS: Dict[str, List[str]] = {}
for _i, _a in enumerate(TABLE1):
    CODONS.setdefault(_a, []).append(TCAG[_i // 16] + TCAG[(_i // 4) % 4] + TCAG[_i % 4])

RANK_NAMES = [
    "no rank", "superkingdom", "domain", "realm", "kingdom", "subkingdom", "superphylum",
    "phylum", "subphylum", "superclass", "class", "subclass", "infraclass", "superorder",
    "order", "suborder", "infraorder", "parvorder", "superfamily", "family", "subfamily",
    "tribe", "subtribe", "genus", "subgenus", "species group", "species subgroup", "species",
    "subspecies", "varietas", "forma", "strain",
]


def make_taxonomy(n: int, seed: int = 1, max_depth: int = 40) -> List[Tuple[int, str, int, int, bool]]:
    """Random recursive tree with sparse ids; root id 1 (parent 1).  Returns oracle Taxon tuples
    (id, name, rank index, parent, valid)."""
    rng = random.Random(seed)
    ids = [1] + sorted(rng.sample(range(2, 3 * n + 2), n - 1))
    rng.shuffle(ids[1:])
    depth = {1: 0}
    rank_of = {1: 0}
    taxa = [(1, "t1", 0, 1, True)]
    placed = [1]
    for tid in ids[1:]:
        while True:
            par = placed[int(rng.random() ** 0.6 * len(placed))] if rng.random() < 0.7 else rng.choice(placed)
            if depth[par] < max_depth:
                break
        depth[tid] = depth[par] + 1
        # ranks non-decreasing along a path, ~35 % "no rank"
        if rng.random() < 0.35:
            rk = 0
        else:
            lo = max([rank_of[par]] + [0])
            rk = min(31, max(lo, 1) + rng.randrange(0, 4))
        rank_of[tid] = rk if rk else rank_of[par]
        taxa.append((tid, f"t{tid}", rk, par, rng.random() >= 0.03))
        placed.append(tid)
    rng.shuffle(taxa)
    return taxa


def taxonomy_arrays(taxa):
    import numpy as np
    return (np.array([t[0] for t in taxa], dtype=np.uint64), np.array([t[3] for t in taxa], dtype=np.uint64),
            np.array([t[2] for t in taxa], dtype=np.uint8), np.array([1 if t[4] else 0 for t in taxa], dtype=np.uint8))


def make_proteome(n_proteins: int, seed: int = 2, lo: int = 60, hi: int = 200) -> List[str]:
    rng = random.Random(seed)
    out = []
    for _ in range(n_proteins):
        L = rng.randrange(lo, hi)
        s = [rng.choice(AAS) for _ in range(L)]
        for i in range(L):
            if rng.random() < 0.001:
                s[i] = rng.choice("BXZUO")
        out.append("".join(s))
    return out


def make_index(proteins: List[str], tax, seed: int = 2, k: int = 9) -> Dict[bytes, int]:
    """All k-mers of all proteins; value = home taxon (70 %), one of its ancestors (20 %) or an
    unrelated taxon (10 %); duplicates merge to the LCA (tax is an oracle Taxonomy)."""
    rng = random.Random(seed * 7919 + 1)
    all_ids = [t[0] for t in tax.by_id if t is not None]
    index: Dict[bytes, int] = {}
    for p in proteins:
        home = rng.choice(all_ids)
        path = tax.root_path(home)
        for i in range(len(p) - k + 1):
            u = rng.random()
            v = home if u < 0.7 else (rng.choice(path) if u < 0.9 else rng.choice(all_ids))
            key = p[i:i + k].encode()
            if key in index:
                a, b = tax.root_path(index[key]), tax.root_path(v)
                j = 0
                while j < min(len(a), len(b)) and a[j] == b[j]:
                    j += 1
                v = a[j - 1]
            index[key] = v
    return index


def revcomp(s: str) -> str:
    return s[::-1].translate(str.maketrans("ACGTN", "TGCAN"))


def make_reads(proteins: List[str], n_pairs: int, seed: int = 3, read_len: int = 150,
               hit_frac: float = 0.7) -> List[Tuple[str, str]]:
    """Paired reads `r<i>/1`, `r<i>/2`: fragments of one protein reverse-translated with random
    synonymous codons (random frame shift and strand, 1 % substitutions, 0.1 % N) or uniform
    random nucleotides."""
    rng = random.Random(seed)
    reads = []
    for i in range(n_pairs):
        hit = rng.random() < hit_frac
        prot = rng.choice(proteins)
        for mate in (1, 2):
            if hit and len(prot) * 3 >= read_len:
                naa = read_len // 3 + 1
                o = rng.randrange(0, max(1, len(prot) - naa + 1))
                frag = prot[o:o + naa]
                nt = "".join(rng.choice(CODONS.get(a, CODONS["A"])) for a in frag)
                sh = rng.randrange(0, 3)
                nt = ("".join(rng.choice("ACGT") for _ in range(sh)) + nt)[:read_len]
                nt = nt + "".join(rng.choice("ACGT") for _ in range(read_len - len(nt)))
                nt = list(nt)
                for x in range(read_len):
                    r = rng.random()
                    if r < 0.001:
                        nt[x] = "N"
                    elif r < 0.011:
                        nt[x] = rng.choice("ACGT")
                nt = "".join(nt)
                if rng.random() < 0.5:
                    nt = revcomp(nt)
            else:
                nt = "".join(rng.choice("ACGT") for _ in range(read_len))
            reads.append((f"r{i}/{mate}", nt))
    return reads
