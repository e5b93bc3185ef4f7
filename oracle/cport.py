"""ctypes loader for the C restatement (oracle/c/umgap_ref.c).  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c")
_LIB: Optional[C.CDLL] = None


class RefOpts(C.Structure):
    _fields_ = [("table", C.c_int), ("methionine", C.c_int), ("one_on_one", C.c_int), ("seedextend", C.c_int),
                ("min_seed_size", C.c_int), ("max_gap_size", C.c_int), ("strategy", C.c_int),
                ("factor", C.c_float), ("lower_bound", C.c_float), ("ranked_only", C.c_int), ("k", C.c_int)]


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        path = os.path.join(_DIR, "libumgap_ref.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-C", _DIR], check=True, stdout=subprocess.DEVNULL)
        l = C.CDLL(path)
        l.ref_fst_build.restype = C.c_void_p
        l.ref_fst_get.restype = C.c_int
        l.ref_tax_new.restype = C.c_void_p
        l.ref_seedextend.restype = C.c_uint32
        l.ref_aggregate.restype = C.c_uint32
        l.ref_translate.restype = C.c_uint32
        _LIB = l
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def fst_build(keys: Sequence[bytes], values: Sequence[int]) -> bytes:
    """Sorted keys -> fst Map image (bytes)."""
    off = np.zeros(len(keys) + 1, dtype=np.uint64)
    np.cumsum(np.fromiter((len(k) for k in keys), dtype=np.uint64, count=len(keys)), out=off[1:])
    blob = np.frombuffer(b"".join(keys) or b"\0", dtype=np.uint8)
    return fst_build_blob(blob, off, np.asarray(values, dtype=np.uint64))


def fst_build_blob(blob: np.ndarray, off: np.ndarray, values: np.ndarray) -> bytes:
    size = C.c_uint64()
    blob = np.ascontiguousarray(blob, dtype=np.uint8)
    off = np.ascontiguousarray(off, dtype=np.uint64)
    values = np.ascontiguousarray(values, dtype=np.uint64)
    ptr = lib().ref_fst_build(_p(blob), _p(off), _p(values), C.c_uint64(len(values)), C.byref(size))
    if not ptr:
        raise ValueError("fst build failed: keys must be strictly increasing")
    try:
        return C.string_at(ptr, size.value)
    finally:
        lib().ref_free(C.c_void_p(ptr))


class FstImage:
    def __init__(self, data: bytes):
        self.data = np.frombuffer(data, dtype=np.uint8)

    def get(self, key: bytes) -> Optional[int]:
        v = C.c_uint64()
        kb = np.frombuffer(key or b"\0", dtype=np.uint8)
        ok = lib().ref_fst_get(_p(self.data), C.c_uint64(len(self.data)), _p(kb), C.c_uint32(len(key)), C.byref(v))
        return v.value if ok else None


class RefTaxonomy:
    def __init__(self, taxa):
        ids = np.array([t[0] for t in taxa], dtype=np.uint64)
        parents = np.array([t[3] for t in taxa], dtype=np.uint64)
        ranks = np.array([t[2] for t in taxa], dtype=np.uint8)
        valid = np.array([1 if t[4] else 0 for t in taxa], dtype=np.uint8)
        self._h = C.c_void_p(lib().ref_tax_new(_p(ids), _p(parents), _p(ranks), _p(valid), C.c_uint64(len(ids))))

    def __del__(self):
        try:
            lib().ref_tax_free(self._h)
        except Exception:
            pass


def seedextend(ids, s: int, g: int):
    a = np.ascontiguousarray(ids, dtype=np.uint32)
    out = np.zeros(len(a) + 1, dtype=np.uint32)
    m = lib().ref_seedextend(_p(a), C.c_uint32(len(a)), C.c_uint32(s), C.c_uint32(g), _p(out))
    return [int(x) for x in out[:m]]


def aggregate(tax: RefTaxonomy, ids, strategy: int, factor: float = 0.25, lower_bound: float = 0.0,
              ranked: bool = False) -> int:
    a = np.ascontiguousarray(ids, dtype=np.uint32)
    bad = C.c_uint32()
    r = lib().ref_aggregate(tax._h, _p(a), C.c_uint32(len(a)), C.c_int(strategy), C.c_float(factor),
                            C.c_float(lower_bound), C.c_int(int(ranked)), C.byref(bad))
    if r == 0xFFFFFFFE:
        raise KeyError(f"Unknown Taxon ID: {bad.value}")
    return int(r)


def translate(nt: bytes, frame: int, table: int = 1, methionine: bool = False) -> str:
    a = np.frombuffer(nt or b"\0", dtype=np.uint8)
    out = np.zeros(len(nt) // 3 + 2, dtype=np.uint8)
    m = lib().ref_translate(C.c_int(table), C.c_int(int(methionine)), _p(a), C.c_uint32(len(nt)), C.c_int(frame), _p(out))
    if m == 0xFFFFFFFF:
        raise ValueError("Unknown table")
    return bytes(out[:m]).decode()


def classify(img: FstImage, tax: RefTaxonomy, opts: RefOpts, nt: np.ndarray, read_off: np.ndarray,
             group_off: np.ndarray, threads: int = 1, lookups_only: bool = False):
    """Returns (taxon per group [0xFFFFFFFF = no record], lookups, hits)."""
    nt = np.ascontiguousarray(nt, dtype=np.uint8)
    read_off = np.ascontiguousarray(read_off, dtype=np.uint64)
    group_off = np.ascontiguousarray(group_off, dtype=np.uint64)
    ng = len(group_off) - 1
    out = np.zeros(max(ng, 1), dtype=np.uint32)
    nl, nh, bad = C.c_uint64(), C.c_uint64(), C.c_uint32()
    rc = lib().ref_classify(_p(img.data), C.c_uint64(len(img.data)), tax._h, C.byref(opts), _p(nt), _p(read_off),
                            _p(group_off), C.c_uint64(ng), _p(out), C.c_int(threads), C.c_int(int(lookups_only)),
                            C.byref(nl), C.byref(nh), C.byref(bad))
    if rc == -4:
        raise KeyError(f"Unknown Taxon ID: {bad.value}")
    if rc != 0:
        raise ValueError("Unknown table")
    return out[:ng], nl.value, nh.value


class RefTrypOpts(C.Structure):
    _fields_ = [("minlen", C.c_int), ("maxlen", C.c_int), ("keep", C.c_char_p), ("drop", C.c_char_p), ("strategy", C.c_int),
                ("factor", C.c_float), ("lower_bound", C.c_float), ("ranked_only", C.c_int)]


def classify_peptides(img: FstImage, tax: RefTaxonomy, opts: RefTrypOpts, aa: np.ndarray, line_off: np.ndarray,
                      group_off: np.ndarray, threads: int = 1):
    """prot2tryp2lca | uniq -d / | taxa2agg per group of peptide lines: (taxon per group, lookups, hits)."""
    aa = np.ascontiguousarray(aa, dtype=np.uint8)
    line_off = np.ascontiguousarray(line_off, dtype=np.uint64)
    group_off = np.ascontiguousarray(group_off, dtype=np.uint64)
    ng = len(group_off) - 1
    out = np.zeros(max(ng, 1), dtype=np.uint32)
    nl, nh, bad = C.c_uint64(), C.c_uint64(), C.c_uint32()
    rc = lib().ref_classify_peptides(_p(img.data), C.c_uint64(len(img.data)), tax._h, C.byref(opts), _p(aa), _p(line_off),
                                     _p(group_off), C.c_uint64(ng), _p(out), C.c_int(threads), C.byref(nl), C.byref(nh),
                                     C.byref(bad))
    if rc == -4:
        raise KeyError(f"Unknown Taxon ID: {bad.value}")
    return out[:ng], nl.value, nh.value


def pipeline_staged(img: FstImage, tax: RefTaxonomy, opts: RefOpts, fasta: bytes, threads: int = 1):
    """The reference's five-process structure over in-memory text (ref_pipeline_staged): returns
    (taxon per uniq group, seconds per stage [translate, prot2kmer2lca, seedextend, uniq, taxa2agg], lookups)."""
    cap = fasta.count(b">") + 1
    out = np.zeros(cap, dtype=np.uint32)
    stage = (C.c_double * 5)()
    nl = C.c_uint64()
    lib().ref_pipeline_staged.restype = C.c_int64
    n = lib().ref_pipeline_staged(_p(img.data), C.c_uint64(len(img.data)), tax._h, C.byref(opts), C.c_char_p(fasta),
                                  C.c_uint64(len(fasta)), C.c_int(threads), _p(out), C.c_uint64(cap), stage, C.byref(nl))
    if n < 0:
        raise ValueError(f"ref_pipeline_staged failed: {n}")
    return out[:n], [float(x) for x in stage], nl.value


def synth_fst(seed: int, n_proteins: int, protein_len: int, home_pct: int, ancestor_pct: int, pre, threads: int = 0):
    """ref_synth_fst: the fst image of the synthetic proteome's 9-mer index (what oracle.synth.build_index +
    fst_build_blob give, generated, sorted and merged in C).  `pre`: oracle.synth.Preorder.  Returns (bytes, n_keys)."""
    size, nkeys = C.c_uint64(), C.c_uint64()
    id_of = np.ascontiguousarray(pre.id_arr, dtype=np.uint64)
    depth = np.asarray(pre.depth, dtype=np.uint32)
    parent = np.asarray(pre.parent_dense, dtype=np.uint32)
    lib().ref_synth_fst.restype = C.c_void_p
    ptr = lib().ref_synth_fst(C.c_uint64(seed), C.c_uint64(n_proteins), C.c_uint32(protein_len), C.c_uint32(home_pct),
                              C.c_uint32(ancestor_pct), _p(id_of), _p(depth), _p(parent), C.c_uint32(pre.n),
                              C.c_int(threads or (os.cpu_count() or 1)), C.byref(size), C.byref(nkeys))
    if not ptr:
        raise ValueError("ref_synth_fst failed")
    try:
        return C.string_at(ptr, size.value), nkeys.value
    finally:
        lib().ref_free(C.c_void_p(ptr))


def synth_reads(seed: int, n_proteins: int, protein_len: int, read_seed: int, first_pair: int, npairs: int, read_len: int,
                hit_pct: int, threads: int = 0) -> np.ndarray:
    """ref_synth_reads: uint8 [npairs*2, read_len], the C mirror of oracle.synth.reads."""
    out = np.empty((2 * npairs, read_len), dtype=np.uint8)
    lib().ref_synth_reads(C.c_uint64(seed), C.c_uint64(n_proteins), C.c_uint32(protein_len), C.c_uint64(read_seed),
                          C.c_uint64(first_pair), C.c_uint64(npairs), C.c_uint32(read_len), C.c_uint32(hit_pct), _p(out),
                          C.c_int(threads or (os.cpu_count() or 1)))
    return out
