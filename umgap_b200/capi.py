"""ctypes binding of ``libumgap_gpu.so`` (the C ABI declared in ``include/umgap_gpu.h``).

This is the thin host-side mirror the tests and ``bench.py`` use; it holds no algorithm.  Every
compute call goes to the CUDA library and raises :class:`UmgapError` when the library reports a
failure -- there is no CPU fallback here or anywhere else in the package.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("UMGAP_GPU_LIB") or os.path.join(_HERE, "lib", "libumgap_gpu.so")

MISS = 0xFFFFFFFF
ABSENT = 0xFFFFFFFF
AGG_LCA_STAR, AGG_HYBRID, AGG_MRTL = 0, 1, 2
AGG_RMQ_HYBRID = 3

# every symbol include/umgap_gpu.h declares (checked by tests/test_capi_symbols.py)
SYMBOLS = [
    "umgap_last_error", "umgap_abi_version", "umgap_device_count", "umgap_host_alloc", "umgap_host_free",
    "umgap_index_load_fst", "umgap_index_from_pairs", "umgap_index_free", "umgap_index_get_info",
    "umgap_index_set_probe_region", "umgap_index_build_from_proteins",
    "umgap_index_load_fst_shard", "umgap_index_from_pairs_shard", "umgap_index_shard_desc",
    "umgap_index_attach_shards", "umgap_index_attach_shards_local", "umgap_index_build_synthetic_shard",
    "umgap_taxonomy_load", "umgap_taxonomy_from_arrays", "umgap_taxonomy_free",
    "umgap_taxonomy_get_info",
    "umgap_translate_bound", "umgap_translate",
    "umgap_kmer_lookup_bound", "umgap_kmer_lookup",
    "umgap_tryp_lookup_bound", "umgap_tryp_lookup",
    "umgap_seedextend", "umgap_seedextend_ranked", "umgap_aggregate", "umgap_aggregate_scored",
    "umgap_fst_writer_open", "umgap_fst_writer_insert", "umgap_fst_writer_finish", "umgap_fst_writer_abort", "umgap_fst_stream",
    "umgap_kernel_times_ex", "umgap_exchange_bucket_cap", "umgap_exchange_region_bytes", "umgap_exchange_create", "umgap_exchange_create_lane", "umgap_exchange_free",
    "umgap_exchange_classify_dev", "umgap_exchange_status", "umgap_sharded_create", "umgap_sharded_free",
    "umgap_classify_reads_sharded_dev", "umgap_sharded_sync", "umgap_classify_reads_sharded",
    "umgap_index_replicate", "umgap_taxonomy_replicate", "umgap_classify_reads_multi", "umgap_classify_reads_packed_multi",
    "umgap_pipeline_opts_default", "umgap_classify_reads", "umgap_classify_reads_dev",
    "umgap_tryp_opts_default", "umgap_classify_peptides", "umgap_classify_peptides_dev",
    "umgap_translate_lookup_dev",
    "umgap_packed_words", "umgap_pack_reads", "umgap_classify_reads_packed",
    "umgap_classify_reads_async", "umgap_classify_reads_packed_async", "umgap_pending_wait",
    "umgap_route_pack_dev", "umgap_lookup_hashes_dev", "umgap_route_scatter_dev", "umgap_classify_ids_dev",
    "umgap_route_sampled_applies", "umgap_route_pack_sampled_dev", "umgap_route_scatter_hits_dev", "umgap_classify_ids_masked_dev",
    "umgap_kernel_timing", "umgap_kernel_times", "umgap_kernel_launch_count", "umgap_transfer_bytes", "umgap_pipeline_slices", "umgap_pipeline_sampling",
    "umgap_index_build_synthetic", "umgap_synth_reads_dev", "umgap_randsector_bench",
]


class UmgapError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"[{code}] {message}")
        self.code = code
        self.message = message


class IndexInfo(C.Structure):
    _fields_ = [("n_keys", C.c_uint64), ("n_buckets", C.c_uint64), ("bytes", C.c_uint64),
                ("n_skipped", C.c_uint64), ("n_flagged", C.c_uint64), ("n_displaced", C.c_uint64),
                ("max_probe", C.c_uint64), ("k", C.c_int), ("device", C.c_int),
                ("alphabet_size", C.c_int), ("load_factor", C.c_double)]


class TaxonomyInfo(C.Structure):
    _fields_ = [("n_taxa", C.c_uint64), ("max_id", C.c_uint64), ("root", C.c_uint64),
                ("max_depth", C.c_uint32), ("device", C.c_int)]


class PipelineOpts(C.Structure):
    _fields_ = [("table", C.c_int), ("methionine", C.c_int), ("one_on_one", C.c_int),
                ("seedextend", C.c_int), ("min_seed_size", C.c_int), ("max_gap_size", C.c_int),
                ("strategy", C.c_int), ("factor", C.c_float), ("lower_bound", C.c_float),
                ("ranked_only", C.c_int)]


class TrypOpts(C.Structure):
    _fields_ = [("minlen", C.c_int), ("maxlen", C.c_int), ("keep", C.c_char_p), ("drop", C.c_char_p),
                ("strategy", C.c_int), ("factor", C.c_float), ("lower_bound", C.c_float), ("ranked_only", C.c_int)]


class ShardDesc(C.Structure):
    _fields_ = [("ipc", (C.c_ubyte * 64) * 4), ("nlines", C.c_uint32 * 4), ("nlevels", C.c_int), ("shard", C.c_int),
                ("nshards", C.c_int), ("device", C.c_int), ("alphabet_size", C.c_int),
                ("code_of_byte", C.c_ubyte * 256)]


class SynthSpec(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("n_proteins", C.c_uint64), ("protein_len", C.c_uint32),
                ("home_pct", C.c_uint32), ("ancestor_pct", C.c_uint32)]


_lib: Optional[C.CDLL] = None


def load_library() -> C.CDLL:
    """Loads the CUDA library; fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise UmgapError(-3, f"{LIB_PATH} is missing: run `make` (or __graft_entry__.build()); "
                             "umgap_b200 has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    lib.umgap_last_error.restype = C.c_char_p
    lib.umgap_translate_bound.restype = C.c_uint64
    lib.umgap_translate_bound.argtypes = [C.c_uint64, C.c_uint64, C.c_uint8]
    lib.umgap_kmer_lookup_bound.restype = C.c_uint64
    lib.umgap_kmer_lookup_bound.argtypes = [C.c_uint64, C.c_uint64]
    lib.umgap_tryp_lookup_bound.restype = C.c_uint64
    lib.umgap_tryp_lookup_bound.argtypes = [C.c_uint64, C.c_uint64]
    _lib = lib
    return lib


def _check(rc: int) -> None:
    if rc != 0:
        raise UmgapError(rc, load_library().umgap_last_error().decode("utf-8", "replace"))


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _arr(x, dtype) -> np.ndarray:
    return np.ascontiguousarray(x, dtype=dtype)


def device_count() -> int:
    return load_library().umgap_device_count()


class Taxonomy:
    """GPU-resident taxonomy (taxon.rs TaxonList/TaxonTree + snapping)."""

    def __init__(self, handle):
        self._h = handle

    @classmethod
    def from_arrays(cls, ids, parents, ranks, valid, device: int = 0) -> "Taxonomy":
        ids = _arr(ids, np.uint64)
        parents = _arr(parents, np.uint64)
        ranks = _arr(ranks, np.uint8)
        valid = _arr(valid, np.uint8)
        h = C.c_void_p()
        _check(load_library().umgap_taxonomy_from_arrays(_p(ids), _p(parents), _p(ranks), _p(valid),
                                                         C.c_uint64(len(ids)), C.c_int(device),
                                                         C.byref(h)))
        return cls(h)

    @classmethod
    def load(cls, path: str, device: int = 0) -> "Taxonomy":
        h = C.c_void_p()
        _check(load_library().umgap_taxonomy_load(path.encode(), C.c_int(device), C.byref(h)))
        return cls(h)

    def info(self) -> TaxonomyInfo:
        i = TaxonomyInfo()
        _check(load_library().umgap_taxonomy_get_info(self._h, C.byref(i)))
        return i

    def close(self):
        if self._h:
            load_library().umgap_taxonomy_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Index:
    """GPU-resident key -> taxon table (fst::Map replacement)."""

    def __init__(self, handle):
        self._h = handle

    @classmethod
    def from_pairs(cls, keys: Sequence[bytes], values, k: int = 9, device: int = 0,
                   load_factor: float = 0.0, shard: int = 0, nshards: int = 1) -> "Index":
        lens = np.fromiter((len(x) for x in keys), dtype=np.uint64, count=len(keys))
        off = np.zeros(len(keys) + 1, dtype=np.uint64)
        np.cumsum(lens, out=off[1:])
        blob = np.frombuffer(b"".join(keys) or b"\0", dtype=np.uint8)
        return cls.from_blob(blob, off, values, k, device, load_factor, shard, nshards)

    @classmethod
    def from_blob(cls, blob: np.ndarray, off: Optional[np.ndarray], values, k: int = 9,
                  device: int = 0, load_factor: float = 0.0, shard: int = 0, nshards: int = 1) -> "Index":
        blob = _arr(blob, np.uint8)
        values = _arr(values, np.uint64)
        off = None if off is None else _arr(off, np.uint64)
        h = C.c_void_p()
        _check(load_library().umgap_index_from_pairs_shard(_p(blob), _p(off), _p(values),
                                                           C.c_uint64(len(values)), C.c_int(k),
                                                           C.c_int(device), C.c_double(load_factor),
                                                           C.c_int(shard), C.c_int(nshards), C.byref(h)))
        return cls(h)

    @classmethod
    def build_from_proteins(cls, tax: "Taxonomy", proteins: Sequence[bytes], taxa, k: int = 9,
                            load_factor: float = 0.0) -> "Index":
        """umgap_index_build_from_proteins: splitkmers | sort | joinkmers | buildindex on the device."""
        aa, off = pack_strings(list(proteins))
        taxa = _arr(np.asarray(taxa, dtype=np.uint64), np.uint64)
        h = C.c_void_p()
        _check(load_library().umgap_index_build_from_proteins(tax._h, _p(aa), _p(off), _p(taxa), C.c_uint64(len(taxa)),
                                                              C.c_int(k), C.c_double(load_factor), C.byref(h)))
        return cls(h)

    @classmethod
    def load_fst(cls, path: str, k: int = 9, device: int = 0, load_factor: float = 0.0, shard: int = 0,
                 nshards: int = 1) -> "Index":
        h = C.c_void_p()
        _check(load_library().umgap_index_load_fst_shard(path.encode(), C.c_int(k), C.c_int(device),
                                                         C.c_double(load_factor), C.c_int(shard),
                                                         C.c_int(nshards), C.byref(h)))
        return cls(h)

    @classmethod
    def build_synthetic(cls, spec: SynthSpec, tax: Taxonomy, device: int = 0,
                        load_factor: float = 0.0, shard: int = 0, nshards: int = 1) -> "Index":
        h = C.c_void_p()
        _check(load_library().umgap_index_build_synthetic_shard(C.byref(spec), tax._h, C.c_int(device),
                                                                C.c_double(load_factor), C.c_int(shard),
                                                                C.c_int(nshards), C.byref(h)))
        return cls(h)

    def shard_desc(self) -> ShardDesc:
        """Descriptor of this rank's shard (CUDA IPC handles of its levels) for the peers."""
        d = ShardDesc()
        _check(load_library().umgap_index_shard_desc(self._h, C.byref(d)))
        return d

    def attach_shards(self, descs: Sequence[ShardDesc]) -> None:
        """Maps every shard (rank order) so that lookups read remote shards over NVLink."""
        arr = (ShardDesc * len(descs))(*descs)
        _check(load_library().umgap_index_attach_shards(self._h, arr, C.c_int(len(descs))))

    def attach_shards_local(self, shards: Sequence["Index"]) -> None:
        """Same-process variant: the shard handles themselves, in shard order."""
        arr = (C.c_void_p * len(shards))(*[s._h for s in shards])
        _check(load_library().umgap_index_attach_shards_local(self._h, arr, C.c_int(len(shards))))

    def info(self) -> IndexInfo:
        i = IndexInfo()
        _check(load_library().umgap_index_get_info(self._h, C.byref(i)))
        return i

    def set_probe_region(self, nbytes: int) -> None:
        _check(load_library().umgap_index_set_probe_region(self._h, C.c_uint64(nbytes)))

    def randsector_rate(self, n_gathers: int, iters: int = 5) -> float:
        r = C.c_double()
        _check(load_library().umgap_randsector_bench(self._h, C.c_uint64(n_gathers), C.c_int(iters),
                                                     C.byref(r)))
        return r.value

    def close(self):
        if self._h:
            load_library().umgap_index_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def fst_write(path: str, items) -> None:
    """umgap_fst_writer_*: (key bytes, value) pairs in strictly increasing key order -> an fst Map file."""
    lib = load_library()
    w = C.c_void_p()
    _check(lib.umgap_fst_writer_open(path.encode(), C.byref(w)))
    try:
        for k, v in items:
            _check(lib.umgap_fst_writer_insert(w, C.c_char_p(bytes(k)), C.c_size_t(len(k)), C.c_uint64(int(v))))
    except Exception:
        lib.umgap_fst_writer_abort(w)
        raise
    _check(lib.umgap_fst_writer_finish(w))


def fst_items(path: str):
    """umgap_fst_stream: the (key, value) pairs of an fst Map file, in order."""
    out = []
    CB = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_uint8), C.c_size_t, C.c_uint64, C.c_void_p)

    def cb(key, n, value, _user):
        out.append((bytes(key[:n]), int(value)))
        return 0

    n = C.c_uint64()
    _check(load_library().umgap_fst_stream(path.encode(), CB(cb), None, C.byref(n)))
    assert n.value == len(out)
    return out


def default_opts(**kw) -> PipelineOpts:
    o = PipelineOpts()
    load_library().umgap_pipeline_opts_default(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise AttributeError(k)
        setattr(o, k, v)
    return o


def _offsets(lengths) -> np.ndarray:
    off = np.zeros(len(lengths) + 1, dtype=np.uint64)
    np.cumsum(np.asarray(lengths, dtype=np.uint64), out=off[1:])
    return off


def pack_strings(items: Sequence[bytes]):
    """Concatenates byte strings -> (uint8 blob, uint64 offsets)."""
    off = _offsets([len(x) for x in items])
    blob = np.frombuffer(b"".join(items) or b"\0", dtype=np.uint8).copy()
    return blob, off


def translate(nt: np.ndarray, read_off: np.ndarray, table: int = 1, methionine: bool = False,
              frames_mask: int = 0x3F, device: int = 0):
    """umgap_translate: returns (aa blob, aa offsets) with one peptide per (read, frame)."""
    lib = load_library()
    nt = _arr(nt, np.uint8)
    read_off = _arr(read_off, np.uint64)
    nreads = len(read_off) - 1
    nframes = bin(frames_mask & 0x3F).count("1")
    bound = lib.umgap_translate_bound(C.c_uint64(int(read_off[-1])), C.c_uint64(nreads),
                                      C.c_uint8(frames_mask))
    aa = np.zeros(max(int(bound), 1), dtype=np.uint8)
    aa_off = np.zeros(nreads * nframes + 1, dtype=np.uint64)
    _check(lib.umgap_translate(C.c_int(device), _p(nt), _p(read_off), C.c_uint64(nreads),
                               C.c_int(table), C.c_int(int(methionine)), C.c_uint8(frames_mask),
                               _p(aa), _p(aa_off)))
    return aa[: int(aa_off[-1])], aa_off


def kmer_lookup(index: Index, aa: np.ndarray, pep_off: np.ndarray, one_on_one: bool):
    """umgap_kmer_lookup: returns (taxa, taxa offsets, kept flags)."""
    lib = load_library()
    aa = _arr(aa, np.uint8)
    pep_off = _arr(pep_off, np.uint64)
    npeps = len(pep_off) - 1
    bound = lib.umgap_kmer_lookup_bound(C.c_uint64(int(pep_off[-1])), C.c_uint64(npeps))
    taxa = np.zeros(max(int(bound), 1), dtype=np.uint32)
    taxa_off = np.zeros(npeps + 1, dtype=np.uint64)
    kept = np.zeros(max(npeps, 1), dtype=np.uint8)
    _check(lib.umgap_kmer_lookup(index._h, _p(aa), _p(pep_off), C.c_uint64(npeps),
                                 C.c_int(int(one_on_one)), _p(taxa), _p(taxa_off), _p(kept)))
    return taxa[: int(taxa_off[-1])], taxa_off, kept[:npeps]


def tryp_lookup(index: Index, aa: np.ndarray, line_off: np.ndarray, minlen: int = 5,
                maxlen: int = 50, keep: str = "", drop: str = "", one_on_one: bool = False):
    lib = load_library()
    aa = _arr(aa, np.uint8)
    line_off = _arr(line_off, np.uint64)
    nlines = len(line_off) - 1
    bound = lib.umgap_tryp_lookup_bound(C.c_uint64(int(line_off[-1])), C.c_uint64(nlines))
    taxa = np.zeros(max(int(bound), 1), dtype=np.uint32)
    taxa_off = np.zeros(nlines + 1, dtype=np.uint64)
    _check(lib.umgap_tryp_lookup(index._h, _p(aa), _p(line_off), C.c_uint64(nlines),
                                 C.c_int(minlen), C.c_int(maxlen), keep.encode(), drop.encode(),
                                 C.c_int(int(one_on_one)), _p(taxa), _p(taxa_off)))
    return taxa[: int(taxa_off[-1])], taxa_off


def seedextend(taxa: np.ndarray, rec_off: np.ndarray, min_seed_size: int = 2,
               max_gap_size: int = 0, device: int = 0):
    lib = load_library()
    taxa = _arr(taxa, np.uint32)
    rec_off = _arr(rec_off, np.uint64)
    nrecs = len(rec_off) - 1
    out = np.zeros(max(len(taxa), 1), dtype=np.uint32)
    out_off = np.zeros(nrecs + 1, dtype=np.uint64)
    _check(lib.umgap_seedextend(C.c_int(device), _p(taxa), _p(rec_off), C.c_uint64(nrecs),
                                C.c_int(min_seed_size), C.c_int(max_gap_size), _p(out),
                                _p(out_off)))
    return out[: int(out_off[-1])], out_off


def seedextend_ranked(tax: Taxonomy, taxa: np.ndarray, rec_off: np.ndarray, min_seed_size: int = 2, max_gap_size: int = 0,
                      penalty: int = 5):
    """umgap_seedextend_ranked: (kept ids, offsets)."""
    taxa = _arr(taxa, np.uint32)
    rec_off = _arr(rec_off, np.uint64)
    nrecs = len(rec_off) - 1
    out = np.zeros(max(int(rec_off[-1]), 1), dtype=np.uint32)
    out_off = np.zeros(nrecs + 1, dtype=np.uint64)
    _check(load_library().umgap_seedextend_ranked(tax._h, _p(taxa), _p(rec_off), C.c_uint64(nrecs), C.c_int(min_seed_size),
                                                  C.c_int(max_gap_size), C.c_int(penalty), _p(out), _p(out_off)))
    return out[:int(out_off[-1])], out_off


def aggregate(tax: Taxonomy, taxa: np.ndarray, rec_off: np.ndarray, strategy: int,
              factor: float = 0.25, lower_bound: float = 0.0, ranked_only: bool = False):
    lib = load_library()
    taxa = _arr(taxa, np.uint32)
    rec_off = _arr(rec_off, np.uint64)
    nrecs = len(rec_off) - 1
    out = np.zeros(max(nrecs, 1), dtype=np.uint32)
    _check(lib.umgap_aggregate(tax._h, _p(taxa), _p(rec_off), C.c_uint64(nrecs), C.c_int(strategy),
                               C.c_float(factor), C.c_float(lower_bound),
                               C.c_int(int(ranked_only)), _p(out)))
    return out[:nrecs]


def aggregate_scored(tax: Taxonomy, taxa: np.ndarray, scores: np.ndarray, rec_off: np.ndarray, strategy: int,
                     factor: float = 0.25, lower_bound: float = 0.0, ranked_only: bool = False):
    """umgap_aggregate_scored (taxa2agg -s)."""
    taxa = _arr(taxa, np.uint32)
    scores = _arr(scores, np.float32)
    rec_off = _arr(rec_off, np.uint64)
    nrecs = len(rec_off) - 1
    out = np.zeros(max(nrecs, 1), dtype=np.uint32)
    _check(load_library().umgap_aggregate_scored(tax._h, _p(taxa), _p(scores), _p(rec_off), C.c_uint64(nrecs), C.c_int(strategy),
                                                 C.c_float(factor), C.c_float(lower_bound), C.c_int(int(ranked_only)), _p(out)))
    return out[:nrecs]


def tryp_opts(**kw) -> TrypOpts:
    """umgap_tryp_opts_default plus overrides (keep / drop as str or bytes)."""
    o = TrypOpts()
    load_library().umgap_tryp_opts_default(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise AttributeError(k)
        if k in ("keep", "drop") and isinstance(v, str):
            v = v.encode()
        setattr(o, k, v)
    return o


def classify_peptides(index: Index, tax: Taxonomy, opts: TrypOpts, aa: np.ndarray, line_off: np.ndarray,
                      group_off: np.ndarray, out: Optional[np.ndarray] = None) -> np.ndarray:
    """umgap_classify_peptides (host buffers): prot2tryp2lca | uniq | taxa2agg, one taxon per group of lines.
    `out` may be a caller-provided (page-locked) uint32 array of at least ngroups entries."""
    aa = _arr(aa, np.uint8)
    line_off = _arr(line_off, np.uint64)
    group_off = _arr(group_off, np.uint64)
    ngroups = len(group_off) - 1
    if out is None:
        out = np.zeros(max(ngroups, 1), dtype=np.uint32)
    _check(load_library().umgap_classify_peptides(index._h, tax._h, C.byref(opts), _p(aa), _p(line_off),
                                                  C.c_uint64(len(line_off) - 1), _p(group_off), C.c_uint64(ngroups), _p(out)))
    return out[:ngroups]


def classify_peptides_dev(index: Index, tax: Taxonomy, opts: TrypOpts, aa_ptr: int, line_off_ptr: int, nlines: int,
                          total_aa: int, group_off_ptr: int, ngroups: int, out_ptr: int, stream: int = 0) -> None:
    _check(load_library().umgap_classify_peptides_dev(
        index._h, tax._h, C.byref(opts), C.c_void_p(aa_ptr), C.c_void_p(line_off_ptr), C.c_uint64(nlines), C.c_uint64(total_aa),
        C.c_void_p(group_off_ptr), C.c_uint64(ngroups), C.c_void_p(out_ptr), C.c_void_p(stream)))


def classify_reads(index: Index, tax: Taxonomy, opts: PipelineOpts, nt: np.ndarray,
                   read_off: np.ndarray, group_off: np.ndarray, count_lookups: bool = True,
                   out: Optional[np.ndarray] = None):
    """umgap_classify_reads (host buffers): returns (taxon per group, number of lookups or None).
    `out` may be a caller-provided (e.g. pinned) uint32 array of at least ngroups entries."""
    lib = load_library()
    nt = _arr(nt, np.uint8)
    read_off = _arr(read_off, np.uint64)
    group_off = _arr(group_off, np.uint64)
    ngroups = len(group_off) - 1
    if out is None:
        out = np.zeros(max(ngroups, 1), dtype=np.uint32)
    nl = C.c_uint64()
    _check(lib.umgap_classify_reads(index._h, tax._h, C.byref(opts), _p(nt), _p(read_off),
                                    C.c_uint64(len(read_off) - 1), _p(group_off),
                                    C.c_uint64(ngroups), _p(out), C.byref(nl) if count_lookups else None))
    return out[:ngroups], (nl.value if count_lookups else None)


def pack_reads(nt: np.ndarray, threads: int = 0, codes: Optional[np.ndarray] = None, entries: Optional[np.ndarray] = None):
    """umgap_pack_reads: nucleotide bytes -> (codes uint32[words], N entries uint64[count]), 16 nucleotides per word.
    `codes` / `entries`: optional arrays to fill (e.g. views of page-locked memory)."""
    lib = load_library()
    lib.umgap_packed_words.restype = C.c_uint64
    nt = _arr(nt, np.uint8)
    nw = int(lib.umgap_packed_words(C.c_uint64(len(nt))))
    if codes is None:
        codes = np.zeros(max(nw, 1), dtype=np.uint32)
    if entries is None:
        entries = np.zeros(max(nw, 1), dtype=np.uint64)
    assert len(codes) >= nw
    n = C.c_uint64()
    _check(lib.umgap_pack_reads(_p(nt), C.c_uint64(len(nt)), _p(codes), _p(entries), C.c_uint64(len(entries)), C.byref(n), C.c_int(threads)))
    return codes, entries[: n.value]


def classify_reads_packed(index: Index, tax: Taxonomy, opts: PipelineOpts, codes: np.ndarray, entries: Optional[np.ndarray],
                          read_off: np.ndarray, group_off: np.ndarray, count_lookups: bool = True,
                          out: Optional[np.ndarray] = None):
    """umgap_classify_reads_packed (host buffers, 2-bit nucleotides + the list of words holding an N)."""
    lib = load_library()
    read_off = _arr(read_off, np.uint64)
    group_off = _arr(group_off, np.uint64)
    ngroups = len(group_off) - 1
    if out is None:
        out = np.zeros(max(ngroups, 1), dtype=np.uint32)
    nl = C.c_uint64()
    ne = 0 if entries is None else len(entries)
    _check(lib.umgap_classify_reads_packed(index._h, tax._h, C.byref(opts), _p(codes), _p(entries) if ne else None, C.c_uint64(ne),
                                           _p(read_off), C.c_uint64(len(read_off) - 1), _p(group_off), C.c_uint64(ngroups),
                                           _p(out), C.byref(nl) if count_lookups else None))
    return out[:ngroups], (nl.value if count_lookups else None)


class Pending:
    """A batch enqueued by classify_reads_async / classify_reads_packed_async; wait() completes it and returns the
    per-group taxa.  Keeps the host arrays of the batch alive until then."""

    def __init__(self, handle, out, ngroups, keep):
        self._h, self._out, self._n, self._keep = handle, out, ngroups, keep

    def wait(self) -> np.ndarray:
        h, self._h = self._h, None
        if h is not None:
            self._keep = None
            _check(load_library().umgap_pending_wait(h))
        return self._out[: self._n]

    def __del__(self):
        if getattr(self, "_h", None) is not None:
            try:
                load_library().umgap_pending_wait(self._h)
            except Exception:
                pass
            self._h = None


def classify_reads_async(index: Index, tax: Taxonomy, opts: PipelineOpts, nt: np.ndarray, read_off: np.ndarray,
                         group_off: np.ndarray, out: Optional[np.ndarray] = None) -> Pending:
    """umgap_classify_reads_async: enqueue the batch, return at once (page-locked arrays make the copies asynchronous)."""
    nt = _arr(nt, np.uint8)
    read_off = _arr(read_off, np.uint64)
    group_off = _arr(group_off, np.uint64)
    ngroups = len(group_off) - 1
    if out is None:
        out = np.zeros(max(ngroups, 1), dtype=np.uint32)
    h = C.c_void_p()
    _check(load_library().umgap_classify_reads_async(index._h, tax._h, C.byref(opts), _p(nt), _p(read_off), C.c_uint64(len(read_off) - 1),
                                                     _p(group_off), C.c_uint64(ngroups), _p(out), C.byref(h)))
    return Pending(h, out, ngroups, (nt, read_off, group_off, opts))


def classify_reads_packed_async(index: Index, tax: Taxonomy, opts: PipelineOpts, codes: np.ndarray, entries: Optional[np.ndarray],
                                read_off: np.ndarray, group_off: np.ndarray, out: Optional[np.ndarray] = None) -> Pending:
    """umgap_classify_reads_packed_async."""
    read_off = _arr(read_off, np.uint64)
    group_off = _arr(group_off, np.uint64)
    ngroups = len(group_off) - 1
    if out is None:
        out = np.zeros(max(ngroups, 1), dtype=np.uint32)
    ne = 0 if entries is None else len(entries)
    h = C.c_void_p()
    _check(load_library().umgap_classify_reads_packed_async(index._h, tax._h, C.byref(opts), _p(codes), _p(entries) if ne else None,
                                                            C.c_uint64(ne), _p(read_off), C.c_uint64(len(read_off) - 1), _p(group_off),
                                                            C.c_uint64(ngroups), _p(out), C.byref(h)))
    return Pending(h, out, ngroups, (codes, entries, read_off, group_off, opts))


def replicate(index: Index, tax: Taxonomy, devices: Sequence[int]):
    """Replicas of a loaded table and taxonomy on `devices` (umgap_index_replicate / umgap_taxonomy_replicate)."""
    lib = load_library()
    out = []
    for d in devices:
        hi, ht = C.c_void_p(), C.c_void_p()
        _check(lib.umgap_index_replicate(index._h, C.c_int(d), C.byref(hi)))
        _check(lib.umgap_taxonomy_replicate(tax._h, C.c_int(d), C.byref(ht)))
        out.append((Index(hi), Taxonomy(ht)))
    return out


def classify_reads_multi(replicas, opts: PipelineOpts, nt: np.ndarray, read_off: np.ndarray, group_off: np.ndarray,
                         count_lookups: bool = True, out: Optional[np.ndarray] = None, packed=None):
    """umgap_classify_reads_multi over [(Index, Taxonomy), ...] (one pair per GPU); packed = (codes, entries) selects
    umgap_classify_reads_packed_multi."""
    lib = load_library()
    read_off = _arr(read_off, np.uint64)
    group_off = _arr(group_off, np.uint64)
    ngroups = len(group_off) - 1
    if out is None:
        out = np.zeros(max(ngroups, 1), dtype=np.uint32)
    n = len(replicas)
    ih = (C.c_void_p * n)(*[i._h for i, _ in replicas])
    th = (C.c_void_p * n)(*[t._h for _, t in replicas])
    nl = C.c_uint64()
    if packed is None:
        nt = _arr(nt, np.uint8)
        _check(lib.umgap_classify_reads_multi(ih, th, C.c_int(n), C.byref(opts), _p(nt), _p(read_off), C.c_uint64(len(read_off) - 1),
                                              _p(group_off), C.c_uint64(ngroups), _p(out), C.byref(nl) if count_lookups else None))
    else:
        codes, entries = packed
        ne = 0 if entries is None else len(entries)
        _check(lib.umgap_classify_reads_packed_multi(ih, th, C.c_int(n), C.byref(opts), _p(codes), _p(entries) if ne else None,
                                                     C.c_uint64(ne), _p(read_off), C.c_uint64(len(read_off) - 1), _p(group_off),
                                                     C.c_uint64(ngroups), _p(out), C.byref(nl) if count_lookups else None))
    return out[:ngroups], (nl.value if count_lookups else None)


class Sharded:
    """umgap_sharded: one process driving every shard of a key-range-sharded index; the exchange step runs in the
    kernels over NVLink peer mappings (exchange.cu).  shards[i] = shard i of n on its own GPU, taxa[i] on the same GPU."""

    def __init__(self, shards: Sequence[Index], taxa: Sequence[Taxonomy], max_total_nt: int):
        n = len(shards)
        self.n = n
        self._keep = (list(shards), list(taxa))
        ih = (C.c_void_p * n)(*[i._h for i in shards])
        th = (C.c_void_p * n)(*[t._h for t in taxa])
        h = C.c_void_p()
        _check(load_library().umgap_sharded_create(ih, th, C.c_int(n), C.c_uint64(max_total_nt), C.byref(h)))
        self._h = h

    def classify_reads(self, opts: PipelineOpts, nt: np.ndarray, read_off: np.ndarray, group_off: np.ndarray,
                       count_lookups: bool = True):
        nt = _arr(nt, np.uint8)
        read_off = _arr(read_off, np.uint64)
        group_off = _arr(group_off, np.uint64)
        ngroups = len(group_off) - 1
        out = np.zeros(max(ngroups, 1), dtype=np.uint32)
        nl = C.c_uint64()
        _check(load_library().umgap_classify_reads_sharded(self._h, C.byref(opts), _p(nt), _p(read_off), C.c_uint64(len(read_off) - 1),
                                                           _p(group_off), C.c_uint64(ngroups), _p(out),
                                                           C.byref(nl) if count_lookups else None))
        return out[:ngroups], (nl.value if count_lookups else None)

    def classify_reads_dev(self, opts: PipelineOpts, nt_ptrs, read_off_ptrs, nreads, total_nt, group_off_ptrs, ngroups, out_ptrs) -> None:
        """One device-resident batch per shard (lists of raw device addresses and counts); returns at once."""
        n = self.n
        vp = lambda xs: (C.c_void_p * n)(*[C.c_void_p(int(x)) for x in xs])
        u64 = lambda xs: (C.c_uint64 * n)(*[int(x) for x in xs])
        _check(load_library().umgap_classify_reads_sharded_dev(self._h, C.byref(opts), vp(nt_ptrs), vp(read_off_ptrs), u64(nreads),
                                                               u64(total_nt), vp(group_off_ptrs), u64(ngroups), vp(out_ptrs)))

    def sync(self) -> int:
        """Waits for the enqueued batches; raises on bucket overflow / a silent peer / Unknown Taxon ID.  Returns the
        number of lookups routed since the last call."""
        r = C.c_uint64()
        _check(load_library().umgap_sharded_sync(self._h, C.byref(r)))
        return r.value

    def close(self):
        if self._h:
            load_library().umgap_sharded_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def classify_reads_dev(index: Index, tax: Taxonomy, opts: PipelineOpts, nt_ptr: int,
                       read_off_ptr: int, nreads: int, total_nt: int, group_off_ptr: int,
                       ngroups: int, out_ptr: int, stream: int = 0) -> None:
    """Device-resident variant; pointers are raw device addresses (e.g. torch .data_ptr())."""
    _check(load_library().umgap_classify_reads_dev(
        index._h, tax._h, C.byref(opts), C.c_void_p(nt_ptr), C.c_void_p(read_off_ptr),
        C.c_uint64(nreads), C.c_uint64(total_nt), C.c_void_p(group_off_ptr), C.c_uint64(ngroups),
        C.c_void_p(out_ptr), C.c_void_p(stream)))


def translate_lookup_dev(index: Index, opts: PipelineOpts, nt_ptr: int, read_off_ptr: int,
                         nreads: int, total_nt: int, ids_ptr: int, stream: int = 0) -> None:
    _check(load_library().umgap_translate_lookup_dev(
        index._h, C.byref(opts), C.c_void_p(nt_ptr), C.c_void_p(read_off_ptr), C.c_uint64(nreads),
        C.c_uint64(total_nt), C.c_void_p(ids_ptr), C.c_void_p(stream)))


def synth_reads_dev(spec: SynthSpec, read_seed: int, first_pair: int, npairs: int, read_len: int,
                    hit_pct: int, nt_ptr: int, stream: int = 0) -> None:
    _check(load_library().umgap_synth_reads_dev(
        C.byref(spec), C.c_uint64(read_seed), C.c_uint64(first_pair), C.c_uint64(npairs),
        C.c_uint32(read_len), C.c_uint32(hit_pct), C.c_void_p(nt_ptr), C.c_void_p(stream)))


def kernel_timing(enable: bool) -> None:
    _check(load_library().umgap_kernel_timing(C.c_int(int(enable))))


def kernel_times():
    """(lookup_ms, lookup_launches, classify_ms, classify_launches) since the last call."""
    a, b = C.c_double(), C.c_double()
    na, nb = C.c_uint64(), C.c_uint64()
    _check(load_library().umgap_kernel_times(C.byref(a), C.byref(na), C.byref(b), C.byref(nb)))
    return a.value, na.value, b.value, nb.value


def kernel_times_ex():
    """umgap_kernel_times_ex: dict stage -> (ms, launches) since the last call."""
    ms = (C.c_double * 8)()
    cnt = (C.c_uint64 * 8)()
    _check(load_library().umgap_kernel_times_ex(ms, cnt, C.c_int(8)))
    names = ("lookup", "classify", "pack", "wait_for_peers", "scatter")
    return {n: (ms[i], cnt[i]) for i, n in enumerate(names)}


def pipeline_slices(slices: int = 0) -> int:
    """Sets (slices > 0) the number of slices of classify_reads_dev; returns the previous value."""
    return int(load_library().umgap_pipeline_slices(C.c_int(slices)))


def transfer_bytes():
    """(h2d, d2h) bytes classify_reads has moved over PCIe in this process so far."""
    a, b = C.c_uint64(), C.c_uint64()
    _check(load_library().umgap_transfer_bytes(C.byref(a), C.byref(b)))
    return a.value, b.value


def pipeline_sampling(enable: int = -1) -> int:
    """Switches the sampled lookups of the fused path on (1) / off (0); returns the previous setting."""
    return int(load_library().umgap_pipeline_sampling(C.c_int(enable)))


def kernel_launch_count() -> int:
    """Kernels launched by the fused path in this process so far."""
    n = C.c_uint64()
    _check(load_library().umgap_kernel_launch_count(C.byref(n)))
    return n.value


def route_pack_dev(index: Index, opts: PipelineOpts, nt_ptr: int, read_off_ptr: int, nreads: int, total_nt: int,
                   cap: int, send_h_ptr: int, send_pos_ptr: int, cursors_ptr: int, ids_ptr: int, stream: int = 0) -> None:
    _check(load_library().umgap_route_pack_dev(
        index._h, C.byref(opts), C.c_void_p(nt_ptr), C.c_void_p(read_off_ptr), C.c_uint64(nreads), C.c_uint64(total_nt),
        C.c_uint64(cap), C.c_void_p(send_h_ptr), C.c_void_p(send_pos_ptr), C.c_void_p(cursors_ptr), C.c_void_p(ids_ptr),
        C.c_void_p(stream)))


def lookup_hashes_dev(index: Index, h_ptr: int, counts_ptr: int, nsrc: int, cap: int, out_ptr: int, stream: int = 0) -> None:
    _check(load_library().umgap_lookup_hashes_dev(index._h, C.c_void_p(h_ptr), C.c_void_p(counts_ptr), C.c_int(nsrc),
                                                  C.c_uint64(cap), C.c_void_p(out_ptr), C.c_void_p(stream)))


def route_scatter_dev(index: Index, ans_ptr: int, send_pos_ptr: int, cursors_ptr: int, cap: int, ids_ptr: int,
                      stream: int = 0) -> None:
    _check(load_library().umgap_route_scatter_dev(index._h, C.c_void_p(ans_ptr), C.c_void_p(send_pos_ptr),
                                                  C.c_void_p(cursors_ptr), C.c_uint64(cap), C.c_void_p(ids_ptr),
                                                  C.c_void_p(stream)))


def route_sampled_applies(index: Index, opts: PipelineOpts) -> bool:
    """True when the sampled form of the exchange step is exact for these options (k = 9, -o, seedextend -s >= 2)."""
    return bool(load_library().umgap_route_sampled_applies(index._h, C.byref(opts)))


def route_pack_sampled_dev(index: Index, opts: PipelineOpts, phase: int, nt_ptr: int, read_off_ptr: int, nreads: int,
                           total_nt: int, cap: int, send_h_ptr: int, send_pos_ptr: int, cursors_ptr: int,
                           frame_hits_ptr: int, ids_ptr: int, stream: int = 0, group_off_ptr: int = 0, g_lo: int = 0,
                           g_hi: int = 0, slot: int = 0) -> None:
    _check(load_library().umgap_route_pack_sampled_dev(
        index._h, C.byref(opts), C.c_int(phase), C.c_void_p(nt_ptr), C.c_void_p(read_off_ptr), C.c_uint64(nreads),
        C.c_uint64(total_nt), C.c_uint64(cap), C.c_void_p(send_h_ptr), C.c_void_p(send_pos_ptr), C.c_void_p(cursors_ptr),
        C.c_void_p(frame_hits_ptr), C.c_void_p(ids_ptr), C.c_void_p(group_off_ptr), C.c_uint64(g_lo), C.c_uint64(g_hi),
        C.c_int(slot), C.c_void_p(stream)))


def route_scatter_hits_dev(index: Index, ans_ptr: int, send_pos_ptr: int, cursors_ptr: int, cap: int,
                           frame_hits_ptr: int, stream: int = 0) -> None:
    _check(load_library().umgap_route_scatter_hits_dev(index._h, C.c_void_p(ans_ptr), C.c_void_p(send_pos_ptr),
                                                       C.c_void_p(cursors_ptr), C.c_uint64(cap),
                                                       C.c_void_p(frame_hits_ptr), C.c_void_p(stream)))


def classify_ids_masked_dev(index: Index, tax: Taxonomy, opts: PipelineOpts, ids_ptr: int, read_off_ptr: int,
                            total_nt: int, group_off_ptr: int, ngroups: int, frame_hits_ptr: int, frame_major: bool,
                            out_ptr: int, stream: int = 0) -> None:
    _check(load_library().umgap_classify_ids_masked_dev(
        index._h, tax._h, C.byref(opts), C.c_void_p(ids_ptr), C.c_void_p(read_off_ptr), C.c_uint64(total_nt),
        C.c_void_p(group_off_ptr), C.c_uint64(ngroups), C.c_void_p(frame_hits_ptr), C.c_int(1 if frame_major else 0),
        C.c_void_p(out_ptr), C.c_void_p(stream)))


def classify_ids_dev(index: Index, tax: Taxonomy, opts: PipelineOpts, ids_ptr: int, read_off_ptr: int, total_nt: int,
                     group_off_ptr: int, ngroups: int, out_ptr: int, stream: int = 0) -> None:
    _check(load_library().umgap_classify_ids_dev(index._h, tax._h, C.byref(opts), C.c_void_p(ids_ptr),
                                                 C.c_void_p(read_off_ptr), C.c_uint64(total_nt), C.c_void_p(group_off_ptr),
                                                 C.c_uint64(ngroups), C.c_void_p(out_ptr), C.c_void_p(stream)))
