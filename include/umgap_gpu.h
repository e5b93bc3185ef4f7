/*
 * umgap_gpu.h -- C ABI of libumgap_gpu.so, the B200 (sm_100a) implementation of UMGAP's
 * per-read classification hot path.
 *
 * The reference (unipept/umgap, Rust) has no FFI seam of its own; its boundary is the
 * process (argv + FASTA streams + an `fst` index file + a taxonomy TSV).  Each entry point
 * below replaces the body of one reference command loop, taking the *parsed* form of that
 * command's stdin and producing the parsed form of its stdout.  The reference interface each
 * one replaces is cited as file:line relative to the reference repository.  A Rust `-sys`
 * crate binds these symbols one to one.
 *
 * Conventions
 *   - Every function returns 0 on success and a negative umgap_status on failure; the
 *     message for the calling thread is available from umgap_last_error().
 *   - Pointers are caller-owned HOST memory unless the parameter name ends in `_dev`
 *     (device memory on the handle's GPU).  No torch / C++ types cross the boundary.
 *   - Offsets arrays have n+1 entries (CSR style): item i spans [off[i], off[i+1]).
 *   - Handles are opaque; one in-flight call per handle, several handles may coexist.
 *   - There is no CPU fallback: without a usable CUDA device every compute entry point
 *     fails with UMGAP_ERR_CUDA.
 */
#ifndef UMGAP_GPU_H
#define UMGAP_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UMGAP_ABI_VERSION 1

typedef enum umgap_status {
    UMGAP_OK = 0,
    UMGAP_ERR_INVALID = -1,       /* bad argument / invalid invocation (reference: exit 1)   */
    UMGAP_ERR_IO = -2,            /* file could not be read / malformed file                 */
    UMGAP_ERR_CUDA = -3,          /* CUDA runtime failure or no device                       */
    UMGAP_ERR_UNKNOWN_TAXON = -4, /* "Unknown Taxon ID: <id>" (tree/mod.rs:37, rmq/rtl.rs:47) */
    UMGAP_ERR_CAPACITY = -5,      /* value / key outside what the device table can encode    */
    UMGAP_ERR_NOMEM = -6
} umgap_status;

/* Result markers in uint32 outputs. */
#define UMGAP_MISS 0xFFFFFFFFu   /* k-mer absent from the index (before -o mapping)          */
#define UMGAP_ABSENT 0xFFFFFFFFu /* classify: group produced no record at all (every frame
                                    shorter than k, prot2kmer2lca.rs:172)                    */

typedef struct umgap_index umgap_index;       /* GPU-resident key -> taxon table            */
typedef struct umgap_taxonomy umgap_taxonomy; /* GPU-resident tree: ancestor matrix etc.    */

/* Aggregation strategies of `taxa2agg -a` (taxa2agg.rs:111-139, 186-221). */
enum { UMGAP_AGG_LCA_STAR = 0, UMGAP_AGG_HYBRID = 1, UMGAP_AGG_MRTL = 2,
       UMGAP_AGG_RMQ_HYBRID = 3 /* `-m rmq -a hybrid`, the LCA / MRTL mix of rmq/mix.rs:56-93: umgap_aggregate_scored only */ };

/* ---- errors / device ------------------------------------------------------------------ */
const char* umgap_last_error(void);
int umgap_abi_version(void);
int umgap_device_count(void); /* number of CUDA devices, or negative status */

/* Page-locked host memory for the buffers handed to the host-buffer entry points (the read buffer of a parser):
 * copies out of it run asynchronously at full PCIe rate, copies out of pageable memory are staged and serialise.
 * Needs a CUDA device; NULL on failure (umgap_last_error).                                                      */
void* umgap_host_alloc(size_t bytes);
void umgap_host_free(void* p);

/* ---- index: replaces fst::Map::{from_path,from_bytes} + Map::get ------------------------
 * call sites prot2kmer2lca.rs:109-114,176; prot2tryp2lca.rs:89-94,130.                      */

/* Streams an `fst` 0.3.x (format v2) Map file once and builds the device table.
 * k > 0: fixed-length k-mer table (keys of any other length in the file can never be
 * queried by prot2kmer2lca -k and are skipped; k <= 9).  k == 0: variable-length peptide
 * table for prot2tryp2lca.  load_factor in (0,1]; <= 0 selects the default: the sparsest of
 * 0.5 / 0.6 / 0.7 / 0.8 whose table fits in half of the free HBM (fewer second probes), else 0.85. */
int umgap_index_load_fst(const char* path, int k, int device, double load_factor,
                         umgap_index** out);

/* Same table from explicit pairs.  keys: concatenated key bytes, key i spans
 * [key_off[i], key_off[i+1]); key_off may be NULL when k > 0 (then key i = keys[i*k..]). */
int umgap_index_from_pairs(const uint8_t* keys, const uint64_t* key_off, const uint64_t* values,
                           uint64_t n, int k, int device, double load_factor, umgap_index** out);

void umgap_index_free(umgap_index* idx);

/* ---- key-range-sharded index (multi-GPU, one process per GPU; SURVEY 8(e) mode 2) -----------
 * Rank r builds shard r of nshards with one of the *_shard constructors (it keeps the keys whose
 * hash falls in its range), exports a descriptor holding CUDA IPC handles of its table levels,
 * the ranks exchange descriptors (e.g. torch.distributed all_gather) and every rank attaches all
 * of them: afterwards the lookup kernel reads remote shards directly from the owning GPU's HBM
 * over NVLink peer mappings -- there is no routing kernel and no collective on the data path.   */
typedef struct umgap_shard_desc {
    unsigned char ipc[4][64];        /* cudaIpcMemHandle_t of each level                        */
    uint32_t nlines[4];
    int nlevels, shard, nshards, device, alphabet_size;
    unsigned char code_of_byte[256]; /* residue alphabet (must agree across shards)             */
} umgap_shard_desc;
int umgap_index_load_fst_shard(const char* path, int k, int device, double load_factor, int shard,
                               int nshards, umgap_index** out);
int umgap_index_from_pairs_shard(const uint8_t* keys, const uint64_t* key_off, const uint64_t* values,
                                 uint64_t n, int k, int device, double load_factor, int shard,
                                 int nshards, umgap_index** out);
int umgap_index_shard_desc(const umgap_index* idx, umgap_shard_desc* desc);
int umgap_index_attach_shards(umgap_index* idx, const umgap_shard_desc* descs, int nshards);
/* Same for one process that drives several GPUs (or several shards on one GPU): the shards' own
 * handles, in shard order; peer access is enabled as needed.                                    */
int umgap_index_attach_shards_local(umgap_index* idx, umgap_index* const* shards, int nshards);

/* ---- buildindex / printindex: the index FILE itself (host only; buildindex.rs:32-48, printindex.rs:38-51).
 * umgap_fst_writer_*: fst::MapBuilder -- keys in strictly increasing byte order, each with its u64 value; writes an
 * `fst` format-v2 Map file to `path` (NULL or "-": stdout).  No suffix sharing: larger than the crate's files, same
 * content.  finish() and abort() release the writer.  umgap_fst_stream: fst::Map::stream -- calls fn for every key in
 * order (a non-zero return stops with UMGAP_ERR_IO); *n_keys receives the count stored in the file's footer.      */
typedef struct umgap_fst_writer umgap_fst_writer;
int umgap_fst_writer_open(const char* path, umgap_fst_writer** out);
int umgap_fst_writer_insert(umgap_fst_writer* w, const uint8_t* key, size_t len, uint64_t value);
int umgap_fst_writer_finish(umgap_fst_writer* w);
void umgap_fst_writer_abort(umgap_fst_writer* w);
typedef int (*umgap_fst_key_fn)(const uint8_t* key, size_t len, uint64_t value, void* user);
int umgap_fst_stream(const char* path, umgap_fst_key_fn fn, void* user, uint64_t* n_keys);

typedef struct umgap_index_info {
    uint64_t n_keys;        /* distinct keys resident                                      */
    uint64_t n_buckets;     /* 32-byte buckets                                             */
    uint64_t bytes;         /* device bytes of the table                                   */
    uint64_t n_skipped;     /* file keys skipped (length != k)                             */
    uint64_t n_flagged;     /* buckets whose overflow flag is set                          */
    uint64_t n_displaced;   /* keys stored outside their home bucket                       */
    uint64_t max_probe;     /* longest probe sequence of any resident key (buckets)        */
    int k;                  /* 0 = variable-length table                                   */
    int device;
    int alphabet_size;      /* distinct residue bytes seen in keys                         */
    double load_factor;     /* slots used / slots of the main level, as chosen at build     */
} umgap_index_info;
int umgap_index_get_info(const umgap_index* idx, umgap_index_info* info);
/* splitkmers | sort | joinkmers | buildindex on the device (splitkmers.rs:44-66, joinkmers.rs:53-105,
 * buildindex.rs:32-48): the k-mer table straight from a protein table.  Protein r = aa[prot_off[r] ..
 * prot_off[r+1]) with taxon id prot_taxon[r] (the two columns of the TSV `splitkmers` reads).  Every k-mer
 * maps to the hybrid (factor 0.95) aggregate of its proteins' taxa, each first replaced by its nearest
 * valid ancestor, the result snapped to a ranked taxon; taxa the taxonomy does not hold are dropped.
 * The table lives on the taxonomy's device.                                                        */
int umgap_index_build_from_proteins(const umgap_taxonomy* tax, const uint8_t* aa, const uint64_t* prot_off,
                                    const uint64_t* prot_taxon, uint64_t nprot, int k, double load_factor,
                                    umgap_index** out);

/* A level-0 table larger than `bytes` is probed one hash-prefix region of at most that size per
 * lookup-kernel pass (B200's random-access rate collapses beyond ~64 GiB of footprint; default
 * 60 GiB, 0 restores it).  Results do not depend on it.                                          */
int umgap_index_set_probe_region(umgap_index* idx, uint64_t bytes);

/* ---- taxonomy: replaces taxon::read_taxa_file + TaxonTree::new + TaxonList::new +
 * TaxonTree::snapping (taxon.rs:89-128, 135-163, 224-247, 251-301; taxa2agg.rs:103-109). */
int umgap_taxonomy_load(const char* tsv_path, int device, umgap_taxonomy** out);
/* rank[i]: index into the reference's Rank enum (rank.rs:9-44), 0 = "no rank". */
int umgap_taxonomy_from_arrays(const uint64_t* ids, const uint64_t* parents, const uint8_t* rank,
                               const uint8_t* valid, uint64_t n, int device,
                               umgap_taxonomy** out);
void umgap_taxonomy_free(umgap_taxonomy* tax);
typedef struct umgap_taxonomy_info {
    uint64_t n_taxa;
    uint64_t max_id;
    uint64_t root;
    uint32_t max_depth; /* depth of the deepest taxon (root = 0) */
    int device;
} umgap_taxonomy_info;
int umgap_taxonomy_get_info(const umgap_taxonomy* tax, umgap_taxonomy_info* info);

/* ---- translate: replaces the record loop of translate.rs:114-133 (+ dna/mod.rs:23-103,
 * dna/translation.rs:125-144).  frames_mask bit i selects frame i in the reference order
 * 1,2,3,1R,2R,3R (translate.rs:83-90).  Output: one peptide per (read, selected frame) in
 * that order; aa_off has nreads*popcount(frames_mask)+1 entries.  aa_out must hold
 * umgap_translate_bound() bytes.  Unknown table -> UMGAP_ERR_INVALID ("Unknown table").   */
uint64_t umgap_translate_bound(uint64_t total_nt, uint64_t nreads, uint8_t frames_mask);
int umgap_translate(int device, const uint8_t* nt, const uint64_t* read_off, uint64_t nreads,
                    int table, int methionine, uint8_t frames_mask, uint8_t* aa_out,
                    uint64_t* aa_off);

/* ---- prot2kmer2lca: replaces the per-record body of stream_prot2kmer2lca
 * (prot2kmer2lca.rs:168-185).  Peptides shorter than k produce NO record: kept[i] = 0 and an
 * empty span (the reference drops the header too, :172).  With one_on_one (-o) a miss is
 * written as 0, otherwise it is omitted (:115,176).  taxa_out must hold
 * umgap_kmer_lookup_bound() entries; taxa_off has npeps+1 entries.                          */
uint64_t umgap_kmer_lookup_bound(uint64_t total_aa, uint64_t npeps);
int umgap_kmer_lookup(const umgap_index* idx, const uint8_t* aa, const uint64_t* pep_off,
                      uint64_t npeps, int one_on_one, uint32_t* taxa_out, uint64_t* taxa_off,
                      uint8_t* kept);

/* ---- prot2tryp2lca: replaces prot2tryp2lca.rs:105-134 for the default cleavage pattern
 * ([KR])([^P]).  Each input item is one physical LINE of a record (the reader is
 * unwrap=false, :100,109); rec_of_line is not needed: the caller concatenates per record.
 * keep / drop: NUL-terminated residue sets ("" = none).                                    */
uint64_t umgap_tryp_lookup_bound(uint64_t total_aa, uint64_t nlines);
int umgap_tryp_lookup(const umgap_index* idx, const uint8_t* aa, const uint64_t* line_off,
                      uint64_t nlines, int minlen, int maxlen, const char* keep,
                      const char* drop, int one_on_one, uint32_t* taxa_out, uint64_t* taxa_off);

/* ---- seedextend (unranked): replaces seedextend.rs:92-149,167-176.  out holds at most
 * rec_off[nrecs] entries; out_off has nrecs+1 entries.                                     */
int umgap_seedextend(int device, const uint32_t* taxa, const uint64_t* rec_off, uint64_t nrecs,
                     int min_seed_size, int max_gap_size, uint32_t* out, uint64_t* out_off);
/* seedextend -r <taxon file> [-p penalty] (seedextend.rs:151-164): of a record's extended seeds only the one with
 * the highest summed rank score is kept (the last of equal maxima); an id scores TaxonList::score (taxon.rs:181-191,
 * rank.rs:86-99), `penalty` (reference default 5) when that is None.                                          */
int umgap_seedextend_ranked(const umgap_taxonomy* tax, const uint32_t* taxa, const uint64_t* rec_off, uint64_t nrecs,
                            int min_seed_size, int max_gap_size, int penalty, uint32_t* out, uint64_t* out_off);

/* ---- taxa2agg: replaces the record loop of taxa2agg.rs:159-181 with the aggregators
 * tree::lca (tree/lca.rs:34-40), tree::mix (tree/mix.rs:43-64) and rmq::rtl
 * (rmq/rtl.rs:39-57), unscored input.  taxon_out[i] is the snapped aggregate of record i
 * ("1" for an empty record, :174-175).  Ties in hybrid / MRTL, which the reference breaks
 * by hash iteration order, are broken deterministically.                    */
int umgap_aggregate(const umgap_taxonomy* tax, const uint32_t* taxa, const uint64_t* rec_off,
                    uint64_t nrecs, int strategy, float factor, float lower_bound,
                    int ranked_only, uint32_t* taxon_out);
/* taxa2agg -s (taxa2agg.rs:141-148, 162-170): every id carries an f32 score ("taxon=score"); a taxon's weight is the sum
 * of its scores in input order, the lower bound applies to the sums, and every later f32 addition follows the
 * reference's order (rmq/rtl.rs:43-46 walking up; tree/mod.rs:75-80 and 95-96 with a node's children in ascending
 * taxon id, one of the HashSet orders the reference can take).  An unknown taxon raises only when it survives the
 * lower bound.  One thread per record: the mode is used by no preset of scripts/umgap-analyse.sh.  Also takes
 * UMGAP_AGG_RMQ_HYBRID (taxa2agg -m rmq -a hybrid, rmq/mix.rs:56-93; unscored input = scores of 1.0).            */
int umgap_aggregate_scored(const umgap_taxonomy* tax, const uint32_t* taxa, const float* scores, const uint64_t* rec_off,
                           uint64_t nrecs, int strategy, float factor, float lower_bound, int ranked_only,
                           uint32_t* taxon_out);

/* ---- fused path: translate -a | prot2kmer2lca [-o] | [seedextend] | uniq -d | taxa2agg
 * (scripts/umgap-analyse.sh:276-311) without materialising text between the stages.       */
typedef struct umgap_pipeline_opts {
    int table;          /* translate -t, default 1                                        */
    int methionine;     /* translate -m                                                   */
    int one_on_one;     /* prot2kmer2lca -o                                               */
    int seedextend;     /* 0: stage absent (tryptic presets), 1: present                  */
    int min_seed_size;  /* seedextend -s, default 2                                       */
    int max_gap_size;   /* seedextend -g, default 0                                       */
    int strategy;       /* UMGAP_AGG_*                                                    */
    float factor;       /* taxa2agg -f, default 0.25                                      */
    float lower_bound;  /* taxa2agg -l, default 0                                         */
    int ranked_only;    /* taxa2agg -r                                                    */
} umgap_pipeline_opts;
void umgap_pipeline_opts_default(umgap_pipeline_opts* o);

/* Reads r in [group_off[g], group_off[g+1]) are the records `uniq -d` joins (uniq.rs:56-84):
 * for paired-end input two consecutive reads.  taxon_out has ngroups entries; a group none
 * of whose frames reaches k residues yields UMGAP_ABSENT (no record in the reference).
 * n_lookups (optional) receives the number of k-mer lookups performed.                      */
int umgap_classify_reads(const umgap_index* idx, const umgap_taxonomy* tax,
                         const umgap_pipeline_opts* opts, const uint8_t* nt,
                         const uint64_t* read_off, uint64_t nreads, const uint64_t* group_off,
                         uint64_t ngroups, uint32_t* taxon_out, uint64_t* n_lookups);

/* Device-resident variant: all arrays already in HBM on idx's device; runs asynchronously on
 * `stream` (a cudaStream_t, NULL = default stream) using the handle's cached workspace.
 * total_nt = read_off[nreads] must be supplied by the caller (no device->host sync inside). */
int umgap_classify_reads_dev(const umgap_index* idx, const umgap_taxonomy* tax,
                             const umgap_pipeline_opts* opts, const uint8_t* nt_dev,
                             const uint64_t* read_off_dev, uint64_t nreads, uint64_t total_nt,
                             const uint64_t* group_off_dev, uint64_t ngroups,
                             uint32_t* taxon_out_dev, void* stream);

/* Packed host form of the reads: 2 bits per nucleotide plus a short list of the words that hold an N -- a quarter
 * of the bytes umgap_classify_reads moves over PCIe, and a form a FASTA parser can emit directly.  Nucleotide x of
 * the concatenated reads sits at bits 2 (x % 16) of codes[x / 16] with A, C, G, T = 0, 1, 2, 3
 * (umgap_packed_words(total_nt) words).  Every 16-nucleotide word that holds a byte which is none of those four
 * letters (lower case included: N, dna/mod.rs:34-44) has one entry in n_entries: word index << 16 | flags, bit
 * x % 16 of flags up for each such nucleotide; the entries ascend by word index.  read_off stays in nucleotides.
 * umgap_pack_reads fills both from the bytes on `threads` host threads (<= 0: all); it sets *n_count and fails with
 * UMGAP_ERR_CAPACITY when n_entries (n_cap entries) is too small (umgap_packed_words(total_nt) always suffices).
 * Results are those of umgap_classify_reads on the same reads.                                                  */
uint64_t umgap_packed_words(uint64_t total_nt);
int umgap_pack_reads(const uint8_t* nt, uint64_t total_nt, uint32_t* codes, uint64_t* n_entries, uint64_t n_cap,
                     uint64_t* n_count, int threads);
int umgap_classify_reads_packed(const umgap_index* idx, const umgap_taxonomy* tax, const umgap_pipeline_opts* opts,
                                const uint32_t* codes, const uint64_t* n_entries, uint64_t n_count,
                                const uint64_t* read_off, uint64_t nreads, const uint64_t* group_off,
                                uint64_t ngroups, uint32_t* taxon_out, uint64_t* n_lookups);

/* Asynchronous forms of the two host-buffer calls: the batch is enqueued on the index's internal streams and the call
 * returns; umgap_pending_wait blocks until taxon_out holds the batch's results, raises what the batch raised
 * (Unknown Taxon ID) and releases the ticket.  A host that hands over batch i + 1 before it waits for batch i keeps
 * the GPU busy across the seam: the uploads and first kernels of one batch run in the tail of the other (one batch at
 * a time leaves ~0.5 ms per 1 M pairs idle).  All host arrays of a batch -- page-locked, else the copies are not
 * asynchronous -- stay valid and unchanged until its wait returns; batches complete in the order they were enqueued;
 * at most 32 in flight per index; calls on one index come from one thread at a time; every ticket is waited for
 * before its index is freed.  The reference's stages are synchronous filters; this is the shape of its pipe (one
 * stage reads while the next computes).                                                                          */
typedef struct umgap_pending umgap_pending;
int umgap_classify_reads_async(const umgap_index* idx, const umgap_taxonomy* tax, const umgap_pipeline_opts* opts,
                               const uint8_t* nt, const uint64_t* read_off, uint64_t nreads, const uint64_t* group_off,
                               uint64_t ngroups, uint32_t* taxon_out, umgap_pending** out);
int umgap_classify_reads_packed_async(const umgap_index* idx, const umgap_taxonomy* tax, const umgap_pipeline_opts* opts,
                                      const uint32_t* codes, const uint64_t* n_entries, uint64_t n_count,
                                      const uint64_t* read_off, uint64_t nreads, const uint64_t* group_off, uint64_t ngroups,
                                      uint32_t* taxon_out, umgap_pending** out);
int umgap_pending_wait(umgap_pending* p);

/* ---- multi-GPU, replicated index (SURVEY 8(e) mode 1: reads are independent units, index and taxonomy
 * replicated per GPU, no collective on the data path) inside one process -- what the reference gets from running
 * several pipelines side by side.  umgap_index_replicate / umgap_taxonomy_replicate copy a loaded table / tree to
 * another device (device to device, NVLink between peers) instead of streaming the file again.
 * umgap_classify_reads_multi cuts the groups of a batch into one contiguous, nucleotide-balanced range per replica
 * (never inside a uniq group, uniq.rs:56-84) and drives each replica's chunked host-buffer path from its own host
 * thread; results are those of umgap_classify_reads, in input order.  idx[i] and tax[i] must live on the same
 * device; a replica may not appear twice.                                                                      */
int umgap_index_replicate(const umgap_index* src, int device, umgap_index** out);
int umgap_taxonomy_replicate(const umgap_taxonomy* src, int device, umgap_taxonomy** out);
int umgap_classify_reads_multi(const umgap_index* const* idx, const umgap_taxonomy* const* tax, int ngpus,
                               const umgap_pipeline_opts* opts, const uint8_t* nt, const uint64_t* read_off,
                               uint64_t nreads, const uint64_t* group_off, uint64_t ngroups, uint32_t* taxon_out,
                               uint64_t* n_lookups);
int umgap_classify_reads_packed_multi(const umgap_index* const* idx, const umgap_taxonomy* const* tax, int ngpus,
                                      const umgap_pipeline_opts* opts, const uint32_t* codes, const uint64_t* n_entries,
                                      uint64_t n_count, const uint64_t* read_off, uint64_t nreads,
                                      const uint64_t* group_off, uint64_t ngroups, uint32_t* taxon_out,
                                      uint64_t* n_lookups);

/* Stage kernels on device-resident data, used by the benchmark to time the lookup kernel in
 * isolation: ids_dev receives 2*total_nt entries (position-major, both strands).           */
int umgap_translate_lookup_dev(const umgap_index* idx, const umgap_pipeline_opts* opts,
                               const uint8_t* nt_dev, const uint64_t* read_off_dev,
                               uint64_t nreads, uint64_t total_nt, uint32_t* ids_dev,
                               void* stream);

/* ---- fused peptide path: prot2tryp2lca | uniq -d / | taxa2agg (the tryptic presets of
 * scripts/umgap-analyse.sh:291-300, behind the gene predictor) without text between the stages.
 * Lines l in [group_off[g], group_off[g+1]) are the records `uniq` joins.  Every peptide line is digested
 * (prot2tryp2lca.rs:112-117), the peptides of minlen..maxlen bytes that pass the keep / drop sets are looked up
 * in the variable-length table `idx` (k = 0), and the taxa of a group are aggregated (zeros dropped,
 * taxa2agg.rs:169; a group without a hit yields 1, :174-175; a group without lines UMGAP_ABSENT).
 * The host-buffer call sends the batch in ranges of whole groups on rotating streams; the ranges overlap (copy in,
 * kernels, results out) when aa, the offset arrays and taxon_out are page-locked (umgap_host_alloc).   */
typedef struct umgap_tryp_opts {
    int minlen;         /* prot2tryp2lca -l, default 5                                     */
    int maxlen;         /* prot2tryp2lca -L, default 50                                    */
    const char* keep;   /* prot2tryp2lca -k (NULL or "" = none)                            */
    const char* drop;   /* prot2tryp2lca -d                                                */
    int strategy;       /* UMGAP_AGG_*                                                     */
    float factor;       /* taxa2agg -f                                                     */
    float lower_bound;  /* taxa2agg -l                                                     */
    int ranked_only;    /* taxa2agg -r                                                     */
} umgap_tryp_opts;
void umgap_tryp_opts_default(umgap_tryp_opts* o);
int umgap_classify_peptides(const umgap_index* idx, const umgap_taxonomy* tax, const umgap_tryp_opts* opts,
                            const uint8_t* aa, const uint64_t* line_off, uint64_t nlines,
                            const uint64_t* group_off, uint64_t ngroups, uint32_t* taxon_out);
/* Device-resident variant, asynchronous on `stream`; total_aa = line_off[nlines]; aa_dev 8-byte aligned. */
int umgap_classify_peptides_dev(const umgap_index* idx, const umgap_taxonomy* tax, const umgap_tryp_opts* opts,
                                const uint8_t* aa_dev, const uint64_t* line_off_dev, uint64_t nlines,
                                uint64_t total_aa, const uint64_t* group_off_dev, uint64_t ngroups,
                                uint32_t* taxon_out_dev, void* stream);

/* ---- routed variant of the sharded mode: the exchange step of SURVEY 8(e).  Per batch and rank:
 *   umgap_route_pack_dev     reads -> 45-bit k-mer hashes bucketed by owning shard: send_h_dev and
 *                            send_pos_dev hold nshards buckets of `cap` entries, cursors_dev
 *                            (2*nshards u64) the bucket fills and, after them, overflow flags;
 *                            k-mers that cannot be keys are answered UMGAP_MISS in ids_dev directly
 *   (host)                   all-to-all of the bucket fills and of the buckets (NCCL)
 *   umgap_lookup_hashes_dev  looks the received hashes up in this rank's shard
 *   (host)                   all-to-all of the answers
 *   umgap_route_scatter_dev  answers -> ids_dev (the layout umgap_translate_lookup_dev produces)
 *   umgap_classify_ids_dev   seedextend | uniq | taxa2agg over ids_dev
 * umgap_b200/sharded.py drives this with torch.distributed.                                        */
int umgap_route_pack_dev(const umgap_index* idx, const umgap_pipeline_opts* opts, const uint8_t* nt_dev,
                         const uint64_t* read_off_dev, uint64_t nreads, uint64_t total_nt, uint64_t cap,
                         uint64_t* send_h_dev, uint32_t* send_pos_dev, uint64_t* cursors_dev,
                         uint32_t* ids_dev, void* stream);
int umgap_lookup_hashes_dev(const umgap_index* idx, const uint64_t* h_dev, const uint64_t* counts_dev,
                            int nsrc, uint64_t cap, uint32_t* out_dev, void* stream);
int umgap_route_scatter_dev(const umgap_index* idx, const uint32_t* ans_dev, const uint32_t* send_pos_dev,
                            const uint64_t* cursors_dev, uint64_t cap, uint32_t* ids_dev, void* stream);
int umgap_classify_ids_dev(const umgap_index* idx, const umgap_taxonomy* tax, const umgap_pipeline_opts* opts,
                           const uint32_t* ids_dev, const uint64_t* read_off_dev, uint64_t total_nt,
                           const uint64_t* group_off_dev, uint64_t ngroups, uint32_t* taxon_out_dev,
                           void* stream);

/* Sampled form of the exchange step (behind `-o | seedextend -s S`, S >= 2, k = 9: umgap_route_sampled_applies
 * returns 1): what umgap_classify_reads_dev does against a local table, with the lookups routed.
 *   umgap_route_pack_sampled_dev(phase 1)  every min(S,4)-th position of every frame record -> buckets
 *                                          (send_pos = read * 8 + frame); clears frame_hits_dev
 *   (host) exchange, umgap_lookup_hashes_dev, exchange back
 *   umgap_route_scatter_hits_dev           non-zero answers -> frame masks (one byte per read)
 *   umgap_route_pack_sampled_dev(phase 2)  every position of the frames whose mask bit is up -> buckets
 *                                          (send_pos = index into ids_dev, frame-major layout)
 *   (host) exchange, umgap_lookup_hashes_dev, exchange back, umgap_route_scatter_dev
 *   umgap_classify_ids_masked_dev          seedextend | uniq | taxa2agg over the flagged frames
 * Both phases of a batch go through the same index handle, phase 1 first (it leaves the list of reads
 * longer than a warp batch for phase 2).  frame_hits_dev: nreads bytes rounded up to 4, 4-byte aligned.
 * group_off_dev != NULL packs only the reads of groups [g_lo, g_hi) -- a batch can be cut into group ranges
 * that run on different streams with their own buckets and `slot` (0..5: the work list of the range), all
 * against the same frame_hits_dev / ids_dev, which the caller then clears once per batch.               */
int umgap_route_sampled_applies(const umgap_index* idx, const umgap_pipeline_opts* opts);
int umgap_route_pack_sampled_dev(const umgap_index* idx, const umgap_pipeline_opts* opts, int phase,
                                 const uint8_t* nt_dev, const uint64_t* read_off_dev, uint64_t nreads,
                                 uint64_t total_nt, uint64_t cap, uint64_t* send_h_dev, uint32_t* send_pos_dev,
                                 uint64_t* cursors_dev, uint8_t* frame_hits_dev, uint32_t* ids_dev,
                                 const uint64_t* group_off_dev, uint64_t g_lo, uint64_t g_hi, int slot, void* stream);
int umgap_route_scatter_hits_dev(const umgap_index* idx, const uint32_t* ans_dev, const uint32_t* send_pos_dev,
                                 const uint64_t* cursors_dev, uint64_t cap, uint8_t* frame_hits_dev, void* stream);
int umgap_classify_ids_masked_dev(const umgap_index* idx, const umgap_taxonomy* tax, const umgap_pipeline_opts* opts,
                                  const uint32_t* ids_dev, const uint64_t* read_off_dev, uint64_t total_nt,
                                  const uint64_t* group_off_dev, uint64_t ngroups, const uint8_t* frame_hits_dev,
                                  int frame_major, uint32_t* taxon_out_dev, void* stream);

/* ---- the exchange step with the kernels doing the transfers (exchange.cu): no collective library, no host in the
 * loop.  Every rank owns a shard and an *exchange region* of umgap_exchange_region_bytes() in its HBM that every other
 * rank maps (peer access inside one process; cuMem / IPC / symmetric-memory mappings across processes -- the library
 * takes the mapped pointers, however they were made).  The pack kernels store each k-mer hash straight into the
 * owner's inbox over NVLink, epoch flags order the rounds, the owner's lookup kernel stores the answers straight into
 * the requester's answer box; one batch = two rounds behind `-o | seedextend -s S` (S >= 2), else one, then the
 * classify kernel -- all enqueued on one stream without a host synchronisation.  Every rank needs a GPU of its own
 * (the ranks' kernels wait on each other).  regions[o] = this rank's mapping of rank o's region (zero-initialised by
 * its owner before the first batch); every rank calls umgap_exchange_classify_dev for every batch (a rank without
 * reads passes nreads = ngroups = 0).  umgap_exchange_status (after the caller synchronised the stream) raises bucket
 * overflow (UMGAP_ERR_CAPACITY), a peer that never answered (UMGAP_ERR_CUDA) and Unknown Taxon ID.               */
typedef struct umgap_exchange umgap_exchange;
uint64_t umgap_exchange_bucket_cap(int nranks, uint64_t max_total_nt);
uint64_t umgap_exchange_region_bytes(int nranks, uint64_t max_total_nt);
int umgap_exchange_create(const umgap_index* shard, const umgap_taxonomy* tax, int rank, int nranks, uint64_t max_total_nt,
                          void* const* regions, umgap_exchange** out);
/* A rank that keeps several batches in flight makes one context per lane (0..5), each with regions of its own: the
 * lanes use separate workspace sets of the shard's handle.                                                         */
int umgap_exchange_create_lane(const umgap_index* shard, const umgap_taxonomy* tax, int rank, int nranks,
                               uint64_t max_total_nt, void* const* regions, int lane, umgap_exchange** out);
void umgap_exchange_free(umgap_exchange* ex);
int umgap_exchange_classify_dev(umgap_exchange* ex, const umgap_pipeline_opts* opts, const uint8_t* nt_dev,
                                const uint64_t* read_off_dev, uint64_t nreads, uint64_t total_nt,
                                const uint64_t* group_off_dev, uint64_t ngroups, uint32_t* taxon_out_dev, void* stream);
int umgap_exchange_status(umgap_exchange* ex, uint64_t* lookups_routed);

/* One process driving every shard (the CLI, a Rust host with one thread): shards[i] is shard i of n, each on its own
 * GPU; peer access is enabled and the regions allocated here.  umgap_classify_reads_sharded takes host buffers like
 * umgap_classify_reads (the groups cut into one range per GPU, as many passes as the per-GPU buffers of max_total_nt
 * nucleotides require); the _dev form takes one device-resident batch per shard and returns at once --
 * umgap_sharded_sync waits and raises what umgap_exchange_status raises.                                          */
typedef struct umgap_sharded umgap_sharded;
int umgap_sharded_create(const umgap_index* const* shards, const umgap_taxonomy* const* tax, int n, uint64_t max_total_nt,
                         umgap_sharded** out);
void umgap_sharded_free(umgap_sharded* s);
int umgap_classify_reads_sharded_dev(umgap_sharded* s, const umgap_pipeline_opts* opts, const uint8_t* const* nt_dev,
                                     const uint64_t* const* read_off_dev, const uint64_t* nreads, const uint64_t* total_nt,
                                     const uint64_t* const* group_off_dev, const uint64_t* ngroups,
                                     uint32_t* const* taxon_out_dev);
int umgap_sharded_sync(umgap_sharded* s, uint64_t* lookups_routed);
int umgap_classify_reads_sharded(umgap_sharded* s, const umgap_pipeline_opts* opts, const uint8_t* nt,
                                 const uint64_t* read_off, uint64_t nreads, const uint64_t* group_off, uint64_t ngroups,
                                 uint32_t* taxon_out, uint64_t* n_lookups);

/* ---- measurement aid: when enabled, every launch of the two hot-path kernels is bracketed by CUDA
 * events on the stream it is launched on.  umgap_kernel_times() waits for the recorded launches,
 * returns the summed durations (ms) and launch counts since the last call, and clears them.      */
int umgap_kernel_timing(int enable);
int umgap_kernel_times(double* lookup_ms, uint64_t* lookup_launches, double* classify_ms,
                       uint64_t* classify_launches);
/* The same with the brackets of the exchange step: kind 0 lookup, 1 classify, 2 pack (hashes stored into the owners'
 * inboxes), 3 the time the wait kernels spun for peers (exposed transfers and skew), 4 scatter; nkinds <= 8, summed
 * over every device this process drives.                                                                        */
int umgap_kernel_times_ex(double* ms, uint64_t* launches, int nkinds);
/* Number of kernels the fused path (umgap_classify_reads[_dev], umgap_translate_lookup_dev,
 * umgap_classify_ids_dev) has launched in this process; the lookup stage is up to three launches
 * (residue-code pre-pass, sampled lookup kernel, long-read pass).                                  */
int umgap_kernel_launch_count(uint64_t* launches);
/* Bytes umgap_classify_reads has copied host -> device and device -> host in this process (offset arrays
 * that are arithmetic progressions are regenerated on the device instead of being uploaded).        */
int umgap_transfer_bytes(uint64_t* h2d, uint64_t* d2h);
/* umgap_classify_reads_dev can cut a batch into `slices` group ranges whose lookup and classify kernels
 * alternate on two internal streams (the classify kernel of a slice beside the lookup kernel of the next);
 * 1 (default; environment UMGAP_SLICES) = one lookup and one classify launch on the caller's stream.  With
 * the lookup kernel handing out its units dynamically the slices no longer pay (profiles/README.md 2c); the
 * mechanism stays for workloads with a heavier classify stage.  Returns the previous setting; slices <= 0
 * only queries.                                                                                          */
int umgap_pipeline_slices(int slices);
/* In front of `seedextend -s S` (S >= 2, with -o) the fused path probes every min(S,4)-th k-mer position first
 * and the others only for frames with a hit -- the same results with half the memory traffic.  0 switches
 * this off (every position is probed, as without seedextend), 1 on (default; environment UMGAP_NO_SAMPLING
 * starts with it off), a negative value only queries.  Returns the previous setting.                     */
int umgap_pipeline_sampling(int enable);

/* ---- benchmark / test aids (synthetic data of SURVEY 8(d); not part of the reference) ---- */
typedef struct umgap_synth_spec {
    uint64_t seed;
    uint64_t n_proteins;   /* proteins of protein_len residues each                        */
    uint32_t protein_len;  /* index keys = all 9-mers of all proteins                      */
    uint32_t home_pct;     /* % of k-mers valued with the protein's own taxon              */
    uint32_t ancestor_pct; /* % valued with a random ancestor; rest: unrelated taxon       */
} umgap_synth_spec;
/* Builds the table on the device from the counter-based proteome (no host copy). */
int umgap_index_build_synthetic(const umgap_synth_spec* spec, const umgap_taxonomy* tax,
                                int device, double load_factor, umgap_index** out);
int umgap_index_build_synthetic_shard(const umgap_synth_spec* spec, const umgap_taxonomy* tax,
                                      int device, double load_factor, int shard, int nshards,
                                      umgap_index** out);
/* Fills nt_dev with npairs*2 reads of read_len nucleotides drawn from the same proteome
 * (hit_pct % of pairs) or uniformly at random.                                              */
int umgap_synth_reads_dev(const umgap_synth_spec* spec, uint64_t read_seed, uint64_t first_pair,
                          uint64_t npairs, uint32_t read_len, uint32_t hit_pct,
                          uint8_t* nt_dev, void* stream);
/* Random 32-byte-sector gather over the index' own table memory: the measured denominator
 * of the "random-sector roofline".  Returns sectors/second in *rate.                       */
int umgap_randsector_bench(const umgap_index* idx, uint64_t n_gathers, int iters, double* rate);

#ifdef __cplusplus
}
#endif
#endif /* UMGAP_GPU_H */
