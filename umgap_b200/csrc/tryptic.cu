// prot2tryp2lca on the device (prot2tryp2lca.rs:88-140): variable-length peptide table + tryptic
// digest / filter / lookup kernel.
//
// Table.  Tryptic peptides are 5..50 bytes of arbitrary content, so the key bytes themselves are
// kept (one byte pool in HBM) and every tag match is verified against them: exact, no false
// positives.  Slots are 16 bytes {tag:32, value:32, pool offset:56 | length:8}, two per 32-byte
// sector, open addressing with linear probing at load factor <= 0.5; the home slot is
// floor(hash * nslots / 2^64) and the tag the low 32 bits of the 64-bit hash.
//
// Digest.  The reference applies the regex ([KR])([^P]) -> "$1\n$2" twice, turns '*' into a line
// break and drops empty pieces (:112-117).  In closed form (oracle.lookup.tryptic_digest, checked
// against the regex form): a peptide starts at i iff c[i] != '*' and (i == 0 or c[i-1] == '*' or
// (c[i-1] in KR and c[i] != 'P')); it ends before the next start or '*'.  One thread per start
// position walks its peptide once, hashing as it goes.
#include <algorithm>

#include "index.h"

namespace umgap {

struct VarSlot {
    uint32_t tag;
    uint32_t value;
    uint64_t off_len;  // pool offset << 8 | length; ~0 = empty
};
constexpr uint64_t kVarEmpty = ~0ull;

struct VarTable {
    VarSlot* slots = nullptr;
    uint8_t* pool = nullptr;
    uint64_t nslots = 0, pool_bytes = 0;
};

__host__ __device__ __forceinline__ uint64_t pep_hash_step(uint64_t h, uint8_t c) {
    h = (h ^ c) * 0x100000001B3ull;  // FNV-1a step
    return h;
}
__host__ __device__ __forceinline__ uint64_t pep_hash_finish(uint64_t h, uint32_t len) {
    h ^= len;
    h ^= h >> 32;
    h *= 0xD6E8FEB86659FD93ull;
    h ^= h >> 32;
    h *= 0xD6E8FEB86659FD93ull;
    h ^= h >> 32;
    return h;
}
constexpr uint64_t kFnvBasis = 0xCBF29CE484222325ull;

__global__ void var_fill_kernel(VarSlot* s, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        s[i].tag = 0;
        s[i].value = kNoValue;
        s[i].off_len = kVarEmpty;
    }
}

__global__ void var_insert_kernel(VarSlot* slots, uint64_t nslots, const uint8_t* __restrict__ pool,
                                  const uint64_t* __restrict__ key_off, const uint32_t* __restrict__ vals,
                                  uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t off = key_off[i];
        const uint32_t len = (uint32_t)(key_off[i + 1] - off);
        uint64_t h = kFnvBasis;
        for (uint32_t j = 0; j < len; ++j) h = pep_hash_step(h, pool[off + j]);
        h = pep_hash_finish(h, len);
        uint64_t s = __umul64hi(h, nslots);
        const unsigned long long mine = (off << 8) | len;
        for (;;) {
            unsigned long long* p = reinterpret_cast<unsigned long long*>(&slots[s].off_len);
            if (*reinterpret_cast<volatile unsigned long long*>(p) == kVarEmpty &&
                atomicCAS(p, kVarEmpty, mine) == kVarEmpty) {
                slots[s].tag = (uint32_t)h;
                slots[s].value = vals[i];
                break;
            }
            s = (s + 1 == nslots) ? 0 : s + 1;
        }
    }
}

struct TrypParams {
    uint32_t minlen, maxlen;
    uint32_t keep_all;       // bitmask with one bit per keep residue
    int filter_sets;         // keep or drop set non-empty (prot2tryp2lca.rs:122)
};

constexpr uint32_t kTrypNone = 0xFFFFFFFEu;  // position that starts no (kept) peptide

// What starts at byte i of the line [b, e): kTrypNone (no kept peptide starts here), kNoValue (a kept peptide
// the table does not hold) or the peptide's value.
__device__ __forceinline__ uint32_t tryp_at(const VarSlot* __restrict__ slots, uint64_t nslots, const uint8_t* __restrict__ pool,
                                            const uint8_t* __restrict__ aa, uint64_t i, uint64_t b, uint64_t e,
                                            const TrypParams& tp, const uint8_t* __restrict__ set_lut) {
    const uint8_t c = aa[i];
    bool start = c != '*';
    if (start && i > b) {
        const uint8_t p = aa[i - 1];
        start = p == '*' || ((p == 'K' || p == 'R') && c != 'P');
    }
    if (!start) return kTrypNone;
    uint64_t h = kFnvBasis;
    uint32_t len = 0, keep_seen = 0;
    bool dropped = false, too_long = false;
    for (uint64_t j = i; j < e; ++j) {
        const uint8_t x = aa[j];
        if (x == '*') break;
        if (len == tp.maxlen) {  // one more residue would exceed -L: filtered out
            too_long = true;
            break;
        }
        h = pep_hash_step(h, x);
        ++len;
        const uint8_t s = set_lut[x];
        dropped |= (s & 0x80) != 0;
        if (s & 0x3F) keep_seen |= 1u << ((s & 0x3F) - 1);
        if ((x == 'K' || x == 'R') && j + 1 < e && aa[j + 1] != 'P') break;
    }
    bool keep = !too_long && len >= tp.minlen;
    if (keep && tp.filter_sets) keep = !dropped && keep_seen == tp.keep_all;
    if (!keep) return kTrypNone;
    h = pep_hash_finish(h, len);
    uint64_t s = __umul64hi(h, nslots);
    for (;;) {
        const VarSlot sl = slots[s];
        if (sl.off_len == kVarEmpty) return kNoValue;
        if (sl.tag == (uint32_t)h && (uint32_t)(sl.off_len & 0xFF) == len) {
            const uint8_t* k = pool + (sl.off_len >> 8);
            bool same = true;
            for (uint32_t q = 0; q < len; ++q) same &= k[q] == aa[i + q];
            if (same) return sl.value;
        }
        s = (s + 1 == nslots) ? 0 : s + 1;
    }
}

// out[i] for every input byte i: kTrypNone, kNoValue (kept peptide, miss) or the value.
__global__ void __launch_bounds__(256)
tryp_lookup_kernel(const VarSlot* __restrict__ slots, uint64_t nslots, const uint8_t* __restrict__ pool,
                   const uint8_t* __restrict__ aa, const uint64_t* __restrict__ line_off, uint64_t nlines,
                   const uint32_t* __restrict__ line_of_byte, uint64_t total, TrypParams tp,
                   const uint8_t* __restrict__ set_lut /* [256]: bit7 = drop, low 6 bits = keep index+1 */,
                   uint32_t* __restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const uint64_t line = line_of_byte[i];
        out[i] = tryp_at(slots, nslots, pool, aa, i, line_off[line], line_off[line + 1], tp, set_lut);
    }
}

// Fused path (umgap_classify_peptides): ONE THREAD PER LINE.  The lane walks its line once, eight bytes per load,
// carrying the digest state (prot2tryp2lca.rs:112-117 in the closed form above: a peptide ends before a '*', before a
// residue that follows K/R and is not P, and at the end of the line), the running hash, the length and the keep / drop
// flags; a finished peptide that passes the filters is parked in a four-entry list in shared memory, and the lookups
// run from the list afterwards (or when it is full), so that the lanes of a warp probe together instead of one
// lane at a time in the middle of the walk.  out[i] = the taxon of the kept peptide starting at byte i; the caller
// zeroed out[] (taxa2agg drops zeros, taxa2agg.rs:169), so `out` with the groups' byte ranges as records is what
// the aggregation kernel reads.  (A warp per line with the lanes striding over its bytes was the first form: 2.95 ms
// per 1 M pairs of 50-residue lines, instruction-bound -- five of 32 lanes walk a peptide at a time.)
constexpr int kPepList = 4;
constexpr int kPepThreads = 128;

__device__ __forceinline__ void tryp_probe(const VarSlot* __restrict__ slots, uint64_t nslots, const uint8_t* __restrict__ pool,
                                           const uint8_t* __restrict__ aa, uint64_t h, uint64_t start, uint32_t len,
                                           uint32_t* __restrict__ out, uint32_t shift) {
    uint64_t s = __umul64hi(h, nslots);
    for (;;) {
        const VarSlot sl = slots[s];
        if (sl.off_len == kVarEmpty) return;
        if (sl.tag == (uint32_t)h && (uint32_t)(sl.off_len & 0xFF) == len) {
            const uint8_t* k = pool + (sl.off_len >> 8);
            bool same = true;
            for (uint32_t q = 0; q < len; ++q) same &= k[q] == aa[start + q];
            if (same) {
                out[start >> shift] = sl.value;
                return;
            }
        }
        s = (s + 1 == nslots) ? 0 : s + 1;
    }
}

__global__ void __launch_bounds__(kPepThreads)
tryp_lookup_lines_kernel(const VarSlot* __restrict__ slots, uint64_t nslots, const uint8_t* __restrict__ pool,
                         const uint8_t* __restrict__ aa, uint64_t total_aa, const uint64_t* __restrict__ line_off, uint64_t nlines,
                         TrypParams tp, const uint8_t* __restrict__ set_lut, uint32_t* __restrict__ out, uint32_t shift) {
    __shared__ uint8_t s_lut[256];
    __shared__ uint64_t s_hash[kPepList][kPepThreads];
    __shared__ uint64_t s_where[kPepList][kPepThreads];  // start << 8 | length
    for (int i = threadIdx.x; i < 256; i += kPepThreads) s_lut[i] = set_lut[i];
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * kPepThreads;
    const int t = threadIdx.x;
    for (uint64_t l = (uint64_t)blockIdx.x * kPepThreads + t; l < nlines; l += stride) {
        const uint64_t b = line_off[l], e = line_off[l + 1];
        uint64_t h = kFnvBasis, pstart = b, word = 0;
        uint32_t len = 0, keep_seen = 0, npend = 0;
        bool dropped = false, too_long = false, prev_kr = false;
        auto flush = [&]() {
            for (uint32_t p = 0; p < npend; ++p)
                tryp_probe(slots, nslots, pool, aa, s_hash[p][t], s_where[p][t] >> 8, (uint32_t)(s_where[p][t] & 0xFF), out, shift);
            npend = 0;
        };
        auto finish = [&]() {  // the peptide [pstart, pstart + len) is complete
            bool keep = !too_long && len >= tp.minlen && len > 0;
            if (keep && tp.filter_sets) keep = !dropped && keep_seen == tp.keep_all;
            if (keep) {
                if (npend == (uint32_t)kPepList) flush();
                s_hash[npend][t] = pep_hash_finish(h, len);
                s_where[npend][t] = (pstart << 8) | len;
                ++npend;
            }
            h = kFnvBasis;
            len = 0;
            keep_seen = 0;
            dropped = too_long = false;
        };
        for (uint64_t i = b; i < e; ++i) {
            if (i == b || (i & 7u) == 0) {  // next aligned word (bytes past the end of the buffer are never loaded)
                const uint64_t w0 = i & ~7ull;
                if (w0 + 8 <= total_aa) {
                    word = __ldg(reinterpret_cast<const unsigned long long*>(aa + w0));
                } else {
                    word = 0;
                    for (uint64_t q = w0; q < total_aa; ++q) word |= (uint64_t)aa[q] << (8 * (q - w0));
                }
            }
            const uint8_t x = (uint8_t)(word >> (8 * (i & 7u)));
            if (x == '*') {
                finish();
                prev_kr = false;
                pstart = i + 1;
                continue;
            }
            if (prev_kr && x != 'P' && len > 0) {
                finish();
                pstart = i;
            }
            if (len == 0) pstart = i;
            if (len == tp.maxlen) {
                too_long = true;  // one more residue would exceed -L: the peptide is filtered out, its tail skipped
            } else if (!too_long) {
                h = pep_hash_step(h, x);
                ++len;
                const uint8_t sl = s_lut[x];
                dropped |= (sl & 0x80) != 0;
                if (sl & 0x3F) keep_seen |= 1u << ((sl & 0x3F) - 1);
            }
            prev_kr = x == 'K' || x == 'R';
        }
        finish();
        flush();
    }
}

// rec_off[g] = line_off[group_off[g]] >> shift: the range of `out` that holds the kept peptides of the lines `uniq` joins
// into group g.  A kept peptide has at least minlen >= 2^shift residues and peptides do not overlap, so start >> shift is a
// slot of its own, and a kept peptide of group g starts at least 2^shift bytes before the group's end: its slot lies below
// the first slot of group g + 1.  (shift 2 for -l >= 4 -- the presets use -l 9: a quarter of the zeroed array to clear and
// to aggregate over.)
__global__ void group_bytes_kernel(const uint64_t* __restrict__ line_off, const uint64_t* __restrict__ group_off, uint64_t ngroups,
                                   uint64_t* __restrict__ rec_off, uint32_t shift) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g <= ngroups; g += stride)
        rec_off[g] = line_off[group_off[g]] >> shift;
}

__global__ void line_of_byte_kernel(const uint64_t* __restrict__ line_off, uint64_t nlines, uint32_t* __restrict__ lob) {
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (uint64_t l = warp; l < nlines; l += nwarps)
        for (uint64_t i = line_off[l] + lane; i < line_off[l + 1]; i += 32) lob[i] = (uint32_t)l;
}

namespace {
struct VarSink : FstSink {
    std::vector<uint8_t> pool;
    std::vector<uint64_t> off{0};
    std::vector<uint32_t> vals;
    uint64_t skipped = 0;
    void on_key(const uint8_t* key, size_t len, uint64_t value) override {
        if (len == 0 || len > 255) {
            ++skipped;
            return;
        }
        if (value >= 0xFFFFFFFFull)
            UMGAP_FAIL(UMGAP_ERR_CAPACITY, "index value %llu does not fit 32 bits", (unsigned long long)value);
        pool.insert(pool.end(), key, key + len);
        off.push_back(pool.size());
        vals.push_back((uint32_t)value);
    }
};
}  // namespace

static void build_var_table(umgap_index* idx, const std::vector<uint8_t>& pool, const std::vector<uint64_t>& off,
                            const std::vector<uint32_t>& vals, double load_factor) {
    use_device(idx->device);
    if (load_factor <= 0 || load_factor > 0.9) load_factor = 0.5;
    const uint64_t n = vals.size();
    VarTable* t = new VarTable();
    idx->var_table = t;
    t->nslots = std::max<uint64_t>(1024, (uint64_t)((double)n / load_factor) + 1);
    t->pool_bytes = pool.size();
    UMGAP_CUDA(cudaMalloc((void**)&t->slots, t->nslots * sizeof(VarSlot)));
    UMGAP_CUDA(cudaMalloc((void**)&t->pool, std::max<size_t>(pool.size(), 16)));
    var_fill_kernel<<<148 * 8, 256>>>(t->slots, t->nslots);
    UMGAP_CUDA(cudaGetLastError());
    if (n) {
        UMGAP_CUDA(cudaMemcpy(t->pool, pool.data(), pool.size(), cudaMemcpyHostToDevice));
        DevBuf<uint64_t> d_off(n + 1);
        DevBuf<uint32_t> d_vals(n);
        UMGAP_CUDA(cudaMemcpy(d_off.p, off.data(), (n + 1) * 8, cudaMemcpyHostToDevice));
        UMGAP_CUDA(cudaMemcpy(d_vals.p, vals.data(), n * 4, cudaMemcpyHostToDevice));
        var_insert_kernel<<<(unsigned)std::min<uint64_t>(ceil_div(n, 256), 148 * 32), 256>>>(t->slots, t->nslots, t->pool,
                                                                                           d_off.p, d_vals.p, n);
        UMGAP_CUDA(cudaGetLastError());
        UMGAP_CUDA(cudaDeviceSynchronize());
    }
    idx->n_keys = n;
    idx->bytes = t->nslots * sizeof(VarSlot) + pool.size();
}

int build_var_table_from_fst(const char* path, umgap_index* idx, double load_factor) {
    return guarded([&] {
        VarSink sink;
        fst_stream_file(path, sink, nullptr);
        idx->n_skipped = sink.skipped;
        build_var_table(idx, sink.pool, sink.off, sink.vals, load_factor);
    });
}

int build_var_table_from_pairs(umgap_index* idx, const uint8_t* keys, const uint64_t* key_off, const uint64_t* values,
                               uint64_t n, double load_factor) {
    return guarded([&] {
        VarSink sink;
        for (uint64_t i = 0; i < n; ++i) sink.on_key(keys + key_off[i], (size_t)(key_off[i + 1] - key_off[i]), values[i]);
        idx->n_skipped = sink.skipped;
        build_var_table(idx, sink.pool, sink.off, sink.vals, load_factor);
    });
}

void free_var_table(void* p) {
    VarTable* t = (VarTable*)p;
    if (!t) return;
    if (t->slots) cudaFree(t->slots);
    if (t->pool) cudaFree(t->pool);
    delete t;
}

}  // namespace umgap

using namespace umgap;

static TrypParams make_tryp_params(int minlen, int maxlen, const char* keep, const char* drop, uint8_t (&lut)[256]) {
    if (minlen < 0 || maxlen < 0) UMGAP_FAIL(UMGAP_ERR_INVALID, "negative length bound");
    TrypParams tp{};
    tp.minlen = (uint32_t)minlen;
    tp.maxlen = (uint32_t)maxlen;
    int nkeep = 0;
    for (const char* p = keep ? keep : ""; *p; ++p) {
        uint8_t& e = lut[(uint8_t)*p];
        if (e & 0x3F) continue;
        if (nkeep == 32) UMGAP_FAIL(UMGAP_ERR_INVALID, "more than 32 distinct residues in --keep");
        e |= (uint8_t)(++nkeep);
    }
    for (const char* p = drop ? drop : ""; *p; ++p) lut[(uint8_t)*p] |= 0x80;
    tp.keep_all = nkeep == 32 ? 0xFFFFFFFFu : ((1u << nkeep) - 1);
    tp.filter_sets = (keep && *keep) || (drop && *drop);
    return tp;
}

namespace umgap {
// stages.cu: the aggregation kernel over device-resident records (zeros dropped, empty record -> 1)
void launch_aggregate(const umgap_taxonomy* tax, int strategy, float factor, float lower_bound, int ranked_only,
                      const uint32_t* taxa_dev, const uint64_t* rec_off_dev, uint64_t nrecs, uint32_t* scratch_dev,
                      uint32_t* out_dev, unsigned int* err_dev, cudaStream_t st);
}

// workspace slots of a peptide-table handle (a k = 0 index never runs the k-mer pipeline, whose slots these overlap)
enum { TWS_LUT = 0, TWS_OUT = 1, TWS_REC = 2, TWS_SCRATCH = 3, TWS_ERR = 4, TWS_AA = 5, TWS_LOFF = 6, TWS_GOFF = 7, TWS_RES = 8 };

// What the launches of a range of groups need: set up once per call by peptides_prepare.
struct PepRun {
    const VarTable* t;
    TrypParams tp;
    uint32_t shift;  // one slot of `out` per 2^shift residues (group_bytes_kernel)
    uint8_t* d_lut;
    uint32_t* d_out;
    uint64_t* d_rec;
    uint32_t* d_scratch;
    unsigned int* d_err;
};

// Checks, workspace, the cleared error slot and the keep / drop table, all on `st`.
static PepRun peptides_prepare(const umgap_index* idx, const umgap_taxonomy* tax, const umgap_tryp_opts* o, uint64_t total_aa,
                               uint64_t ngroups, cudaStream_t st) {
    if (!idx || !tax || !o) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
    if (idx->k != 0 || !idx->var_table) UMGAP_FAIL(UMGAP_ERR_INVALID, "index is not a variable-length peptide table");
    if (tax->device != idx->device) UMGAP_FAIL(UMGAP_ERR_INVALID, "index and taxonomy live on different devices");
    if (o->strategy < UMGAP_AGG_LCA_STAR || o->strategy > UMGAP_AGG_MRTL) UMGAP_FAIL(UMGAP_ERR_INVALID, "unknown aggregation strategy %d", o->strategy);
    uint8_t lut[256] = {};
    PepRun r{};
    r.tp = make_tryp_params(o->minlen, o->maxlen, o->keep, o->drop, lut);
    use_device(idx->device);
    r.t = (const VarTable*)idx->var_table;
    r.d_lut = (uint8_t*)idx->ws.get(TWS_LUT, 256);
    r.d_out = (uint32_t*)idx->ws.get(TWS_OUT, (total_aa + 1) * sizeof(uint32_t));
    r.d_rec = (uint64_t*)idx->ws.get(TWS_REC, (ngroups + 1) * sizeof(uint64_t));
    r.d_scratch = (uint32_t*)idx->ws.get(TWS_SCRATCH, (3 * total_aa + ngroups + 8) * sizeof(uint32_t));
    r.d_err = (unsigned int*)idx->ws.get(TWS_ERR, 2 * sizeof(unsigned int));
    UMGAP_CUDA(cudaMemsetAsync(r.d_err, 0, 2 * sizeof(unsigned int), st));
    UMGAP_CUDA(cudaMemcpyAsync(r.d_lut, lut, 256, cudaMemcpyHostToDevice, st));  // pageable source: staged before the call returns
    const uint32_t minlen = std::max<uint32_t>(1, r.tp.minlen);
    r.shift = minlen >= 4 ? 2 : minlen >= 2 ? 1 : 0;
    return r;
}

// Groups [g0, g1) = lines [l0, l1) = bytes [b0, b1) of the batch, on `st`.  Offsets and slots are those of the whole
// batch (aa_dev, line_off_dev, group_off_dev and the workspace are indexed absolutely), so ranges on different streams
// touch disjoint parts: a kept peptide of the range starts in [b0, b1 - 2^shift], its slot lies in
// [b0 >> shift, b1 >> shift) (group_bytes_kernel), and the scratch of record g is A = scratch + 3 * rec_off[g] + g.
// The walk never loads a byte at or beyond b1 (`total_aa` of the kernel), so a range can run while the next one is
// still on its way to the device.
static void peptides_range(const PepRun& r, const umgap_taxonomy* tax, const umgap_tryp_opts* o, const uint8_t* aa_dev,
                           const uint64_t* line_off_dev, const uint64_t* group_off_dev, uint64_t g0, uint64_t g1, uint64_t l0,
                           uint64_t l1, uint64_t b0, uint64_t b1, bool last, uint32_t* taxon_out_dev, cudaStream_t st) {
    if (g1 <= g0) return;
    const uint64_t s0 = b0 >> r.shift, s1 = (b1 >> r.shift) + (last ? 1 : 0);
    if (s1 > s0) UMGAP_CUDA(cudaMemsetAsync(r.d_out + s0, 0, (s1 - s0) * sizeof(uint32_t), st));
    if (l1 > l0 && b1 > b0) {
        tryp_lookup_lines_kernel<<<(unsigned)std::min<uint64_t>(ceil_div(l1 - l0, kPepThreads), 148 * 16), kPepThreads, 0, st>>>(
            r.t->slots, r.t->nslots, r.t->pool, aa_dev, b1, line_off_dev + l0, l1 - l0, r.tp, r.d_lut, r.d_out, r.shift);
        UMGAP_CUDA(cudaGetLastError());
    }
    group_bytes_kernel<<<(unsigned)std::min<uint64_t>(ceil_div(g1 - g0 + 1, 256), 148 * 8), 256, 0, st>>>(line_off_dev, group_off_dev + g0,
                                                                                                        g1 - g0, r.d_rec + g0, r.shift);
    UMGAP_CUDA(cudaGetLastError());
    launch_aggregate(tax, o->strategy, o->factor, o->lower_bound, o->ranked_only, r.d_out, r.d_rec + g0, g1 - g0, r.d_scratch + g0,
                     taxon_out_dev + g0, r.d_err, st);
}

static void classify_peptides_dev(const umgap_index* idx, const umgap_taxonomy* tax, const umgap_tryp_opts* o,
                                  const uint8_t* aa_dev, const uint64_t* line_off_dev, uint64_t nlines, uint64_t total_aa,
                                  const uint64_t* group_off_dev, uint64_t ngroups, uint32_t* taxon_out_dev, cudaStream_t st) {
    const PepRun r = peptides_prepare(idx, tax, o, total_aa, ngroups, st);
    if (!ngroups) return;
    if (nlines && total_aa && ((uintptr_t)aa_dev & 7u) != 0) UMGAP_FAIL(UMGAP_ERR_INVALID, "aa_dev must be 8-byte aligned");
    peptides_range(r, tax, o, aa_dev, line_off_dev, group_off_dev, 0, ngroups, 0, nlines, 0, total_aa, true, taxon_out_dev, st);
}

extern "C" {

void umgap_tryp_opts_default(umgap_tryp_opts* o) {
    if (!o) return;
    o->minlen = 5;                 // prot2tryp2lca -l / -L defaults (prot2tryp2lca.rs:61-67)
    o->maxlen = 50;
    o->keep = nullptr;
    o->drop = nullptr;
    o->strategy = UMGAP_AGG_HYBRID;  // taxa2agg defaults (taxa2agg.rs:111-125)
    o->factor = 0.25f;
    o->lower_bound = 0.0f;
    o->ranked_only = 0;
}

int umgap_classify_peptides_dev(const umgap_index* idx, const umgap_taxonomy* tax, const umgap_tryp_opts* opts,
                                const uint8_t* aa_dev, const uint64_t* line_off_dev, uint64_t nlines, uint64_t total_aa,
                                const uint64_t* group_off_dev, uint64_t ngroups, uint32_t* taxon_out_dev, void* stream) {
    return guarded([&] {
        classify_peptides_dev(idx, tax, opts, aa_dev, line_off_dev, nlines, total_aa, group_off_dev, ngroups, taxon_out_dev,
                              (cudaStream_t)stream);
    });
}

int umgap_classify_peptides(const umgap_index* idx, const umgap_taxonomy* tax, const umgap_tryp_opts* opts, const uint8_t* aa,
                            const uint64_t* line_off, uint64_t nlines, const uint64_t* group_off, uint64_t ngroups,
                            uint32_t* taxon_out) {
    return guarded([&] {
        if (!idx || !line_off || !group_off || (ngroups && !taxon_out)) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        if (!ngroups) return;
        if (group_off[ngroups] > nlines) UMGAP_FAIL(UMGAP_ERR_INVALID, "group_off exceeds the number of lines");
        const uint64_t total = nlines ? line_off[nlines] : 0;
        if (total && !aa) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        use_device(idx->device);
        uint8_t* d_aa = (uint8_t*)idx->ws.get(TWS_AA, total + 16);
        uint64_t* d_loff = (uint64_t*)idx->ws.get(TWS_LOFF, (nlines + 1) * sizeof(uint64_t));
        uint64_t* d_goff = (uint64_t*)idx->ws.get(TWS_GOFF, (ngroups + 1) * sizeof(uint64_t));
        uint32_t* d_res = (uint32_t*)idx->ws.get(TWS_RES, ngroups * sizeof(uint32_t));
        // The batch goes over in ranges of whole groups of about 16 MB of residues (UMGAP_PEP_CHUNK_BYTES; the first ones
        // 1/8, 1/4, 1/2 of that), rotating over the handle's chunk streams: each range's bytes and offsets are copied
        // to their places in the batch-sized device arrays, its kernels run behind its own copy and beside the next
        // range's, its results go back on the same stream.  The ranges overlap only when `taxon_out` (and the
        // inputs) are page-locked (umgap_host_alloc): a copy into pageable memory holds the host until the range is done.
        // (One copy of everything, then the kernels, then the results: 4.8 ms per 1 M pairs of 50-residue lines, of
        // which PCIe needs 2.3.)
        const char* e = getenv("UMGAP_PEP_CHUNK_BYTES");  // read per call, so that tests can move the seams
        const uint64_t chunk = std::max<uint64_t>(64, e && atoll(e) > 0 ? (uint64_t)atoll(e) : 8ull << 20);
        const int kStreams = 4;
        if (!idx->chunk_stream[0])
            for (int i = 0; i < 6; ++i) {
                UMGAP_CUDA(cudaStreamCreateWithFlags(&idx->chunk_stream[i], cudaStreamNonBlocking));
                UMGAP_CUDA(cudaEventCreateWithFlags(&idx->chunk_done[i], cudaEventDisableTiming));
            }
        cudaStream_t* st = idx->chunk_stream;
        const PepRun r = peptides_prepare(idx, tax, opts, total, ngroups, st[0]);
        UMGAP_CUDA(cudaEventRecord(idx->chunk_done[0], st[0]));
        for (int i = 1; i < kStreams; ++i) UMGAP_CUDA(cudaStreamWaitEvent(st[i], idx->chunk_done[0], 0));
        auto bytes_before = [&](uint64_t g) {
            if (group_off[g] > nlines) UMGAP_FAIL(UMGAP_ERR_INVALID, "group_off exceeds the number of lines");
            return line_off[group_off[g]];
        };
        std::vector<uint64_t> empty;  // groups without lines, found while the device works on the range
        try {
            uint64_t g0 = 0;
            for (int c = 0; g0 < ngroups; ++c) {
                const uint64_t b0 = bytes_before(g0), limit = c < 3 ? chunk >> (3 - c) : chunk;
                uint64_t lo = g0 + 1, hi = ngroups;  // largest g1 with at most `limit` bytes in [g0, g1), at least g0 + 1
                while (lo < hi) {
                    const uint64_t mid = lo + (hi - lo + 1) / 2;
                    if (bytes_before(mid) - b0 <= limit) lo = mid; else hi = mid - 1;
                }
                const uint64_t g1 = lo, l0 = group_off[g0], l1 = group_off[g1], b1 = line_off[l1];
                if (l1 < l0 || b1 < b0) UMGAP_FAIL(UMGAP_ERR_INVALID, "offsets are not ascending");
                cudaStream_t s = st[c % kStreams];
                if (b1 > b0) UMGAP_CUDA(cudaMemcpyAsync(d_aa + b0, aa + b0, b1 - b0, cudaMemcpyHostToDevice, s));
                UMGAP_CUDA(cudaMemcpyAsync(d_loff + l0, line_off + l0, (l1 - l0 + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
                UMGAP_CUDA(cudaMemcpyAsync(d_goff + g0, group_off + g0, (g1 - g0 + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
                peptides_range(r, tax, opts, d_aa, d_loff, d_goff, g0, g1, l0, l1, b0, b1, g1 == ngroups, d_res, s);
                UMGAP_CUDA(cudaMemcpyAsync(taxon_out + g0, d_res + g0, (g1 - g0) * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
                for (uint64_t g = g0; g < g1; ++g)
                    if (group_off[g + 1] == group_off[g]) empty.push_back(g);
                g0 = g1;
            }
            for (int i = 0; i < kStreams; ++i) UMGAP_CUDA(cudaStreamSynchronize(st[i]));
        } catch (...) {
            cudaDeviceSynchronize();
            throw;
        }
        unsigned int he[2] = {0, 0};
        UMGAP_CUDA(cudaMemcpy(he, r.d_err, sizeof he, cudaMemcpyDeviceToHost));
        if (he[0]) UMGAP_FAIL(UMGAP_ERR_UNKNOWN_TAXON, "Unknown Taxon ID: %u", he[1]);
        for (uint64_t g : empty) taxon_out[g] = UMGAP_ABSENT;  // a group without lines has no record in the reference
    });
}

uint64_t umgap_tryp_lookup_bound(uint64_t total_aa, uint64_t nlines) {
    (void)nlines;
    return total_aa + 1;
}

int umgap_tryp_lookup(const umgap_index* idx, const uint8_t* aa, const uint64_t* line_off, uint64_t nlines,
                      int minlen, int maxlen, const char* keep, const char* drop, int one_on_one,
                      uint32_t* taxa_out, uint64_t* taxa_off) {
    return guarded([&] {
        if (!idx || !line_off || !taxa_off) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        if (idx->k != 0 || !idx->var_table) UMGAP_FAIL(UMGAP_ERR_INVALID, "index is not a variable-length peptide table");
        if (minlen < 0 || maxlen < 0) UMGAP_FAIL(UMGAP_ERR_INVALID, "negative length bound");
        for (uint64_t i = 0; i <= nlines; ++i) taxa_off[i] = 0;
        const uint64_t total = nlines ? line_off[nlines] : 0;
        if (!total) return;
        if (!aa || !taxa_out) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        if (total >= (1ull << 32) || nlines >= (1ull << 32)) UMGAP_FAIL(UMGAP_ERR_INVALID, "batch too large (>= 2^32 bytes)");
        uint8_t lut[256] = {};
        const TrypParams tp = make_tryp_params(minlen, maxlen, keep, drop, lut);
        use_device(idx->device);
        const VarTable* t = (const VarTable*)idx->var_table;
        DevBuf<uint8_t> d_aa(total), d_lut(256);
        DevBuf<uint64_t> d_off(nlines + 1);
        DevBuf<uint32_t> d_lob(total), d_out(total);
        UMGAP_CUDA(cudaMemcpy(d_aa.p, aa, total, cudaMemcpyHostToDevice));
        UMGAP_CUDA(cudaMemcpy(d_lut.p, lut, 256, cudaMemcpyHostToDevice));
        UMGAP_CUDA(cudaMemcpy(d_off.p, line_off, (nlines + 1) * 8, cudaMemcpyHostToDevice));
        line_of_byte_kernel<<<(unsigned)std::min<uint64_t>(ceil_div(nlines, 8), 148 * 16), 256>>>(d_off.p, nlines, d_lob.p);
        UMGAP_CUDA(cudaGetLastError());
        tryp_lookup_kernel<<<(unsigned)std::min<uint64_t>(ceil_div(total, 256), 148 * 16), 256>>>(
            t->slots, t->nslots, t->pool, d_aa.p, d_off.p, nlines, d_lob.p, total, tp, d_lut.p, d_out.p);
        UMGAP_CUDA(cudaGetLastError());
        std::vector<uint32_t> raw(total);
        UMGAP_CUDA(cudaMemcpy(raw.data(), d_out.p, total * 4, cudaMemcpyDeviceToHost));
        uint64_t w = 0;
        for (uint64_t l = 0; l < nlines; ++l) {
            taxa_off[l] = w;
            for (uint64_t i = line_off[l]; i < line_off[l + 1]; ++i) {
                const uint32_t v = raw[i];
                if (v == kTrypNone) continue;
                if (v == kNoValue) {
                    if (one_on_one) taxa_out[w++] = 0;  // prot2tryp2lca.rs:95,130
                } else {
                    taxa_out[w++] = v;
                }
            }
        }
        taxa_off[nlines] = w;
    });
}

}  // extern "C"
