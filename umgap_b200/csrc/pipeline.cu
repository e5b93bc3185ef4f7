// Hot-path kernels and their C ABI entry points:
//   translate_lookup_kernel  reads -> six-frame translation -> rolling 9-mer keys -> one-sector
//                            table probes                      (translate.rs:114-133,
//                            dna/mod.rs:23-103, dna/translation.rs:125-144,
//                            prot2kmer2lca.rs:168-185 + fst::Map::get)
//   classify_kernel          per group: seedextend per frame record, uniq join, taxa2agg
//                            (seedextend.rs:92-176, uniq.rs:56-84, taxa2agg.rs:159-181)
//   translate_kernel, kmer_lookup_kernel, seedextend_kernel, aggregate_kernel: the same stages
//                            one at a time, behind the per-command entry points.
#include <algorithm>

#include "index.h"
#include "warp_agg.cuh"

namespace umgap {

// ---- genetic codes (NCBI gc.prt, TCAG order; the same 19 tables as translation.rs:47-104) ----
struct GeneticCode {
    int id;
    const char* aas;
    const char* starts;
};
static const GeneticCode kCodes[] = {
    {1, "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
     "---M---------------M---------------M----------------------------"},
    {2, "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNKKSS**VVVVAAAADDEEGGGG",
     "--------------------------------MMMM---------------M------------"},
    {3, "FFLLSSSSYY**CCWWTTTTPPPPHHQQRRRRIIMMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
     "----------------------------------MM----------------------------"},
    {4, "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
     "--MM---------------M------------MMMM---------------M------------"},
    {5, "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNKKSSSSVVVVAAAADDEEGGGG",
     "---M----------------------------MMMM---------------M------------"},
    {6, "FFLLSSSSYYQQCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
     "-----------------------------------M----------------------------"},
    {9, "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNNKSSSSVVVVAAAADDEEGGGG",
     "-----------------------------------M---------------M------------"},
    {10, "FFLLSSSSYY**CCCWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
     "-----------------------------------M----------------------------"},
    {11, "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
     "---M---------------M------------MMMM---------------M------------"},
    {12, "FFLLSSSSYY**CC*WLLLSPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
     "-------------------M---------------M----------------------------"},
    {13, "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNKKSSGGVVVVAAAADDEEGGGG",
     "---M------------------------------MM---------------M------------"},
    {14, "FFLLSSSSYYY*CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNNKSSSSVVVVAAAADDEEGGGG",
     "-----------------------------------M----------------------------"},
    {15, "FFLLSSSSYY*QCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
     "-----------------------------------M----------------------------"},
    {16, "FFLLSSSSYY*LCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
     "-----------------------------------M----------------------------"},
    {21, "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNNKSSSSVVVVAAAADDEEGGGG",
     "-----------------------------------M---------------M------------"},
    {22, "FFLLSS*SYY*LCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
     "-----------------------------------M----------------------------"},
    {23, "FF*LSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
     "--------------------------------M--M---------------M------------"},
};

// 65-entry codon -> residue table: [0..63] in TCAG order, [64] = codon holding an N ('-',
// translation.rs:126).  `ascii` for the translate stage, `code` (5-bit index alphabet code,
// 0xFF = residue that no index key contains) for the fused lookup.
struct CodonLut {
    uint8_t v[72];
};

static void make_ascii_lut(int table, int methionine, CodonLut& lut) {
    const GeneticCode* gc = nullptr;
    for (const GeneticCode& c : kCodes)
        if (c.id == table) gc = &c;
    if (!gc) UMGAP_FAIL(UMGAP_ERR_INVALID, "Unknown table: %d", table);  // translation.rs:176-185
    for (int i = 0; i < 64; ++i)
        lut.v[i] = (methionine && gc->starts[i] == 'M') ? 'M' : (uint8_t)gc->aas[i];
    lut.v[64] = '-';
}

void make_ascii_lut_public(int table, int methionine, uint8_t* out65) {
    CodonLut a;
    make_ascii_lut(table, methionine, a);
    memcpy(out65, a.v, 65);
}

static void make_code_lut(const umgap_index* idx, int table, int methionine, CodonLut& lut) {
    CodonLut a;
    make_ascii_lut(table, methionine, a);
    for (int i = 0; i < 65; ++i) lut.v[i] = idx->code_of_byte[a.v[i]];
}

// nucleotide byte -> 0..3 in codon order T,C,A,G (translation.rs:20); anything else, lowercase
// included, is N = 4 (dna/mod.rs:34-44).  Complement is code ^ 2 (dna/mod.rs:23-31).
__device__ __forceinline__ uint32_t nt_code(uint8_t c) {
    return c == 'T' ? 0u : c == 'C' ? 1u : c == 'A' ? 2u : c == 'G' ? 3u : 4u;
}

constexpr int kTile = 128;          // k-mer start positions per warp pass (4 per lane)
constexpr int kLookupWarps = 8;     // warps (= reads in flight) per CTA
constexpr int kQueue = 2 * kTile;   // pending second probes of one tile (both strands)

// One warp per read.  Per tile of 128 start positions the warp stages the nucleotide codes in
// shared memory, translates every codon start once for both strands (F = forward codon at x,
// R = codon of the reverse strand whose lowest forward coordinate is x), then each lane packs
// 4 forward + 4 reverse k-mers and issues their 8 sector loads back to back.  First probes are
// resolved branch-free; the few lookups that ended on a flagged sector are compacted into a
// per-warp shared-memory queue and re-probed 32 at a time, so the warp stays converged.
// ids layout: forward k-mer starting at p -> ids[2*off + p]; reverse-strand k-mer starting at
// reverse coordinate q -> ids[2*off + n + q]  (frame f record = entries f-1, f+2, f+5, ...).
template <int K>
__global__ void __launch_bounds__(kLookupWarps * 32)
translate_lookup_kernel(TableView t, CodonLut lut, const uint8_t* __restrict__ nt,
                        const uint64_t* __restrict__ read_off, uint64_t nreads,
                        uint32_t* __restrict__ ids) {
    constexpr int W = kTile + 3 * (K - 1);  // codon starts needed per tile
    __shared__ uint8_t s_lut[72];
    __shared__ uint8_t s_nt[kLookupWarps][W + 2 + 2];
    __shared__ uint8_t s_f[kLookupWarps][W + 4];
    __shared__ uint8_t s_r[kLookupWarps][W + 4];
    __shared__ uint64_t q_h[kLookupWarps][kQueue];    // hash | next distance << 45 | level << 48
    __shared__ uint32_t q_pos[kLookupWarps][kQueue];  // index into the read's ids slice
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1;
    if (threadIdx.x < 72) s_lut[threadIdx.x] = lut.v[threadIdx.x];
    __syncthreads();
    const ulonglong4* __restrict__ level0 = t.level[0];
    const uint32_t nlines0 = t.nlines[0];
    const uint64_t nwarps = (uint64_t)gridDim.x * kLookupWarps;
    for (uint64_t r = (uint64_t)blockIdx.x * kLookupWarps + warp; r < nreads; r += nwarps) {
        const uint64_t off = read_off[r];
        const uint32_t n = (uint32_t)(read_off[r + 1] - off);
        if (n < 3u * K) continue;  // no frame reaches K residues
        const uint32_t npos = n - 3u * K + 1;
        uint32_t* out = ids + 2 * off;  // forward ids at [p], reverse ids at [n + q]
        for (uint32_t w0 = 0; w0 < npos; w0 += kTile) {
            for (int i = lane; i < W + 2; i += 32) {
                const uint32_t x = w0 + i;
                s_nt[warp][i] = x < n ? (uint8_t)nt_code(nt[off + x]) : (uint8_t)4;
            }
            __syncwarp();
            for (int i = lane; i < W; i += 32) {
                const uint32_t a = s_nt[warp][i], b = s_nt[warp][i + 1], c = s_nt[warp][i + 2];
                const bool has_n = ((a | b | c) & 4u) != 0;
                s_f[warp][i] = s_lut[has_n ? 64 : 16 * a + 4 * b + c];
                s_r[warp][i] = s_lut[has_n ? 64 : 16 * (c ^ 2) + 4 * (b ^ 2) + (a ^ 2)];
            }
            __syncwarp();
            uint64_t h[8];       // 0..3 forward, 4..7 reverse
            ulonglong4 sec[8];
            bool valid[8];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int pl = lane + 32 * u;
                uint64_t kf = 0, kr = 0;
                uint32_t bad_f = 0, bad_r = 0;
#pragma unroll
                for (int i = 0; i < K; ++i) {
                    const uint32_t cf = s_f[warp][pl + 3 * i];
                    const uint32_t cr = s_r[warp][pl + 3 * (K - 1 - i)];
                    bad_f |= cf;
                    bad_r |= cr;
                    kf = (kf << 5) | (cf & 31u);
                    kr = (kr << 5) | (cr & 31u);
                }
                const bool live = w0 + pl < npos;
                valid[u] = live && !(bad_f & 0x80u);
                valid[u + 4] = live && !(bad_r & 0x80u);
                h[u] = mix45(kf);
                h[u + 4] = mix45(kr);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (valid[u]) sec[u] = load_sector(level0 + probe_sector(h[u], nlines0, 0));
            uint32_t qn = 0;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const uint32_t p = w0 + lane + 32 * (u & 3);
                const uint32_t pos = u < 4 ? p : n + (npos - 1 - p);
                bool more = false;
                uint32_t v = kNoValue;
                if (valid[u]) v = probe_sector_data(sec[u], (uint32_t)h[u] & kTagMask, more);
                if (p < npos && !more) out[pos] = v;
                const unsigned m = __ballot_sync(0xffffffffu, more);
                if (more) {
                    const uint32_t at = qn + __popc(m & lt_mask);
                    q_h[warp][at] = h[u] | (1ull << 45);
                    q_pos[warp][at] = pos;
                }
                qn += __popc(m);
            }
            __syncwarp();
            // re-probe the flagged ones, densely packed: distance d, then d+1, ... then next level
            while (qn) {
                uint32_t qnext = 0;
                for (uint32_t c = 0; c < qn; c += 32) {
                    const uint32_t i = c + lane;
                    const bool active = i < qn;
                    uint64_t hq = 0;
                    uint32_t pos = 0;
                    if (active) {
                        hq = q_h[warp][i];
                        pos = q_pos[warp][i];
                    }
                    uint32_t d = (uint32_t)(hq >> 45) & 7u;
                    uint32_t lv = (uint32_t)(hq >> 48);
                    const uint64_t hh = hq & kKeyMask;
                    bool more = false;
                    if (active) {
                        const ulonglong4 s2 = load_sector(t.level[lv] + probe_sector(hh, t.nlines[lv], d));
                        const uint32_t v = probe_sector_data(s2, (d << 28) | ((uint32_t)hh & kTagMask), more);
                        if (more && ++d == (uint32_t)kMaxDisp) {
                            d = 0;
                            if (++lv == (uint32_t)t.nlevels) more = false;  // v is kNoValue here
                        }
                        if (!more) out[pos] = v;
                    }
                    __syncwarp();
                    const unsigned m = __ballot_sync(0xffffffffu, more);
                    if (more) {
                        const uint32_t at = qnext + __popc(m & lt_mask);
                        q_h[warp][at] = hh | ((uint64_t)d << 45) | ((uint64_t)lv << 48);
                        q_pos[warp][at] = pos;
                    }
                    qnext += __popc(m);
                }
                __syncwarp();
                qn = qnext;
            }
        }
    }
}

// ---- classify: seedextend + uniq join + aggregate, one warp per group ----------------------------
struct ClassifyParams {
    int k;
    int one_on_one;
    int seedextend;
    uint32_t min_seed, max_gap;
    AggParams agg;
};

constexpr int kAggWarps = 4;
constexpr uint32_t kAggCap = 512;  // ids of a 2 x 150 nt pair: <= 496

struct DevError {  // first error raised by a kernel
    unsigned int flag;
    unsigned int taxon;
};

__global__ void __launch_bounds__(kAggWarps * 32)
classify_kernel(TaxView tv, ClassifyParams cp, const uint32_t* __restrict__ ids,
                const uint64_t* __restrict__ read_off, const uint64_t* __restrict__ group_off,
                uint64_t ngroups, uint32_t* __restrict__ scratch, uint32_t* __restrict__ taxon_out,
                DevError* err) {
    __shared__ uint32_t s_a[kAggWarps][kAggCap];
    __shared__ uint32_t s_p[kAggWarps][kAggCap + 1];
    __shared__ uint32_t s_l[kAggWarps][kAggCap];
    __shared__ uint32_t s_cnt[kAggWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t nwarps = (uint64_t)gridDim.x * kAggWarps;
    for (uint64_t g = (uint64_t)blockIdx.x * kAggWarps + warp; g < ngroups; g += nwarps) {
        const uint64_t r0 = group_off[g], r1 = group_off[g + 1];
        const uint64_t nrec = (r1 - r0) * 6;
        uint32_t* A = s_a[warp];
        uint32_t* P = s_p[warp];
        uint32_t* L = s_l[warp];
        uint32_t cap = kAggCap;
        bool present = false;
        uint32_t total = 0;
        for (int attempt = 0; attempt < 2; ++attempt) {
            if (lane == 0) s_cnt[warp] = 0;
            __syncwarp();
            bool any = false;
            for (uint64_t rec = lane; rec < nrec; rec += 32) {
                const uint64_t r = r0 + rec / 6;
                const uint32_t fr = (uint32_t)(rec % 6);  // 0,1,2 forward; 3,4,5 reverse
                const uint64_t off = read_off[r];
                const uint32_t n = (uint32_t)(read_off[r + 1] - off);
                const uint32_t f = fr % 3;
                const uint32_t plen = n >= f ? (n - f) / 3 : 0;  // peptide length of the frame
                if (plen < (uint32_t)cp.k) continue;             // record dropped (:172)
                any = true;
                const uint32_t cnt = plen - cp.k + 1;
                const uint32_t* base = ids + 2 * off + (fr >= 3 ? n : 0) + f;
                auto push = [&](uint32_t v) {
                    if (v == 0) return;  // taxa2agg.rs:169
                    const uint32_t at = atomicAdd(&s_cnt[warp], 1u);
                    if (at < cap) A[at] = v;
                };
                if (cp.seedextend) {
                    seedextend_stream(base, 3, cnt, cp.one_on_one != 0, cp.min_seed, cp.max_gap, push);
                } else {
                    for (uint32_t i = 0; i < cnt; ++i) {
                        const uint32_t v = base[3 * i];
                        if (v != kNoValue) push(v);
                    }
                }
            }
            present = __any_sync(0xffffffffu, any);
            __syncwarp();
            total = s_cnt[warp];
            if (total <= cap) break;
            // rare: more kept ids than the shared-memory list holds -> redo into global scratch
            // (the group's own slice of a buffer as large as `ids`, split in three)
            // (scratch holds 6 words per nucleotide: the group's slice is split in three lists of
            // gsize words; total < gsize because a read yields fewer k-mers than 2x its length)
            const uint64_t gbase = 6 * read_off[r0];
            const uint64_t gsize = 2 * (read_off[r1] - read_off[r0]);
            A = scratch + gbase;
            P = scratch + gbase + gsize;
            L = scratch + gbase + 2 * gsize;
            cap = 0xFFFFFFFFu;
            __syncwarp();
        }
        uint32_t res;
        if (!present) {
            res = UMGAP_ABSENT;
        } else {
            uint32_t bad = 0;
            res = warp_aggregate(tv, A, P, L, total, cp.agg, lane, &bad);
            const uint32_t bad_any = __reduce_max_sync(0xffffffffu, bad);
            if (res == kAggUnknown) {
                if (lane == 0 && atomicCAS(&err->flag, 0u, 1u) == 0u) err->taxon = bad_any;
                res = UMGAP_ABSENT;
            }
        }
        if (lane == 0) taxon_out[g] = res;
        __syncwarp();
    }
}

}  // namespace umgap

using namespace umgap;

// ---- workspace slots of an index handle --------------------------------------------------------
enum { WS_IDS = 0, WS_SCRATCH = 1, WS_ERR = 2, WS_NT = 3, WS_ROFF = 5, WS_GOFF = 7, WS_OUT = 9 };

static ClassifyParams make_params(const umgap_index* idx, const umgap_pipeline_opts* o) {
    ClassifyParams cp{};
    cp.k = idx->k;
    cp.one_on_one = o->one_on_one;
    cp.seedextend = o->seedextend;
    cp.min_seed = (uint32_t)std::max(0, o->min_seed_size);
    cp.max_gap = (uint32_t)std::max(0, o->max_gap_size);
    cp.agg.strategy = o->strategy;
    cp.agg.factor = o->factor;
    cp.agg.lower_bound = o->lower_bound;
    cp.agg.ranked_only = o->ranked_only;
    return cp;
}

static void check_opts(const umgap_index* idx, const umgap_taxonomy* tax, const umgap_pipeline_opts* o) {
    if (!idx || !o) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
    if (idx->k <= 0) UMGAP_FAIL(UMGAP_ERR_INVALID, "index is not a fixed-length k-mer table");
    if (tax && tax->device != idx->device)
        UMGAP_FAIL(UMGAP_ERR_INVALID, "index and taxonomy live on different devices");
    if (o->strategy < UMGAP_AGG_LCA_STAR || o->strategy > UMGAP_AGG_MRTL)
        UMGAP_FAIL(UMGAP_ERR_INVALID, "unknown aggregation strategy %d", o->strategy);
}

// ---- optional per-launch timing (umgap_kernel_timing) --------------------------------------------
namespace {
struct TimedLaunch {
    cudaEvent_t a, b;
    int kind;  // 0 lookup, 1 classify
};
bool g_timing = false;
std::vector<TimedLaunch> g_launches;
std::vector<cudaEvent_t> g_event_pool;
cudaEvent_t take_event() {
    if (!g_event_pool.empty()) {
        cudaEvent_t e = g_event_pool.back();
        g_event_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    UMGAP_CUDA(cudaEventCreate(&e));
    return e;
}
struct LaunchTimer {
    cudaStream_t st;
    TimedLaunch t{};
    bool on;
    LaunchTimer(int kind, cudaStream_t s) : st(s), on(g_timing) {
        if (!on) return;
        t.kind = kind;
        t.a = take_event();
        t.b = take_event();
        UMGAP_CUDA(cudaEventRecord(t.a, st));
    }
    void stop() {
        if (!on) return;
        UMGAP_CUDA(cudaEventRecord(t.b, st));
        g_launches.push_back(t);
    }
};
}  // namespace

static void launch_translate_lookup(const umgap_index* idx, const umgap_pipeline_opts* o,
                                    const uint8_t* nt_dev, const uint64_t* read_off_dev,
                                    uint64_t nreads, uint32_t* ids_dev, cudaStream_t st) {
    if (!nreads) return;
    CodonLut lut{};
    make_code_lut(idx, o->table, o->methionine, lut);
    const TableView tv = idx->view();
    const unsigned blocks = (unsigned)std::min<uint64_t>(ceil_div(nreads, kLookupWarps), 148ull * 32);
    LaunchTimer timer(0, st);
    switch (idx->k) {
#define UMGAP_CASE(KK)                                                                           \
    case KK:                                                                                     \
        translate_lookup_kernel<KK><<<blocks, kLookupWarps * 32, 0, st>>>(tv, lut, nt_dev,        \
                                                                          read_off_dev, nreads,  \
                                                                          ids_dev);              \
        break;
        UMGAP_CASE(1) UMGAP_CASE(2) UMGAP_CASE(3) UMGAP_CASE(4) UMGAP_CASE(5) UMGAP_CASE(6)
        UMGAP_CASE(7) UMGAP_CASE(8) UMGAP_CASE(9)
#undef UMGAP_CASE
        default:
            UMGAP_FAIL(UMGAP_ERR_INVALID, "unsupported k %d", idx->k);
    }
    UMGAP_CUDA(cudaGetLastError());
    timer.stop();
}

static void launch_classify(const umgap_index* idx, const umgap_taxonomy* tax,
                            const umgap_pipeline_opts* o, const uint32_t* ids_dev,
                            const uint64_t* read_off_dev, const uint64_t* group_off_dev,
                            uint64_t ngroups, uint32_t* scratch_dev, uint32_t* out_dev, DevError* err,
                            cudaStream_t st) {
    if (!ngroups) return;
    const unsigned blocks = (unsigned)std::min<uint64_t>(ceil_div(ngroups, kAggWarps), 148ull * 64);
    LaunchTimer timer(1, st);
    classify_kernel<<<blocks, kAggWarps * 32, 0, st>>>(tax->view, make_params(idx, o), ids_dev,
                                                       read_off_dev, group_off_dev, ngroups,
                                                       scratch_dev, out_dev, err);
    UMGAP_CUDA(cudaGetLastError());
    timer.stop();
}

static void raise_dev_error(const DevError& e) {
    if (e.flag) UMGAP_FAIL(UMGAP_ERR_UNKNOWN_TAXON, "Unknown Taxon ID: %u", e.taxon);
}

extern "C" {

int umgap_kernel_timing(int enable) {
    g_timing = enable != 0;
    return UMGAP_OK;
}

int umgap_kernel_times(double* lookup_ms, uint64_t* lookup_launches, double* classify_ms,
                       uint64_t* classify_launches) {
    return guarded([&] {
        double ms[2] = {0, 0};
        uint64_t cnt[2] = {0, 0};
        for (TimedLaunch& t : g_launches) {
            UMGAP_CUDA(cudaEventSynchronize(t.b));
            float e = 0;
            UMGAP_CUDA(cudaEventElapsedTime(&e, t.a, t.b));
            ms[t.kind] += e;
            cnt[t.kind]++;
            g_event_pool.push_back(t.a);
            g_event_pool.push_back(t.b);
        }
        g_launches.clear();
        if (lookup_ms) *lookup_ms = ms[0];
        if (lookup_launches) *lookup_launches = cnt[0];
        if (classify_ms) *classify_ms = ms[1];
        if (classify_launches) *classify_launches = cnt[1];
    });
}

void umgap_pipeline_opts_default(umgap_pipeline_opts* o) {
    if (!o) return;
    o->table = 1;
    o->methionine = 0;
    o->one_on_one = 1;
    o->seedextend = 1;
    o->min_seed_size = 2;
    o->max_gap_size = 0;
    o->strategy = UMGAP_AGG_HYBRID;
    o->factor = 0.25f;
    o->lower_bound = 0.0f;
    o->ranked_only = 0;
}

int umgap_translate_lookup_dev(const umgap_index* idx, const umgap_pipeline_opts* opts,
                               const uint8_t* nt_dev, const uint64_t* read_off_dev, uint64_t nreads,
                               uint64_t total_nt, uint32_t* ids_dev, void* stream) {
    (void)total_nt;
    return guarded([&] {
        check_opts(idx, nullptr, opts);
        use_device(idx->device);
        launch_translate_lookup(idx, opts, nt_dev, read_off_dev, nreads, ids_dev, (cudaStream_t)stream);
    });
}

int umgap_classify_reads_dev(const umgap_index* idx, const umgap_taxonomy* tax,
                             const umgap_pipeline_opts* opts, const uint8_t* nt_dev,
                             const uint64_t* read_off_dev, uint64_t nreads, uint64_t total_nt,
                             const uint64_t* group_off_dev, uint64_t ngroups, uint32_t* taxon_out_dev,
                             void* stream) {
    return guarded([&] {
        check_opts(idx, tax, opts);
        if (!tax) UMGAP_FAIL(UMGAP_ERR_INVALID, "null taxonomy");
        use_device(idx->device);
        cudaStream_t st = (cudaStream_t)stream;
        uint32_t* ids = (uint32_t*)idx->ws.get(WS_IDS, (2 * total_nt + 64) * sizeof(uint32_t));
        uint32_t* scratch = (uint32_t*)idx->ws.get(WS_SCRATCH, (6 * total_nt + 64) * sizeof(uint32_t));
        DevError* err = (DevError*)idx->ws.get(WS_ERR, sizeof(DevError));
        UMGAP_CUDA(cudaMemsetAsync(err, 0, sizeof(DevError), st));
        launch_translate_lookup(idx, opts, nt_dev, read_off_dev, nreads, ids, st);
        launch_classify(idx, tax, opts, ids, read_off_dev, group_off_dev, ngroups, scratch,
                        taxon_out_dev, err, st);
    });
}

int umgap_classify_reads(const umgap_index* idx, const umgap_taxonomy* tax,
                         const umgap_pipeline_opts* opts, const uint8_t* nt, const uint64_t* read_off,
                         uint64_t nreads, const uint64_t* group_off, uint64_t ngroups,
                         uint32_t* taxon_out, uint64_t* n_lookups) {
    return guarded([&] {
        check_opts(idx, tax, opts);
        if (!tax) UMGAP_FAIL(UMGAP_ERR_INVALID, "null taxonomy");
        if ((nreads && (!nt || !read_off)) || (ngroups && (!group_off || !taxon_out)))
            UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        use_device(idx->device);
        if (n_lookups) {
            uint64_t c = 0;
            const uint64_t span = 3ull * idx->k;
            for (uint64_t r = 0; r < nreads; ++r) {
                const uint64_t n = read_off[r + 1] - read_off[r];
                if (n >= span) c += 2 * (n - span + 1);
            }
            *n_lookups = c;
        }
        if (!ngroups) return;
        if (group_off[ngroups] != nreads || group_off[0] != 0)
            UMGAP_FAIL(UMGAP_ERR_INVALID, "group_off must cover reads 0..nreads");
        // Chunked, double-buffered: while chunk c computes on its stream, chunk c+1 uploads.
        const uint64_t kChunkNt = 96ull << 20;  // nucleotides per chunk
        cudaStream_t st[2];
        cudaEvent_t done[2];
        for (int i = 0; i < 2; ++i) {
            UMGAP_CUDA(cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking));
            UMGAP_CUDA(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
        }
        DevError* err = (DevError*)idx->ws.get(WS_ERR, sizeof(DevError));
        UMGAP_CUDA(cudaMemset(err, 0, sizeof(DevError)));
        std::vector<uint64_t> h_roff[2], h_goff[2];
        uint64_t g0 = 0;
        int buf = 0;
        bool used[2] = {false, false};
        try {
            while (g0 < ngroups) {
                // extend the chunk group by group up to kChunkNt nucleotides (at least one group)
                uint64_t g1 = g0;
                const uint64_t nt0 = read_off[group_off[g0]];
                while (g1 < ngroups && (g1 == g0 || read_off[group_off[g1 + 1]] - nt0 <= kChunkNt)) ++g1;
                const uint64_t r0 = group_off[g0], r1 = group_off[g1];
                const uint64_t cnt_nt = read_off[r1] - nt0, cnt_r = r1 - r0, cnt_g = g1 - g0;
                if (used[buf]) UMGAP_CUDA(cudaEventSynchronize(done[buf]));  // host vectors reusable
                h_roff[buf].resize(cnt_r + 1);
                for (uint64_t i = 0; i <= cnt_r; ++i) h_roff[buf][i] = read_off[r0 + i] - nt0;
                h_goff[buf].resize(cnt_g + 1);
                for (uint64_t i = 0; i <= cnt_g; ++i) h_goff[buf][i] = group_off[g0 + i] - r0;
                uint8_t* d_nt = (uint8_t*)idx->ws.get(WS_NT + buf, cnt_nt + 64);
                uint64_t* d_roff = (uint64_t*)idx->ws.get(WS_ROFF + buf, (cnt_r + 1) * 8);
                uint64_t* d_goff = (uint64_t*)idx->ws.get(WS_GOFF + buf, (cnt_g + 1) * 8);
                uint32_t* d_out = (uint32_t*)idx->ws.get(WS_OUT + buf, cnt_g * 4 + 16);
                // ids/scratch are shared by both chunks' kernels: kernels of consecutive chunks
                // are ordered through `done` below, copies overlap freely.
                uint32_t* ids = (uint32_t*)idx->ws.get(WS_IDS, (2 * std::max(cnt_nt, kChunkNt) + 64) * 4);
                uint32_t* scratch = (uint32_t*)idx->ws.get(WS_SCRATCH, (6 * std::max(cnt_nt, kChunkNt) + 64) * 4);
                cudaStream_t s = st[buf];
                UMGAP_CUDA(cudaMemcpyAsync(d_nt, nt + nt0, cnt_nt, cudaMemcpyHostToDevice, s));
                UMGAP_CUDA(cudaMemcpyAsync(d_roff, h_roff[buf].data(), (cnt_r + 1) * 8, cudaMemcpyHostToDevice, s));
                UMGAP_CUDA(cudaMemcpyAsync(d_goff, h_goff[buf].data(), (cnt_g + 1) * 8, cudaMemcpyHostToDevice, s));
                if (used[buf ^ 1]) UMGAP_CUDA(cudaStreamWaitEvent(s, done[buf ^ 1], 0));
                launch_translate_lookup(idx, opts, d_nt, d_roff, cnt_r, ids, s);
                launch_classify(idx, tax, opts, ids, d_roff, d_goff, cnt_g, scratch, d_out, err, s);
                UMGAP_CUDA(cudaEventRecord(done[buf], s));
                UMGAP_CUDA(cudaMemcpyAsync(taxon_out + g0, d_out, cnt_g * 4, cudaMemcpyDeviceToHost, s));
                used[buf] = true;
                buf ^= 1;
                g0 = g1;
            }
            for (int i = 0; i < 2; ++i) UMGAP_CUDA(cudaStreamSynchronize(st[i]));
            DevError he;
            UMGAP_CUDA(cudaMemcpy(&he, err, sizeof he, cudaMemcpyDeviceToHost));
            for (int i = 0; i < 2; ++i) {
                cudaStreamDestroy(st[i]);
                cudaEventDestroy(done[i]);
            }
            raise_dev_error(he);
        } catch (...) {
            cudaDeviceSynchronize();
            throw;
        }
    });
}

}  // extern "C"
