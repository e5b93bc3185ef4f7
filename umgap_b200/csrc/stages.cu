// Per-command entry points: one reference command loop each, host buffers in and out.
//   umgap_translate    translate.rs:114-133       umgap_kmer_lookup  prot2kmer2lca.rs:168-185
//   umgap_seedextend   seedextend.rs:92-176       umgap_aggregate    taxa2agg.rs:159-181
#include <algorithm>

#include "index.h"
#include "warp_agg.cuh"

namespace umgap {

struct AsciiLut {
    uint8_t v[72];
};
void make_ascii_lut_public(int table, int methionine, uint8_t* out65);  // pipeline.cu

__device__ __forceinline__ uint32_t nt_code2(uint8_t c) {
    return c == 'T' ? 0u : c == 'C' ? 1u : c == 'A' ? 2u : c == 'G' ? 3u : 4u;
}

// One warp per read; every lane translates whole codons of the selected frames.  Peptide j of
// read r starts at aa_off[r*nframes + j] (offsets are computed on the host from the read
// lengths: frame f of an n-nt read has max(n-f,0)/3 residues).
__global__ void __launch_bounds__(256)
translate_kernel(AsciiLut lut, const uint8_t* __restrict__ nt, const uint64_t* __restrict__ read_off,
                 uint64_t nreads, uint32_t frames_mask, int nframes,
                 const uint64_t* __restrict__ aa_off, uint8_t* __restrict__ aa) {
    __shared__ uint8_t s_lut[72];
    if (threadIdx.x < 72) s_lut[threadIdx.x] = lut.v[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t r = warp; r < nreads; r += nwarps) {
        const uint64_t off = read_off[r];
        const uint32_t n = (uint32_t)(read_off[r + 1] - off);
        int slot = 0;
        for (int fr = 0; fr < 6; ++fr) {
            if (!(frames_mask >> fr & 1)) continue;
            const uint32_t f = fr % 3;
            const bool rev = fr >= 3;
            const uint32_t plen = n >= f ? (n - f) / 3 : 0;
            uint8_t* out = aa + aa_off[r * nframes + slot];
            ++slot;
            for (uint32_t j = lane; j < plen; j += 32) {
                uint32_t a, b, c;
                if (!rev) {
                    const uint8_t* p = nt + off + f + 3 * j;
                    a = nt_code2(p[0]);
                    b = nt_code2(p[1]);
                    c = nt_code2(p[2]);
                } else {  // reverse strand position i is the complement of forward n-1-i
                    const uint8_t* p = nt + off + (n - 1 - f - 3 * j);
                    a = nt_code2(p[0]);
                    b = nt_code2(*(p - 1));
                    c = nt_code2(*(p - 2));
                    if (!((a | b | c) & 4u)) {
                        a ^= 2u;
                        b ^= 2u;
                        c ^= 2u;
                    }
                }
                out[j] = s_lut[((a | b | c) & 4u) ? 64 : 16 * a + 4 * b + c];
            }
        }
    }
}

// One warp per peptide; lanes stride the k-mer start positions.  out has one entry per
// position (UMGAP_MISS for a miss); the host applies -o / compaction.
template <class TV>
__global__ void __launch_bounds__(256)
kmer_lookup_kernel(const __grid_constant__ TV t, const uint8_t* __restrict__ code_of_byte,
                   const uint8_t* __restrict__ aa, const uint64_t* __restrict__ pep_off,
                   uint64_t npeps, const uint64_t* __restrict__ out_off, uint32_t* __restrict__ out) {
    __shared__ uint8_t s_code[256];
    s_code[threadIdx.x & 255] = code_of_byte[threadIdx.x & 255];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const int k = t.k;
    for (uint64_t p = warp; p < npeps; p += nwarps) {
        const uint64_t off = pep_off[p];
        const uint64_t len = pep_off[p + 1] - off;
        if (len < (uint64_t)k) continue;
        const uint64_t cnt = len - k + 1;
        uint32_t* o = out + out_off[p];
        for (uint64_t i = lane; i < cnt; i += 32) {
            uint64_t key = 0;
            uint32_t bad = 0;
            for (int j = 0; j < k; ++j) {
                const uint32_t c = s_code[aa[off + i + j]];
                bad |= c;
                key = (key << 5) | (c & 31u);
            }
            o[i] = table_lookup(t, (bad & 0x80u) ? kInvalidKey : key);
        }
    }
}

// One lane per record.
__global__ void __launch_bounds__(128)
seedextend_kernel(const uint32_t* __restrict__ taxa, const uint64_t* __restrict__ rec_off,
                  uint64_t nrecs, uint32_t min_seed, uint32_t max_gap, uint32_t* __restrict__ out,
                  uint32_t* __restrict__ out_len) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < nrecs; r += stride) {
        const uint64_t off = rec_off[r];
        const uint32_t cnt = (uint32_t)(rec_off[r + 1] - off);
        uint32_t* o = out + off;
        uint32_t m = 0;
        seedextend_stream(taxa + off, 1, cnt, true, min_seed, max_gap,
                          [&](uint32_t v) { o[m++] = v; });
        out_len[r] = m;
    }
}

// seedextend -r (seedextend.rs:151-164): of the extended seeds of a record only the one with the highest summed rank
// score survives (`max_by_key`: the last of equal maxima); an id scores TaxonList::score, `penalty` when that is None
// (unranked lineage, species and below -- see taxonomy.cu -- the 0 of a miss, ids the taxonomy does not hold).
// One lane per record, the reference's loop as written (:101-149), over the ids plus the sentinel 0.
__global__ void __launch_bounds__(128)
seedextend_ranked_kernel(TaxView tv, const uint32_t* __restrict__ taxa, const uint64_t* __restrict__ rec_off, uint64_t nrecs,
                         uint32_t min_seed, uint32_t max_gap, uint32_t penalty, uint32_t* __restrict__ out,
                         uint32_t* __restrict__ out_len) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < nrecs; r += stride) {
        const uint64_t off = rec_off[r];
        const uint32_t n = (uint32_t)(rec_off[r + 1] - off);
        const uint32_t* ids = taxa + off;
        auto at = [&](uint32_t i) { return i < n ? ids[i] : 0u; };
        auto score = [&](uint32_t t) -> uint64_t {
            if (t > tv.max_id) return penalty;  // (taxon 0 is the reference's "unknown" taxon unless the file defines it)
            const uint32_t d = __ldg(tv.dense_of + t);
            if (d == kNoTaxon) return penalty;
            const uint32_t s = __ldg(tv.seed_score + d);
            return s ? s : penalty;
        };
        bool have = false;
        uint32_t bs = 0, be = 0;
        uint64_t best = 0;
        auto seed = [&](uint32_t s, uint32_t e) {
            if (e < s) e = s;
            uint64_t sum = 0;
            for (uint32_t i = s; i < e; ++i) sum += score(at(i));
            if (!have || sum >= best) {
                have = true;
                best = sum;
                bs = s;
                be = e;
            }
        };
        const uint32_t len = n + 1;
        uint32_t start = 0, end = 1, last_tid = at(0), same_tid = 1, same_max = 1;
        while (end < len) {
            const uint32_t cur = at(end);
            if (last_tid == cur) {
                ++same_tid;
                ++end;
                continue;
            }
            if (last_tid == 0 && same_tid > max_gap) {
                if (same_max >= min_seed) seed(start, end - same_tid);
                start = end;
                last_tid = cur;
                same_tid = 1;
                same_max = 1;
                ++end;
                continue;
            }
            if (last_tid == 0 && (end - start) == same_tid) {
                ++end;
                start = end;
                continue;
            }
            if (last_tid != 0) same_max = max(same_max, same_tid);
            last_tid = cur;
            same_tid = 1;
            ++end;
        }
        if (same_max >= min_seed) {
            if (last_tid == 0) end -= same_tid;
            seed(start, end);
        }
        uint32_t m = 0;
        if (have)
            for (uint32_t i = bs; i < be; ++i) out[off + m++] = at(i);
        out_len[r] = m;
    }
}

constexpr int kStageAggWarps = 4;
constexpr uint32_t kStageAggCap = 512;

// One warp per record; small records aggregate in shared memory, larger ones in their slice of
// a global scratch buffer (3 words per input id).
__global__ void __launch_bounds__(kStageAggWarps * 32)
aggregate_kernel(TaxView tv, AggParams ap, const uint32_t* __restrict__ taxa,
                 const uint64_t* __restrict__ rec_off, uint64_t nrecs, uint32_t* __restrict__ scratch,
                 uint32_t* __restrict__ out, unsigned int* err) {
    __shared__ uint32_t s_a[kStageAggWarps][kStageAggCap];
    __shared__ uint32_t s_p[kStageAggWarps][kStageAggCap + 1];
    __shared__ uint32_t s_l[kStageAggWarps][kStageAggCap];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t nwarps = (uint64_t)gridDim.x * kStageAggWarps;
    for (uint64_t r = (uint64_t)blockIdx.x * kStageAggWarps + warp; r < nrecs; r += nwarps) {
        const uint64_t off = rec_off[r];
        const uint64_t cnt = rec_off[r + 1] - off;
        uint32_t *A, *P, *L;
        if (cnt <= kStageAggCap) {
            A = s_a[warp];
            P = s_p[warp];
            L = s_l[warp];
        } else {  // slice [3*off + r, 3*(off+cnt) + r + 1): cnt + (cnt+1) + cnt words
            A = scratch + 3 * off + r;
            P = A + cnt;
            L = P + cnt + 1;
        }
        // drop zeros (taxa2agg.rs:169), order is irrelevant to the aggregators
        uint32_t n = 0;
        for (uint64_t base = 0; base < cnt; base += 32) {
            const uint64_t i = base + lane;
            const uint32_t v = i < cnt ? taxa[off + i] : 0u;
            const unsigned m = __ballot_sync(0xffffffffu, v != 0);
            if (v != 0) A[n + __popc(m & ((1u << lane) - 1))] = v;
            n += __popc(m);
        }
        __syncwarp();
        uint32_t bad = 0;
        uint32_t res = warp_aggregate<false>(tv, A, nullptr, P, L, n, ap, lane, &bad);
        bad = __reduce_max_sync(0xffffffffu, bad);
        if (res == kAggUnknown) {
            if (lane == 0 && atomicCAS(&err[0], 0u, 1u) == 0u) err[1] = bad;
            res = UMGAP_ABSENT;
        }
        if (lane == 0) out[r] = res;
        __syncwarp();
    }
}

static unsigned grid_for(uint64_t items, unsigned per_block, unsigned max_blocks = 148u * 32) {
    return (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(ceil_div(items, per_block), max_blocks));
}

// The aggregation kernel over device-resident records (the fused peptide path of tryptic.cu); scratch_dev holds
// 3 * rec_off[nrecs] + nrecs + 8 words, err_dev two (flag, offending taxon id).
void launch_aggregate(const umgap_taxonomy* tax, int strategy, float factor, float lower_bound, int ranked_only,
                      const uint32_t* taxa_dev, const uint64_t* rec_off_dev, uint64_t nrecs, uint32_t* scratch_dev,
                      uint32_t* out_dev, unsigned int* err_dev, cudaStream_t st) {
    if (!nrecs) return;
    AggParams ap{strategy, factor, lower_bound, ranked_only};
    aggregate_kernel<<<grid_for(nrecs, kStageAggWarps), kStageAggWarps * 32, 0, st>>>(tax->view, ap, taxa_dev, rec_off_dev, nrecs,
                                                                                   scratch_dev, out_dev, err_dev);
    UMGAP_CUDA(cudaGetLastError());
}


// ---- taxa2agg -s: scored input (taxa2agg.rs:141-148: "taxon=score" pairs) ------------------------------------------
// The weights are f32 sums, so the ORDER of every addition is part of the result.  One thread per record, the
// reference's own orders: a taxon's score is summed in input order (agg/mod.rs:31-34); MRTL adds the ancestors' sums
// walking up from the taxon (rmq/rtl.rs:43-46); the induced tree sums a chain of single children top-down
// (tree/mod.rs:75-80) and a node's children after its own value (tree/mod.rs:95-96) -- in HashSet order there, in
// ascending taxon id here (= preorder of the sorted member list: taxonomy.cu numbers siblings by ascending id), one of
// the orders the reference can take.  Ties as in the unscored kernels: the first maximum in preorder.  An unknown taxon
// raises only if it survives the lower bound (it never reaches the aggregator otherwise, taxa2agg.rs:169-170).
// Scratch per record: its slice of A (u32) and W (f32), as long as the record.
struct ScoredRec {
    const TaxView& tv;
    const uint32_t* A;  // distinct kept members: dense index, ascending
    const float* W;     // their summed scores
};

__device__ uint32_t scored_lca(const TaxView& tv, uint32_t x, uint32_t y) {  // dense x <= y
    if (x == y || __ldg(tv.last + x) >= y) return x;
    uint32_t best = 0;
    const uint32_t dmax = min((uint32_t)__ldg(tv.depth + x), (uint32_t)__ldg(tv.depth + y));
    for (uint32_t d = 0; d <= dmax; ++d) {
        const uint32_t ax = __ldg(tv.anc + (uint64_t)x * tv.stride + d);
        if (ax != __ldg(tv.anc + (uint64_t)y * tv.stride + d)) break;
        best = ax;
    }
    return best;
}
// End of the child group that starts at member j: the members below the same child (depth cd) of the node.
__device__ uint32_t scored_group_end(const ScoredRec& r, uint32_t j, uint32_t hi, uint32_t cd) {
    const uint32_t ch = __ldg(r.tv.anc + (uint64_t)r.A[j] * r.tv.stride + cd);
    uint32_t e = j + 1;
    while (e < hi && __ldg(r.tv.anc + (uint64_t)r.A[e] * r.tv.stride + cd) == ch) ++e;
    return e;
}
// The collapsed node (tree/mod.rs:71-86) over members [lo, hi), all inside one subtree: its root and chain value;
// on return [lo, hi) are the members below it and nk says whether it has children (0, or >= 2).
__device__ void scored_node(const ScoredRec& r, uint32_t& lo, uint32_t hi, uint32_t& root, float& value, uint32_t& nk) {
    value = 0.0f;  // T::default() of the nodes that are not members (tree/mod.rs:58)
    for (;;) {
        const uint32_t c = scored_lca(r.tv, r.A[lo], r.A[hi - 1]);
        if (r.A[lo] == c) value = value + r.W[lo++];  // the node itself is a member; a chain adds top-down
        root = c;
        nk = 0;
        if (lo >= hi) return;  // a leaf of the induced tree
        const uint32_t cd = (uint32_t)__ldg(r.tv.depth + c) + 1;
        nk = scored_group_end(r, lo, hi, cd) == hi ? 1u : 2u;
        if (nk != 1) return;  // a fork ends the chain of single children
    }
}
// Aggregated value (tree/mod.rs:90-101) of the collapsed subtree over members [lo, hi): the node's own value, then its
// children's aggregated values in order.  Depth-first with an explicit stack (one frame per nested fork).
constexpr int kScoredDepth = 256;  // depths are 8-bit
__device__ float scored_subtree(const ScoredRec& r, uint32_t lo, uint32_t hi) {
    uint32_t fj[kScoredDepth], fhi[kScoredDepth], fcd[kScoredDepth];
    float fv[kScoredDepth];
    int sp = 0;
    uint32_t root, nk;
    scored_node(r, lo, hi, root, fv[0], nk);
    fj[0] = lo;
    fhi[0] = hi;
    fcd[0] = (uint32_t)__ldg(r.tv.depth + root) + 1;
    for (;;) {
        if (fj[sp] >= fhi[sp]) {  // every child added
            const float v = fv[sp];
            if (sp == 0) return v;
            --sp;
            fv[sp] = fv[sp] + v;
            continue;
        }
        uint32_t clo = fj[sp];
        const uint32_t chi = scored_group_end(r, clo, fhi[sp], fcd[sp]);
        fj[sp] = chi;
        ++sp;  // nested forks are strictly deeper: sp < 256
        scored_node(r, clo, chi, root, fv[sp], nk);
        fj[sp] = clo;
        fhi[sp] = chi;
        fcd[sp] = (uint32_t)__ldg(r.tv.depth + root) + 1;
    }
}

__global__ void aggregate_scored_kernel(const TaxView tv, AggParams ap, const uint32_t* __restrict__ taxa, const float* __restrict__ scores,
                                        const uint64_t* __restrict__ rec_off, uint64_t nrecs, uint32_t* __restrict__ sa,
                                        float* __restrict__ sw, uint32_t* __restrict__ out, unsigned int* __restrict__ err) {
    const uint64_t rec = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (rec >= nrecs) return;
    const uint64_t lo = rec_off[rec], hi = rec_off[rec + 1];
    uint32_t* A = sa + lo;
    float* W = sw + lo;
    // count (agg/mod.rs:27-36): first sight appends, later sights add in input order; zeros are dropped (taxa2agg.rs:169)
    uint32_t m = 0;
    for (uint64_t i = lo; i < hi; ++i) {
        const uint32_t id = taxa[i];
        if (id == 0) continue;
        uint32_t j = 0;
        while (j < m && A[j] != id) ++j;
        if (j == m) {
            A[m] = id;
            W[m] = 0.0f;
            ++m;
        }
        W[j] = W[j] + scores[i];
    }
    // filter (agg/mod.rs:39-44)
    uint32_t kept = 0;
    for (uint32_t j = 0; j < m; ++j)
        if (W[j] >= ap.lower_bound) {
            A[kept] = A[j];
            W[kept] = W[j];
            ++kept;
        }
    m = kept;
    if (m == 0) {
        out[rec] = 1u;  // taxa2agg.rs:174-175
        return;
    }
    // dense indices, ascending (insertion sort: records are short)
    for (uint32_t j = 0; j < m; ++j) {
        const uint32_t id = A[j];
        const uint32_t d = id <= tv.max_id ? __ldg(tv.dense_of + id) : kNoTaxon;
        if (d == kNoTaxon) {
            if (atomicCAS(&err[0], 0u, 1u) == 0u) err[1] = id;
            out[rec] = UMGAP_ABSENT;
            return;
        }
        const float w = W[j];
        uint32_t k = j;
        while (k > 0 && A[k - 1] > d) {
            A[k] = A[k - 1];
            W[k] = W[k - 1];
            --k;
        }
        A[k] = d;
        W[k] = w;
    }
    const ScoredRec r{tv, A, W};
    uint32_t result;
    if (ap.strategy == UMGAP_AGG_RMQ_HYBRID) {
        // rmq/mix.rs:56-93: every taxon of the closure of the members under pairwise LCA -- the members and the LCAs of
        // neighbours in preorder -- weighs  lca * factor + rtl * (1 - factor),  lca = the summed counts of the members at or
        // below it, rtl = those of the members at or above it; the heaviest wins (the first in preorder among equals;
        // the reference takes the last in HashMap order).  Sums run in ascending preorder.
        const float inv = __fsub_rn(1.0f, ap.factor);
        float best = 0.0f;
        result = kNoTaxon;
        for (uint32_t c = 0; c < 2 * m - 1; ++c) {
            const uint32_t t = (c & 1u) ? scored_lca(tv, A[c >> 1], A[(c >> 1) + 1]) : A[c >> 1];
            const uint32_t tl = __ldg(tv.last + t);
            float wl = 0.0f, wr = 0.0f;
            for (uint32_t i = 0; i < m; ++i) {
                if (A[i] >= t && A[i] <= tl) wl = wl + W[i];
                if (A[i] <= t && __ldg(tv.last + A[i]) >= t) wr = wr + W[i];
            }
            const float val = __fadd_rn(__fmul_rn(wl, ap.factor), __fmul_rn(wr, inv));
            if (result == kNoTaxon || val > best || (val == best && t < result)) {
                best = val;
                result = t;
            }
        }
    } else if (ap.strategy == UMGAP_AGG_MRTL) {
        float best = 0.0f;
        uint32_t best_j = 0;
        for (uint32_t j = 0; j < m; ++j) {
            float w = W[j];
            for (uint32_t i = j; i-- > 0;)  // the ancestors among the members, deepest first (rmq/rtl.rs:43-46)
                if (__ldg(tv.last + A[i]) >= A[j]) w = w + W[i];
            if (j == 0 || w > best) {
                best = w;
                best_j = j;
            }
        }
        result = A[best_j];
    } else {
        uint32_t lo2 = 0, hi2 = m, root, nk;
        float v;
        float bval = ap.strategy == UMGAP_AGG_HYBRID ? scored_subtree(r, 0, m) : 0.0f;
        scored_node(r, lo2, hi2, root, v, nk);  // LCA* = the root of the collapsed tree (tree/lca.rs:34-40)
        if (ap.strategy == UMGAP_AGG_HYBRID) {
            while (nk) {  // no children: stop (tree/mix.rs:51)
                const uint32_t cd = (uint32_t)__ldg(tv.depth + root) + 1;
                uint32_t bl = 0, bh = 0;
                float best = 0.0f;
                for (uint32_t j = lo2; j < hi2;) {  // the heaviest child, the first in order among equals
                    const uint32_t e = scored_group_end(r, j, hi2, cd);
                    const float sv = scored_subtree(r, j, e);
                    if (j == lo2 || sv > best) {
                        best = sv;
                        bl = j;
                        bh = e;
                    }
                    j = e;
                }
                if (__fdiv_rn(best, bval) < ap.factor) break;  // tree/mix.rs:57
                lo2 = bl;
                hi2 = bh;
                bval = best;
                scored_node(r, lo2, hi2, root, v, nk);
            }
        }
        result = root;
    }
    out[rec] = ap.ranked_only ? __ldg(tv.snap_ranked + result) : __ldg(tv.snap_valid + result);
}

}  // namespace umgap

using namespace umgap;

extern "C" {

uint64_t umgap_translate_bound(uint64_t total_nt, uint64_t nreads, uint8_t frames_mask) {
    (void)nreads;
    return ceil_div(total_nt, 3) * (uint64_t)__builtin_popcount(frames_mask & 0x3F) + 1;
}

int umgap_translate(int device, const uint8_t* nt, const uint64_t* read_off, uint64_t nreads,
                    int table, int methionine, uint8_t frames_mask, uint8_t* aa_out,
                    uint64_t* aa_off) {
    return guarded([&] {
        if (!read_off || !aa_off || (nreads && (!nt || !aa_out))) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        AsciiLut lut{};
        make_ascii_lut_public(table, methionine, lut.v);
        frames_mask &= 0x3F;
        const int nframes = __builtin_popcount(frames_mask);
        uint64_t o = 0;
        for (uint64_t r = 0; r < nreads; ++r) {
            const uint64_t n = read_off[r + 1] - read_off[r];
            int slot = 0;
            for (int fr = 0; fr < 6; ++fr) {
                if (!(frames_mask >> fr & 1)) continue;
                const uint64_t f = fr % 3;
                aa_off[r * nframes + slot++] = o;
                o += n >= f ? (n - f) / 3 : 0;
            }
        }
        aa_off[nreads * nframes] = o;
        if (!nreads || !nframes || !o) return;
        use_device(device);
        const uint64_t total_nt = read_off[nreads];
        DevBuf<uint8_t> d_nt(total_nt + 1), d_aa(o);
        DevBuf<uint64_t> d_roff(nreads + 1), d_aoff(nreads * nframes + 1);
        UMGAP_CUDA(cudaMemcpy(d_nt.p, nt, total_nt, cudaMemcpyHostToDevice));
        UMGAP_CUDA(cudaMemcpy(d_roff.p, read_off, (nreads + 1) * 8, cudaMemcpyHostToDevice));
        UMGAP_CUDA(cudaMemcpy(d_aoff.p, aa_off, (nreads * nframes + 1) * 8, cudaMemcpyHostToDevice));
        translate_kernel<<<grid_for(nreads, 8), 256>>>(lut, d_nt.p, d_roff.p, nreads, frames_mask,
                                                       nframes, d_aoff.p, d_aa.p);
        UMGAP_CUDA(cudaGetLastError());
        UMGAP_CUDA(cudaMemcpy(aa_out, d_aa.p, o, cudaMemcpyDeviceToHost));
    });
}

uint64_t umgap_kmer_lookup_bound(uint64_t total_aa, uint64_t npeps) {
    (void)npeps;
    return total_aa + 1;
}

int umgap_kmer_lookup(const umgap_index* idx, const uint8_t* aa, const uint64_t* pep_off,
                      uint64_t npeps, int one_on_one, uint32_t* taxa_out, uint64_t* taxa_off,
                      uint8_t* kept) {
    return guarded([&] {
        if (!idx || !pep_off || !taxa_off || (npeps && (!aa || !taxa_out)))
            UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        if (idx->k <= 0) UMGAP_FAIL(UMGAP_ERR_INVALID, "index is not a fixed-length k-mer table");
        const uint64_t k = idx->k;
        uint64_t o = 0;
        for (uint64_t p = 0; p < npeps; ++p) {
            const uint64_t len = pep_off[p + 1] - pep_off[p];
            taxa_off[p] = o;
            const bool keep = len >= k;  // prot2kmer2lca.rs:172
            if (kept) kept[p] = keep;
            if (keep) o += len - k + 1;
        }
        taxa_off[npeps] = o;
        if (!o) return;
        use_device(idx->device);
        const uint64_t total_aa = pep_off[npeps];
        DevBuf<uint8_t> d_aa(total_aa + 1), d_code(256);
        DevBuf<uint64_t> d_poff(npeps + 1), d_ooff(npeps + 1);
        DevBuf<uint32_t> d_out(o);
        UMGAP_CUDA(cudaMemcpy(d_aa.p, aa, total_aa, cudaMemcpyHostToDevice));
        UMGAP_CUDA(cudaMemcpy(d_code.p, idx->code_of_byte, 256, cudaMemcpyHostToDevice));
        UMGAP_CUDA(cudaMemcpy(d_poff.p, pep_off, (npeps + 1) * 8, cudaMemcpyHostToDevice));
        UMGAP_CUDA(cudaMemcpy(d_ooff.p, taxa_off, (npeps + 1) * 8, cudaMemcpyHostToDevice));
        if (idx->nshards > 1 && !idx->attached)
            UMGAP_FAIL(UMGAP_ERR_INVALID, "sharded index: call umgap_index_attach_shards() before looking up");
        if (idx->nshards > 1)
            kmer_lookup_kernel<ShardedView><<<grid_for(npeps, 8), 256>>>(idx->sharded, d_code.p, d_aa.p, d_poff.p, npeps,
                                                                         d_ooff.p, d_out.p);
        else
            kmer_lookup_kernel<TableView><<<grid_for(npeps, 8), 256>>>(idx->view(), d_code.p, d_aa.p, d_poff.p, npeps,
                                                                       d_ooff.p, d_out.p);
        UMGAP_CUDA(cudaGetLastError());
        UMGAP_CUDA(cudaMemcpy(taxa_out, d_out.p, o * 4, cudaMemcpyDeviceToHost));
        if (one_on_one) {  // miss -> 0 (prot2kmer2lca.rs:115)
            for (uint64_t i = 0; i < o; ++i)
                if (taxa_out[i] == UMGAP_MISS) taxa_out[i] = 0;
        } else {  // miss -> omitted (:176)
            uint64_t w = 0;
            for (uint64_t p = 0; p < npeps; ++p) {
                const uint64_t b = taxa_off[p], e = taxa_off[p + 1];
                taxa_off[p] = w;
                for (uint64_t i = b; i < e; ++i)
                    if (taxa_out[i] != UMGAP_MISS) taxa_out[w++] = taxa_out[i];
            }
            taxa_off[npeps] = w;
        }
    });
}

int umgap_seedextend(int device, const uint32_t* taxa, const uint64_t* rec_off, uint64_t nrecs,
                     int min_seed_size, int max_gap_size, uint32_t* out, uint64_t* out_off) {
    return guarded([&] {
        if (!rec_off || !out_off) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        if (min_seed_size < 0 || max_gap_size < 0) UMGAP_FAIL(UMGAP_ERR_INVALID, "negative size");
        const uint64_t total = rec_off[nrecs];
        for (uint64_t i = 0; i <= nrecs; ++i) out_off[i] = 0;
        if (!nrecs) return;
        if (total && (!taxa || !out)) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        use_device(device);
        DevBuf<uint32_t> d_in(total + 1), d_out(total + 1), d_len(nrecs);
        DevBuf<uint64_t> d_off(nrecs + 1);
        if (total) UMGAP_CUDA(cudaMemcpy(d_in.p, taxa, total * 4, cudaMemcpyHostToDevice));
        UMGAP_CUDA(cudaMemcpy(d_off.p, rec_off, (nrecs + 1) * 8, cudaMemcpyHostToDevice));
        seedextend_kernel<<<grid_for(nrecs, 128), 128>>>(d_in.p, d_off.p, nrecs, (uint32_t)min_seed_size,
                                                         (uint32_t)max_gap_size, d_out.p, d_len.p);
        UMGAP_CUDA(cudaGetLastError());
        std::vector<uint32_t> len(nrecs), tmp(total + 1);
        UMGAP_CUDA(cudaMemcpy(len.data(), d_len.p, nrecs * 4, cudaMemcpyDeviceToHost));
        if (total) UMGAP_CUDA(cudaMemcpy(tmp.data(), d_out.p, total * 4, cudaMemcpyDeviceToHost));
        uint64_t w = 0;
        for (uint64_t r = 0; r < nrecs; ++r) {
            out_off[r] = w;
            memcpy(out + w, tmp.data() + rec_off[r], (size_t)len[r] * 4);
            w += len[r];
        }
        out_off[nrecs] = w;
    });
}

int umgap_seedextend_ranked(const umgap_taxonomy* tax, const uint32_t* taxa, const uint64_t* rec_off, uint64_t nrecs,
                            int min_seed_size, int max_gap_size, int penalty, uint32_t* out, uint64_t* out_off) {
    return guarded([&] {
        if (!tax || !rec_off || !out_off) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        if (min_seed_size < 0 || max_gap_size < 0 || penalty < 0) UMGAP_FAIL(UMGAP_ERR_INVALID, "negative size");
        const uint64_t total = rec_off[nrecs];
        for (uint64_t i = 0; i <= nrecs; ++i) out_off[i] = 0;
        if (!nrecs) return;
        if (total && (!taxa || !out)) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        use_device(tax->device);
        DevBuf<uint32_t> d_in(total + 1), d_out(total + 1), d_len(nrecs);
        DevBuf<uint64_t> d_off(nrecs + 1);
        if (total) UMGAP_CUDA(cudaMemcpy(d_in.p, taxa, total * 4, cudaMemcpyHostToDevice));
        UMGAP_CUDA(cudaMemcpy(d_off.p, rec_off, (nrecs + 1) * 8, cudaMemcpyHostToDevice));
        seedextend_ranked_kernel<<<grid_for(nrecs, 128), 128>>>(tax->view, d_in.p, d_off.p, nrecs, (uint32_t)min_seed_size,
                                                                (uint32_t)max_gap_size, (uint32_t)penalty, d_out.p, d_len.p);
        UMGAP_CUDA(cudaGetLastError());
        std::vector<uint32_t> len(nrecs), tmp(total + 1);
        UMGAP_CUDA(cudaMemcpy(len.data(), d_len.p, nrecs * 4, cudaMemcpyDeviceToHost));
        if (total) UMGAP_CUDA(cudaMemcpy(tmp.data(), d_out.p, total * 4, cudaMemcpyDeviceToHost));
        uint64_t w = 0;
        for (uint64_t r = 0; r < nrecs; ++r) {
            out_off[r] = w;
            memcpy(out + w, tmp.data() + rec_off[r], (size_t)len[r] * 4);
            w += len[r];
        }
        out_off[nrecs] = w;
    });
}

int umgap_aggregate(const umgap_taxonomy* tax, const uint32_t* taxa, const uint64_t* rec_off,
                    uint64_t nrecs, int strategy, float factor, float lower_bound, int ranked_only,
                    uint32_t* taxon_out) {
    return guarded([&] {
        if (!tax || !rec_off || (nrecs && !taxon_out)) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        if (strategy < UMGAP_AGG_LCA_STAR || strategy > UMGAP_AGG_MRTL)
            UMGAP_FAIL(UMGAP_ERR_INVALID, "unknown aggregation strategy %d", strategy);
        if (!nrecs) return;
        const uint64_t total = rec_off[nrecs];
        if (total && !taxa) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        use_device(tax->device);
        DevBuf<uint32_t> d_in(total + 1), d_scratch(3 * total + nrecs + 8), d_out(nrecs);
        DevBuf<uint64_t> d_off(nrecs + 1);
        DevBuf<unsigned int> d_err(2);
        if (total) UMGAP_CUDA(cudaMemcpy(d_in.p, taxa, total * 4, cudaMemcpyHostToDevice));
        UMGAP_CUDA(cudaMemcpy(d_off.p, rec_off, (nrecs + 1) * 8, cudaMemcpyHostToDevice));
        UMGAP_CUDA(cudaMemset(d_err.p, 0, 8));
        AggParams ap{strategy, factor, lower_bound, ranked_only};
        aggregate_kernel<<<grid_for(nrecs, kStageAggWarps), kStageAggWarps * 32>>>(
            tax->view, ap, d_in.p, d_off.p, nrecs, d_scratch.p, d_out.p, d_err.p);
        UMGAP_CUDA(cudaGetLastError());
        unsigned int he[2];
        UMGAP_CUDA(cudaMemcpy(he, d_err.p, 8, cudaMemcpyDeviceToHost));
        UMGAP_CUDA(cudaMemcpy(taxon_out, d_out.p, nrecs * 4, cudaMemcpyDeviceToHost));
        if (he[0]) UMGAP_FAIL(UMGAP_ERR_UNKNOWN_TAXON, "Unknown Taxon ID: %u", he[1]);
    });
}

int umgap_aggregate_scored(const umgap_taxonomy* tax, const uint32_t* taxa, const float* scores, const uint64_t* rec_off,
                           uint64_t nrecs, int strategy, float factor, float lower_bound, int ranked_only, uint32_t* taxon_out) {
    return guarded([&] {
        if (!tax || !rec_off || (nrecs && !taxon_out)) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        if (strategy < UMGAP_AGG_LCA_STAR || strategy > UMGAP_AGG_RMQ_HYBRID)
            UMGAP_FAIL(UMGAP_ERR_INVALID, "unknown aggregation strategy %d", strategy);
        if (!nrecs) return;
        const uint64_t total = rec_off[nrecs];
        if (total && (!taxa || !scores)) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        use_device(tax->device);
        DevBuf<uint32_t> d_in(total + 1), d_a(total + 1), d_out(nrecs);
        DevBuf<float> d_sc(total + 1), d_w(total + 1);
        DevBuf<uint64_t> d_off(nrecs + 1);
        DevBuf<unsigned int> d_err(2);
        if (total) {
            UMGAP_CUDA(cudaMemcpy(d_in.p, taxa, total * 4, cudaMemcpyHostToDevice));
            UMGAP_CUDA(cudaMemcpy(d_sc.p, scores, total * 4, cudaMemcpyHostToDevice));
        }
        UMGAP_CUDA(cudaMemcpy(d_off.p, rec_off, (nrecs + 1) * 8, cudaMemcpyHostToDevice));
        UMGAP_CUDA(cudaMemset(d_err.p, 0, 8));
        AggParams ap{strategy, factor, lower_bound, ranked_only};
        aggregate_scored_kernel<<<(unsigned)ceil_div(nrecs, 64), 64>>>(tax->view, ap, d_in.p, d_sc.p, d_off.p, nrecs, d_a.p, d_w.p, d_out.p,
                                                                       d_err.p);
        UMGAP_CUDA(cudaGetLastError());
        unsigned int he[2];
        UMGAP_CUDA(cudaMemcpy(he, d_err.p, 8, cudaMemcpyDeviceToHost));
        UMGAP_CUDA(cudaMemcpy(taxon_out, d_out.p, nrecs * 4, cudaMemcpyDeviceToHost));
        if (he[0]) UMGAP_FAIL(UMGAP_ERR_UNKNOWN_TAXON, "Unknown Taxon ID: %u", he[1]);
    });
}

}  // extern "C"
