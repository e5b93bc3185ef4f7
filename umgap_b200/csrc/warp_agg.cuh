// Warp-cooperative per-record machinery: the seedextend state machine (one lane per record) and
// the taxa2agg aggregators LCA*, hybrid and MRTL (one warp per record) on the preorder-numbered
// taxonomy.  Restates seedextend.rs:94-149,167-176, agg/mod.rs:27-44, tree/mod.rs:29-101,
// tree/lca.rs:34-40, tree/mix.rs:43-64, rmq/rtl.rs:39-57 and taxa2agg.rs:159-181 in closed form:
//
//   S      = distinct taxa of the record with count >= lower bound, sorted by preorder index
//   leaf   = member of S with no strict descendant in S  <=>  next member lies outside its subtree
//   LCA*   = LCA(first leaf, last member)            (root of the collapsed induced tree)
//   hybrid = from LCA*, descend into the child subtree holding the largest count while
//            count/parent_count >= factor (IEEE f32 division), re-collapsing at every step
//   MRTL   = member maximising the summed counts of its ancestors-or-self in S
#pragma once
#include "taxdev.cuh"

namespace umgap {

constexpr uint32_t kSkip = 0xFFFFFFFEu;  // internal: element removed from a stream

// ---- seedextend -----------------------------------------------------------------------------
// Streams the logical list ids[0], ids[stride], ... (count raw entries).  A raw UMGAP_MISS is a
// 0 when one_on_one, otherwise it is not part of the list at all (prot2kmer2lca.rs:115,176).
// Calls emit(v) for every element of every selected range, in order.
template <class Emit>
__device__ __forceinline__ void seedextend_stream(const uint32_t* ids, uint32_t stride,
                                                  uint32_t count, bool one_on_one,
                                                  uint32_t min_seed, uint32_t max_gap, Emit emit) {
    auto fetch = [&](uint32_t raw) -> uint32_t {
        const uint32_t v = ids[(uint64_t)raw * stride];
        return v == kNoValue ? (one_on_one ? 0u : kSkip) : v;
    };
    auto flush = [&](uint32_t raw_from, uint32_t n) {  // emit n list elements starting at raw_from
        for (uint32_t r = raw_from; n; ++r) {
            const uint32_t v = fetch(r);
            if (v == kSkip) continue;
            emit(v);
            --n;
        }
    };
    // first element t[0] (the sentinel 0 when the list is empty)
    uint32_t raw = 0;
    uint32_t last = 0;
    bool have_first = false;
    for (; raw < count; ++raw) {
        const uint32_t v = fetch(raw);
        if (v != kSkip) {
            last = v;
            have_first = true;
            break;
        }
    }
    if (!have_first) return;  // t = [0]: every candidate range is empty
    uint32_t start = 0, end = 1, same = 1, smax = 1;
    uint32_t start_raw = raw;  // raw index at or before list element `start`
    ++raw;
    for (;;) {
        // next list element t[end]; the sentinel once the raw entries are exhausted
        uint32_t cur = 0;
        bool sentinel = true;
        for (; raw < count; ++raw) {
            const uint32_t v = fetch(raw);
            if (v != kSkip) {
                cur = v;
                sentinel = false;
                break;
            }
        }
        const uint32_t cur_raw = raw;  // raw position of t[end] (== count for the sentinel)
        if (last == cur) {                                   // :109-113
            ++same;
        } else if (last == 0 && same > max_gap) {            // :116-127 gap too long
            // the reference slices start..end-same here; after a leading gap (:130-134) start can lie
            // beyond it (min_seed <= 1, max_gap >= 1), where the reference panics: an empty selection
            const uint32_t e = end - same;
            if (smax >= min_seed && e > start) flush(start_raw, e - start);
            start = end;
            start_raw = cur_raw;
            last = cur;
            same = 1;
            smax = 1;
        } else if (last == 0 && end - start == same) {       // :130-134 do not start with a gap
            start = end + 1;
            start_raw = cur_raw + 1;
        } else {
            if (last != 0) smax = smax > same ? smax : same; // :137-139
            last = cur;                                      // :140-142
            same = 1;
        }
        ++end;
        if (sentinel) break;
        ++raw;
    }
    // :144-149 -- `last` is 0 here (the sentinel was consumed), so the trailing zeros drop out
    if (smax >= min_seed) {
        if (last == 0) end -= same;
        if (end > start) flush(start_raw, end - start);
    }
}

// ---- warp utilities ----------------------------------------------------------------------------
template <bool KV>
__device__ __forceinline__ void cmpex(uint32_t* a, uint32_t* c, uint32_t i, uint32_t l) {
    const uint32_t x = a[i], y = a[l];
    if (x > y) {
        a[i] = y;
        a[l] = x;
        if (KV) {
            const uint32_t t = c[i];
            c[i] = c[l];
            c[l] = t;
        }
    }
}

// Ascending bitonic sort of a[0..n) for any n (indices >= n behave as +infinity); with KV the
// payload c[] moves with its key.
template <bool KV>
__device__ __forceinline__ void warp_sort(uint32_t* a, uint32_t* c, uint32_t n, int lane) {
    if (n < 2) return;
    uint32_t np2 = 1;
    while (np2 < n) np2 <<= 1;
    const uint32_t pairs = np2 >> 1;
    for (uint32_t k = 2; k <= np2; k <<= 1) {
        const uint32_t hk = k >> 1;
        for (uint32_t t = lane; t < pairs; t += 32) {
            const uint32_t blk = t / hk, r = t - blk * hk;
            const uint32_t i = blk * k + r, l = blk * k + (k - 1 - r);
            if (l < n) cmpex<KV>(a, c, i, l);
        }
        __syncwarp();
        for (uint32_t j = k >> 2; j > 0; j >>= 1) {
            for (uint32_t t = lane; t < pairs; t += 32) {
                const uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i | j;
                if (l < n) cmpex<KV>(a, c, i, l);
            }
            __syncwarp();
        }
    }
}

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
    }
    return v;
}

// First index in [lo, hi) with A[idx] > bound (A ascending); hi when none.  All lanes call and
// all receive the result.
__device__ __forceinline__ uint32_t warp_upper_bound(const uint32_t* A, uint32_t lo, uint32_t hi,
                                                     uint32_t bound, int lane) {
    for (uint32_t base = lo; base < hi; base += 32) {
        const uint32_t i = base + lane;
        const bool gt = i < hi && A[i] > bound;
        const unsigned m = __ballot_sync(0xffffffffu, gt);
        if (m) return base + (uint32_t)__ffs(m) - 1;
    }
    return hi;
}

// LCA* of the members [lo, hi) of the sorted distinct list (A = dense index, L = last[A]).
__device__ __forceinline__ uint32_t lca_star_range(const TaxView& tv, const uint32_t* A,
                                                   const uint32_t* L, uint32_t lo, uint32_t hi,
                                                   int lane) {
    uint32_t first_leaf = hi - 1;
    for (uint32_t base = lo; base < hi; base += 32) {
        const uint32_t j = base + lane;
        const bool leaf = j < hi && (j + 1 == hi || A[j + 1] > L[j]);
        const unsigned m = __ballot_sync(0xffffffffu, leaf);
        if (m) {
            first_leaf = base + (uint32_t)__ffs(m) - 1;
            break;
        }
    }
    return warp_lca(tv, A[first_leaf], A[hi - 1], lane);
}

struct AggParams {
    int strategy;
    float factor;
    float lower_bound;
    int ranked_only;
};

constexpr uint32_t kAggUnknown = 0xFFFFFFFDu;  // internal: record holds an id unknown to the tree

// The strategies proper, on the sorted distinct kept members: A[0..m) dense indices ascending, P[0..m]
// exclusive prefix sums of their counts (P[m] = running = total), L[j] = last[A[j]].  All lanes call.
__device__ __forceinline__ uint32_t agg_finish(const TaxView& tv, const uint32_t* A, const uint32_t* P, const uint32_t* L,
                                               uint32_t m, uint32_t running, const AggParams& ap, int lane) {
    uint32_t result;
    if (ap.strategy == UMGAP_AGG_LCA_STAR) {
        result = lca_star_range(tv, A, L, 0, m, lane);
    } else if (ap.strategy == UMGAP_AGG_HYBRID) {
        uint32_t lo = 0, hi = m;
        uint32_t base_node = lca_star_range(tv, A, L, lo, hi, lane);
        uint32_t bval = running;
        lo = warp_upper_bound(A, lo, hi, base_node - 1u, lane);  // first member >= base (base>=0)
        if (base_node == 0) lo = 0;
        for (;;) {
            uint32_t g = lo;
            if (g < hi && A[g] == base_node) ++g;  // the base itself is not one of its children
            if (g >= hi) break;                    // no children: stop (tree/mix.rs:51)
            const uint32_t child_depth = (uint32_t)__ldg(tv.depth + base_node) + 1;
            uint32_t best = 0, best_lo = 0, best_hi = 0;
            while (g < hi) {
                const uint32_t c = __ldg(tv.anc + (uint64_t)A[g] * tv.stride + child_depth);
                const uint32_t e = warp_upper_bound(A, g, hi, __ldg(tv.last + c), lane);
                const uint32_t sub = P[e] - P[g];
                if (sub > best) {  // first maximal child in preorder wins ties
                    best = sub;
                    best_lo = g;
                    best_hi = e;
                }
                g = e;
            }
            if (__fdiv_rn((float)best, (float)bval) < ap.factor) break;  // tree/mix.rs:57
            base_node = lca_star_range(tv, A, L, best_lo, best_hi, lane);
            bval = best;
            hi = best_hi;
            lo = base_node == 0 ? best_lo : warp_upper_bound(A, best_lo, best_hi, base_node - 1u, lane);
        }
        result = base_node;
    } else {  // MRTL
        uint32_t best_w = 0, best_j = 0;
        for (uint32_t base = 0; base < m; base += 32) {
            const uint32_t j = base + lane;
            uint32_t w = 0;
            if (j < m) {
                const uint32_t me = A[j];
                for (uint32_t i = 0; i <= j; ++i)
                    if (L[i] >= me) w += P[i + 1] - P[i];
            }
            // arg-max within the chunk, first index wins ties
            uint32_t bw = w, bj = j;
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                const uint32_t ow = __shfl_xor_sync(0xffffffffu, bw, o);
                const uint32_t oj = __shfl_xor_sync(0xffffffffu, bj, o);
                if (ow > bw || (ow == bw && oj < bj)) {
                    bw = ow;
                    bj = oj;
                }
            }
            if (bw > best_w) {
                best_w = bw;
                best_j = bj;
            }
        }
        result = A[best_j];
    }
    return ap.ranked_only ? __ldg(tv.snap_ranked + result) : __ldg(tv.snap_valid + result);
}

// Aggregates one record.  A[0..n) holds its non-zero taxon ids (any order) and, with KV, C[0..n)
// how many times each entry occurred (run-length compressed input); P needs n+1 and L needs n
// entries of scratch.  Returns the snapped taxon id, the literal 1 for an empty record
// (taxa2agg.rs:174-175), or kAggUnknown with *bad_id set.  All 32 lanes must call.
template <bool KV>
__device__ __forceinline__ uint32_t warp_aggregate(const TaxView& tv, uint32_t* A, uint32_t* C, uint32_t* P,
                                                   uint32_t* L, uint32_t n, const AggParams& ap,
                                                   int lane, uint32_t* bad_id) {
    if (n == 0) return 1u;
    // 1. taxon id -> preorder index.  An id the tree does not hold raises only if it survives the lower bound -- the
    //    reference filters the counts before the aggregator sees them (taxa2agg.rs:169-170) -- so it is counted like
    //    any other, under a marked value that sorts behind every index (ids that differ only in bit 31 would share it)
    constexpr uint32_t kUnknownMark = 0x80000000u;
    for (uint32_t i = lane; i < n; i += 32) {
        const uint32_t id = A[i];
        const uint32_t d = id <= tv.max_id ? __ldg(tv.dense_of + id) : kNoTaxon;
        A[i] = d == kNoTaxon ? (kUnknownMark | id) : d;
    }
    __syncwarp();
    // 2. sort, 3. distinct + run starts
    warp_sort<KV>(A, C, n, lane);
    uint32_t occurrences = n;
    if (KV) {  // C <- exclusive prefix sums of the occurrence counts (run start "positions")
        uint32_t run = 0;
        for (uint32_t base = 0; base < n; base += 32) {
            const uint32_t i = base + lane;
            const uint32_t c = i < n ? C[i] : 0u;
            const uint32_t incl = warp_incl_scan(c, lane);
            if (i < n) C[i] = run + incl - c;
            run += __shfl_sync(0xffffffffu, incl, 31);
        }
        occurrences = run;
        __syncwarp();
    }
    uint32_t m = 0;
    uint32_t carry = kNoTaxon;
    for (uint32_t base = 0; base < n; base += 32) {
        const uint32_t i = base + lane;
        const uint32_t v = i < n ? A[i] : kNoTaxon;
        uint32_t prev = __shfl_up_sync(0xffffffffu, v, 1);
        if (lane == 0) prev = carry;
        const bool head = i < n && (i == 0 || v != prev);
        const unsigned mask = __ballot_sync(0xffffffffu, head);
        const uint32_t rank = __popc(mask & ((1u << lane) - 1));
        carry = __shfl_sync(0xffffffffu, v, 31);
        __syncwarp();
        const uint32_t startpos = KV ? (i < n ? C[i] : 0u) : i;
        __syncwarp();
        if (head) {
            A[m + rank] = v;
            P[m + rank] = startpos;
        }
        m += __popc(mask);
        __syncwarp();
    }
    if (lane == 0) P[m] = occurrences;
    __syncwarp();
    // 4. counts, lower-bound filter (agg/mod.rs:39-44: keep count >= lower_bound, f32 compare),
    //    exclusive prefix sums of the kept counts into P
    uint32_t kept = 0, running = 0;
    for (uint32_t base = 0; base < m; base += 32) {
        const uint32_t j = base + lane;
        uint32_t c = 0, v = 0;
        if (j < m) {
            c = P[j + 1] - P[j];
            v = A[j];
        }
        const bool keep = j < m && (float)c >= ap.lower_bound;
        const unsigned mask = __ballot_sync(0xffffffffu, keep);
        const uint32_t rank = __popc(mask & ((1u << lane) - 1));
        const uint32_t incl = warp_incl_scan(keep ? c : 0u, lane);
        __syncwarp();
        if (keep) {
            A[kept + rank] = v;
            P[kept + rank] = running + incl - c;
        }
        kept += __popc(mask);
        running += __shfl_sync(0xffffffffu, incl, 31);
        __syncwarp();
    }
    m = kept;
    if (m == 0) return 1u;  // everything filtered: the literal "1"
    {
        const uint32_t tail = A[m - 1];  // marked values sort last
        if (tail & kUnknownMark) {
            *bad_id = tail & ~kUnknownMark;
            return kAggUnknown;
        }
    }
    if (lane == 0) P[m] = running;
    for (uint32_t j = lane; j < m; j += 32) L[j] = __ldg(tv.last + A[j]);
    __syncwarp();
    return agg_finish(tv, A, P, L, m, running, ap, lane);
}

// LCA of dense x <= y: the deepest depth at which the two rows of the ancestor matrix agree (rows are padded
// with kNoTaxon past the node's own depth); one round of two independent loads per 32 depths.  Also returns
// that depth (undefined when x == y).  All lanes must call.
__device__ __forceinline__ uint32_t warp_lca_rows(const TaxView& t, uint32_t x, uint32_t y, int lane, uint32_t& depth) {
    depth = 0;
    if (x == y) return x;
    uint32_t best = 0;  // the root (dense 0, depth 0) is on every path
    for (uint32_t base = 0; base < t.stride; base += 32) {
        const uint32_t d = base + lane;
        uint32_t ax = kNoTaxon, ay = kNoTaxon;
        if (d < t.stride) {
            ax = __ldg(t.anc + (uint64_t)x * t.stride + d);
            ay = __ldg(t.anc + (uint64_t)y * t.stride + d);
        }
        const unsigned same = __ballot_sync(0xffffffffu, ax == ay && ax != kNoTaxon);
        if (same) {
            const int top = 31 - __clz(same);
            best = __shfl_sync(0xffffffffu, ax, top);
            depth = base + (uint32_t)top;
        }
        if (same != 0xffffffffu) break;  // the paths agree on a prefix of depths only
    }
    return best;
}

// The strategies LCA* and hybrid on m <= 32 sorted distinct kept members held one per lane (lane j < m: dense
// index a, subtree end l = last[a], p0 = exclusive prefix of the counts; lanes >= m: a = kNoTaxon, p0 = running).
// Same results as agg_finish; a level of the hybrid descent is one parallel step: every member loads its ancestor
// one below the base, equal neighbours form the child ranges, their sums come from the prefix, a warp arg-max picks
// the first heaviest child.  All lanes must call.
__device__ __forceinline__ uint32_t agg_finish32(const TaxView& tv, uint32_t a, uint32_t l, uint32_t p0, uint32_t m,
                                                 uint32_t running, const AggParams& ap, int lane) {
    const unsigned all = m >= 32 ? 0xffffffffu : (1u << m) - 1;
    const uint32_t a_next = __shfl_down_sync(0xffffffffu, a, 1);
    // member j is a leaf of the induced tree iff the next member lies outside its subtree; the last member of any
    // child range is one (what follows it belongs to another child)
    const unsigned leaves = __ballot_sync(0xffffffffu, (uint32_t)lane < m && ((uint32_t)lane + 1 == m || a_next > l));
    uint32_t base_depth = 0;
    auto lca_star = [&](unsigned range) -> uint32_t {  // LCA(first leaf, last member) of a contiguous member range
        const int first_leaf = __ffs((int)(leaves & range)) - 1, last = 31 - __clz(range);
        return warp_lca_rows(tv, __shfl_sync(0xffffffffu, a, first_leaf), __shfl_sync(0xffffffffu, a, last), lane, base_depth);
    };
    uint32_t base_node = lca_star(all);
    if (ap.strategy == UMGAP_AGG_HYBRID) {
        unsigned range = all;   // members inside the subtree of the current base's range
        uint32_t hi = m, bval = running;
        for (;;) {
            // the members below the base: inside the range, not an ancestor of the base, not the base itself
            const bool below = (range >> lane & 1u) && a > base_node;
            const unsigned cmask = __ballot_sync(0xffffffffu, below);
            if (!cmask) break;  // no children: stop (tree/mix.rs:51)
            const uint32_t ch = below ? __ldg(tv.anc + (uint64_t)a * tv.stride + base_depth + 1) : kNoTaxon;
            const uint32_t ch_prev = __shfl_up_sync(0xffffffffu, ch, 1);
            const bool head = below && (lane == __ffs((int)cmask) - 1 || ch != ch_prev);
            const unsigned heads = __ballot_sync(0xffffffffu, head);
            const unsigned above = lane == 31 ? 0u : heads & ~((2u << lane) - 1);
            const uint32_t end = above ? (uint32_t)__ffs((int)above) - 1 : hi;  // one past the child's last member
            const uint32_t p_end = __shfl_sync(0xffffffffu, p0, end & 31);
            const uint32_t sub = head ? (end < 32 ? p_end : running) - p0 : 0u;
            const uint32_t best = __reduce_max_sync(0xffffffffu, sub);
            if (__fdiv_rn((float)best, (float)bval) < ap.factor) break;  // tree/mix.rs:57
            const int bl = __ffs((int)__ballot_sync(0xffffffffu, head && sub == best)) - 1;  // first maximal child in preorder
            const uint32_t best_hi = __shfl_sync(0xffffffffu, end, bl);
            range = (best_hi >= 32 ? 0xffffffffu : (1u << best_hi) - 1) & ~((1u << bl) - 1);
            base_node = lca_star(range);
            bval = best;
            hi = best_hi;
        }
    }
    return ap.ranked_only ? __ldg(tv.snap_ranked + base_node) : __ldg(tv.snap_valid + base_node);
}

// Aggregates a record given as n <= 32 DISTINCT non-zero taxon ids with their occurrence counts, one per
// lane (lane i < n holds id / cnt): filter, rank sort and prefix sums in registers.  P needs n+1 and
// A, L need n entries of scratch.  Same results as warp_aggregate.  All 32 lanes must call.
__device__ __forceinline__ uint32_t warp_aggregate_distinct(const TaxView& tv, uint32_t id, uint32_t cnt, uint32_t n,
                                                            uint32_t* A, uint32_t* P, uint32_t* L, const AggParams& ap,
                                                            int lane, uint32_t* bad_id) {
    if (n == 0) return 1u;
    const bool have = (uint32_t)lane < n;
    uint32_t d = kNoTaxon;
    if (have) d = id <= tv.max_id ? __ldg(tv.dense_of + id) : kNoTaxon;
    // lower-bound filter first (agg/mod.rs:39-44, f32 compare), then a rank sort of the kept members: the dozen
    // members of a typical record take a dozen shuffles, a sorting network over all 32 lanes takes 15 steps of two
    const bool keep = have && (float)cnt >= ap.lower_bound;
    // an id the tree does not hold raises only if it survives the filter (taxa2agg.rs:169-170 filters first)
    if (keep && d == kNoTaxon) *bad_id = id;
    if (__any_sync(0xffffffffu, keep && d == kNoTaxon)) return kAggUnknown;
    const unsigned mask = __ballot_sync(0xffffffffu, keep);
    const uint32_t m = (uint32_t)__popc(mask);
    if (m == 0) return 1u;  // everything filtered: the literal "1"
    uint32_t rank = 0;  // members are distinct: the number of smaller kept members is the sorted position
    for (unsigned rest = mask; rest; rest &= rest - 1) {
        const uint32_t dj = __shfl_sync(0xffffffffu, d, __ffs((int)rest) - 1);
        rank += dj < d ? 1u : 0u;
    }
    __syncwarp();
    if (keep) {
        A[rank] = d;
        L[rank] = cnt;  // the counts travel through L
    }
    __syncwarp();
    // member j to lane j: dense index, subtree end, exclusive prefix of the counts (running beyond the members)
    const bool mem = (uint32_t)lane < m;
    const uint32_t a = mem ? A[lane] : kNoTaxon, c = mem ? L[lane] : 0u;
    const uint32_t incl = warp_incl_scan(c, lane);
    const uint32_t running = __shfl_sync(0xffffffffu, incl, 31);
    const uint32_t p0 = mem ? incl - c : running, l = mem ? __ldg(tv.last + a) : 0u;
    if (ap.strategy == UMGAP_AGG_MRTL) {
        __syncwarp();
        if (mem) {
            P[lane] = p0;
            L[lane] = l;
        }
        if (lane == 0) P[m] = running;
        __syncwarp();
        return agg_finish(tv, A, P, L, m, running, ap, lane);
    }
    return agg_finish32(tv, a, l, p0, m, running, ap, lane);
}

}  // namespace umgap
