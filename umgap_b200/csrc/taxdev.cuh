// Device helpers over the preorder-numbered taxonomy (TaxView, index.h).
#pragma once
#include "index.h"

namespace umgap {

// LCA of two dense indices by one thread: climb from the preorder-smaller node until its
// subtree interval covers the other one.
__device__ __forceinline__ uint32_t lca_dense(const TaxView& t, uint32_t a, uint32_t b) {
    if (a > b) {
        const uint32_t x = a;
        a = b;
        b = x;
    }
    while (t.last[a] < b) a = t.parent[a];
    return a;
}

// LCA of two taxon ids (used when the synthetic index merges duplicate k-mers); ids unknown to
// the taxonomy collapse to the root.
__device__ __forceinline__ uint32_t lca_ids(const TaxView& t, uint32_t ia, uint32_t ib) {
    const uint32_t a = ia <= t.max_id ? t.dense_of[ia] : kNoTaxon;
    const uint32_t b = ib <= t.max_id ? t.dense_of[ib] : kNoTaxon;
    if (a == kNoTaxon || b == kNoTaxon) return t.id_of[0];
    return t.id_of[lca_dense(t, a, b)];
}

// Warp-cooperative LCA of dense a <= b: every lane inspects one depth of a's root path through
// the ancestor matrix; the deepest ancestor whose interval covers b wins.  All lanes must call.
__device__ __forceinline__ uint32_t warp_lca(const TaxView& t, uint32_t a, uint32_t b, int lane) {
    if (a == b) return a;
    const int da = __ldg(t.depth + a);
    uint32_t best = 0;  // the root (dense 0) covers everything
    for (int base = 0; base <= da; base += 32) {
        const int d = base + lane;
        uint32_t x = 0;
        bool covers = false;
        if (d <= da) {
            x = __ldg(t.anc + (uint64_t)a * t.stride + d);
            covers = __ldg(t.last + x) >= b;
        }
        const unsigned m = __ballot_sync(0xffffffffu, covers);
        if (m) {
            const int top = 31 - __clz(m);
            best = __shfl_sync(0xffffffffu, x, top);
        }
        if (m != 0xffffffffu) break;  // covers() is monotone in depth: first false ends the search
    }
    return best;
}

}  // namespace umgap
