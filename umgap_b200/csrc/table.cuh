// GPU-resident k-mer -> taxon table (replaces fst::Map::get on the hot path; call sites
// prot2kmer2lca.rs:176, pept2lca.rs:93).
//
// Layout in HBM.  A table is up to kMaxLevels *levels*; a level is an array of 32-byte buckets
// (one DRAM sector, 32-byte aligned) of four 8-byte slots:
//
//      slot = value:32 | flag:1 | disp:3 | tag:28          EMPTY = all ones
//
// A key is 5 bits per residue (per-index alphabet of <= 32 byte values), first residue most
// significant, at most 9 residues = 45 bits.  h = mix45(key) is a bijection on 45 bits; the
// home bucket is floor(h * nb / 2^45) and the tag is the low 28 bits of h.  All h that share a
// home bucket lie in an interval narrower than 2^28 (nb >= 2^17), so (home, tag) identifies the
// key exactly -- no false positives, no key bytes stored.  A key that does not fit in its home
// bucket goes to one of the next kMaxDisp-1 buckets, recording the distance in `disp` (which
// keeps (bucket, disp, tag) exact) and setting the overflow `flag` (bit 31 of slot 0) of every
// full bucket it passed.  A key that finds kMaxDisp full buckets goes to the next level, a
// smaller table of the same shape.  A probe therefore reads ONE sector when the home bucket
// holds the key or is not flagged -- a miss costs the same as a hit -- and continues only
// through flagged buckets.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace umgap {

constexpr int kMaxLevels = 4;
constexpr int kMaxDisp = 8;             // probe distances 0..7 fit the 3-bit disp field
constexpr int kKeyBits = 45;
constexpr uint64_t kKeyMask = (1ull << kKeyBits) - 1;
constexpr uint64_t kMinBuckets = 1ull << 17;
constexpr uint32_t kTagMask = (1u << 28) - 1;
constexpr uint32_t kFlagBit = 1u << 31;
constexpr uint64_t kEmptySlot = ~0ull;
constexpr uint32_t kNoValue = 0xFFFFFFFFu;  // also UMGAP_MISS
constexpr uint64_t kInvalidKey = ~0ull;     // k-mer holding a byte outside the index alphabet

struct TableView {
    const ulonglong4* level[kMaxLevels];
    uint64_t nb[kMaxLevels];
    int nlevels;
    int k;
};

__host__ __device__ __forceinline__ uint64_t mix45(uint64_t x) {
    x ^= x >> 23;
    x = (x * 0x2545F4914F6CDD1Dull) & kKeyMask;
    x ^= x >> 21;
    x = (x * 0x1B03738712FAD5C9ull) & kKeyMask;
    x ^= x >> 24;
    return x;
}

__device__ __forceinline__ uint64_t home_bucket(uint64_t h, uint64_t nb) {
    return __umul64hi(h << (64 - kKeyBits), nb);
}

// 256-bit read-only load of one bucket: a single LDG.E.256 that touches exactly one sector.
__device__ __forceinline__ ulonglong4 load_bucket(const ulonglong4* p) {
    ulonglong4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(r.x), "=l"(r.y), "=l"(r.z), "=l"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ bool slot_matches(uint64_t slot, uint32_t want) {
    // `want` = disp<<28 | tag; the flag bit is ignored; EMPTY never matches because its value
    // field is kNoValue, which no resident slot may hold.
    return (((uint32_t)slot & ~kFlagBit) == want) & ((uint32_t)(slot >> 32) != kNoValue);
}

// Examines one loaded bucket.  Returns true when the probe is finished (hit or certain miss),
// false when the bucket is flagged and the probe must continue at the next bucket.
__device__ __forceinline__ bool probe_bucket(const ulonglong4& b, uint32_t want, uint32_t& value) {
    if (slot_matches(b.x, want)) { value = (uint32_t)(b.x >> 32); return true; }
    if (slot_matches(b.y, want)) { value = (uint32_t)(b.y >> 32); return true; }
    if (slot_matches(b.z, want)) { value = (uint32_t)(b.z >> 32); return true; }
    if (slot_matches(b.w, want)) { value = (uint32_t)(b.w >> 32); return true; }
    value = kNoValue;
    const bool flagged = ((uint32_t)b.x & kFlagBit) && (b.x != kEmptySlot);
    return !flagged;
}

// Continuation of a probe whose home bucket was flagged (rare path).
static __device__ __noinline__ uint32_t probe_slow(const TableView& t, uint64_t h) {
    uint32_t value = kNoValue;
    const uint32_t tag = (uint32_t)h & kTagMask;
    for (int lv = 0; lv < t.nlevels; ++lv) {
        const uint64_t nb = t.nb[lv];
        uint64_t b = home_bucket(h, nb);
        int d = 0;
        for (; d < kMaxDisp; ++d) {
            const ulonglong4 bk = load_bucket(t.level[lv] + b);
            if (probe_bucket(bk, ((uint32_t)d << 28) | tag, value)) return value;
            b = (b + 1 == nb) ? 0 : b + 1;
        }
        // kMaxDisp flagged buckets in a row: the key, if present, lives in the next level
    }
    return kNoValue;
}

// Full lookup of one packed key (used where lookups are not software-pipelined).
__device__ __forceinline__ uint32_t table_lookup(const TableView& t, uint64_t key) {
    if (key == kInvalidKey) return kNoValue;
    const uint64_t h = mix45(key);
    const ulonglong4 bk = load_bucket(t.level[0] + home_bucket(h, t.nb[0]));
    uint32_t value;
    if (probe_bucket(bk, (uint32_t)h & kTagMask, value)) return value;
    return probe_slow(t, h);
}

}  // namespace umgap
