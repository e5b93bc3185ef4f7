// GPU-resident k-mer -> taxon table (replaces fst::Map::get on the hot path; call sites
// prot2kmer2lca.rs:176, pept2lca.rs:93).
//
// Measured on B200 (bench/randsector.cu, bench/tma_gather.cu; profiles/): an L2 miss always fills
// a whole 128-byte line from HBM -- LDG, cp.async.bulk and TMA with L2 promotion NONE all read
// ~125 B of DRAM per 32-byte gather -- and random line fills top out at ~45 G lines/s.  The
// other three sectors of a fetched line are therefore free L2 hits.  The layout follows that:
//
//   level  = array of 128-byte LINES, each four 32-byte SECTORS (buckets) of four slots
//   sector = meta[4] (u32) | value[4] (u32)                    (structure of arrays, 32 B)
//   meta   = flag:1 | disp:3 | tag:28        empty slot: disp = 7, value = 0xFFFFFFFF
//
// A key is 5 bits per residue (per-index alphabet of <= 32 byte values), first residue most
// significant, at most 9 residues = 45 bits.  h = mix45(key) is a bijection on 45 bits.  The
// home line is floor((h >> 13) * nlines / 2^32), the home sector h & 3, the tag h's low 28 bits:
// every h that shares a home line lies in an interval narrower than 2^28 (nlines >= 2^17), so
// (line, tag) identifies the key exactly -- no false positives and no key bytes stored.
// Probe d (0..6) looks at sector (home + d) & 3 of line home_line + (d >> 2): the first four
// probes stay inside the home line (one DRAM fill, then L2 hits), three more use the next line.
// A key stored at distance d records d in `disp` and sets the overflow `flag` (top bit of
// meta[0]) of every full sector it passed; a key that finds all seven sectors full goes to the
// next, smaller level.  A probe reads ONE sector when the home sector holds the key or is not
// flagged -- a miss costs the same as a hit -- and continues only through flagged sectors.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace umgap {

constexpr int kMaxLevels = 4;
constexpr int kMaxDisp = 7;              // probe distances 0..6; disp == 7 marks an empty slot
constexpr int kKeyBits = 45;
constexpr uint64_t kKeyMask = (1ull << kKeyBits) - 1;
constexpr uint32_t kMinLines = 1u << 17;
constexpr uint32_t kTagMask = (1u << 28) - 1;
constexpr uint32_t kFlagBit = 1u << 31;
constexpr uint32_t kEmptyMeta = 0x7FFFFFFFu;
constexpr uint32_t kNoValue = 0xFFFFFFFFu;  // also UMGAP_MISS
constexpr uint64_t kInvalidKey = ~0ull;     // k-mer holding a byte outside the index alphabet

struct TableView {
    const ulonglong4* level[kMaxLevels];  // sector array, 4 sectors per line
    uint32_t nlines[kMaxLevels];
    int nlevels;
    int k;
    uint32_t nshards;  // > 1: this is one shard; its lines are addressed by the shard-local 32 bits
};

// Key-range-sharded table: shard o = floor((h >> 13) * nshards / 2^32) owns the key; the low 32
// bits of that product spread the shard's keys over its own lines.  level[o] of a remote shard is
// a peer mapping of that GPU's HBM (CUDA IPC over NVLink): the lookup kernel loads remote sectors
// directly, no routing kernels and no collective on the data path.
constexpr int kMaxShards = 8;
struct ShardedView {
    const ulonglong4* level[kMaxShards][kMaxLevels];
    uint32_t nlines[kMaxShards][kMaxLevels];
    int nlevels[kMaxShards];
    int nshards;
    int k;
};

__host__ __device__ __forceinline__ uint64_t mix45(uint64_t x) {
    x ^= x >> 22;
    x = (x * 0x2545F4914F6CDD1Dull) & kKeyMask;
    x ^= x >> 23;
    return x;
}

// Owner shard of hash h and the 32-bit value that places the key inside the shard (with one shard:
// owner 0 and the top 32 bits of h).
__host__ __device__ __forceinline__ uint32_t shard_split(uint64_t h, uint32_t nshards, uint32_t& local32) {
    const uint64_t prod = (uint64_t)(uint32_t)(h >> 13) * nshards;
    local32 = (uint32_t)prod;
    return (uint32_t)(prod >> 32);
}

// Sector index (line * 4 + sector) of probe d for a key placed by local32 in a level of nlines lines.
__device__ __forceinline__ uint64_t probe_sector_local(uint32_t local32, uint64_t h, uint32_t nlines, uint32_t d) {
    uint32_t line = __umulhi(local32, nlines) + (d >> 2);
    if (line >= nlines) line -= nlines;
    return (uint64_t)line * 4 + (((uint32_t)h + d) & 3u);
}
__device__ __forceinline__ uint64_t probe_sector(uint64_t h, uint32_t nlines, uint32_t d) {
    return probe_sector_local((uint32_t)(h >> 13), h, nlines, d);
}

// Address of the sector probe d of hash h reads in level lv, and the number of levels to try.
__device__ __forceinline__ const ulonglong4* sector_addr(const TableView& t, uint64_t h, uint32_t lv, uint32_t d) {
    uint32_t local32;
    shard_split(h, t.nshards, local32);  // nshards == 1: the top 32 bits of h
    return t.level[lv] + probe_sector_local(local32, h, t.nlines[lv], d);
}
__device__ __forceinline__ uint32_t num_levels(const TableView& t, uint64_t) { return (uint32_t)t.nlevels; }
__device__ __forceinline__ const ulonglong4* sector_addr(const ShardedView& t, uint64_t h, uint32_t lv, uint32_t d) {
    uint32_t local32;
    const uint32_t o = shard_split(h, (uint32_t)t.nshards, local32);
    return t.level[o][lv] + probe_sector_local(local32, h, t.nlines[o][lv], d);
}
__device__ __forceinline__ uint32_t num_levels(const ShardedView& t, uint64_t h) {
    uint32_t local32;
    return (uint32_t)t.nlevels[shard_split(h, (uint32_t)t.nshards, local32)];
}

// 256-bit read-only load of one sector: a single LDG.E.256.  UMGAP_PROBE_FILL (".L2::64B" / ".L2::128B", SASS LTC64B /
// LTC128B) is a measurement knob: an L2 miss of the plain form fills the whole 128-byte line, `.L2::64B` fills 64 bytes
// (bench/randsector variants 9-11: half the DRAM bytes at the same request rate).
#ifndef UMGAP_PROBE_FILL
#define UMGAP_PROBE_FILL ""
#endif
__device__ __forceinline__ ulonglong4 load_sector(const ulonglong4* p) {
    ulonglong4 r;
    asm volatile("ld.global.nc.L1::no_allocate" UMGAP_PROBE_FILL ".v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(r.x), "=l"(r.y), "=l"(r.z), "=l"(r.w)
                 : "l"(p));
    return r;
}

// Branch-free examination of one sector.  `want` = disp << 28 | tag.  Returns the value or
// kNoValue; `more` is set when the key was not found and the sector's overflow flag is up.
__device__ __forceinline__ uint32_t probe_sector_data(const ulonglong4& s, uint32_t want, bool& more) {
    const uint32_t m0 = (uint32_t)s.x, m1 = (uint32_t)(s.x >> 32), m2 = (uint32_t)s.y, m3 = (uint32_t)(s.y >> 32);
    uint32_t v = kNoValue;
    v = (m3 == want) ? (uint32_t)(s.w >> 32) : v;
    v = (m2 == want) ? (uint32_t)s.w : v;
    v = (m1 == want) ? (uint32_t)(s.z >> 32) : v;
    v = ((m0 & ~kFlagBit) == want) ? (uint32_t)s.z : v;
    more = (v == kNoValue) & ((m0 >> 31) != 0);
    return v;
}

// Remaining probes of a lookup whose probe d-1 ended on a flagged sector: continues at distance
// d of level lv, then through the overflow levels.
template <class TV>
__device__ __forceinline__ uint32_t probe_continue(const TV& t, uint64_t h, uint32_t lv, uint32_t d) {
    const uint32_t tag = (uint32_t)h & kTagMask;
    const uint32_t nlv = num_levels(t, h);
    for (; lv < nlv; ++lv, d = 0) {
        for (; d < (uint32_t)kMaxDisp; ++d) {
            const ulonglong4 s = load_sector(sector_addr(t, h, lv, d));
            bool more;
            const uint32_t v = probe_sector_data(s, (d << 28) | tag, more);
            if (!more) return v;
        }
        // seven flagged sectors in a row: the key, if present, lives in the next level
    }
    return kNoValue;
}

// Full lookup of one packed key (used where lookups are not software-pipelined).
template <class TV>
__device__ __forceinline__ uint32_t table_lookup(const TV& t, uint64_t key) {
    if (key == kInvalidKey) return kNoValue;
    return probe_continue(t, mix45(key), 0, 0);
}

// Layouts of a read's ids (2n words; strand s in words [s * n, s * n + npos)): position-major, or frame-major = the
// three frame records of a strand one after the other, each contiguous.  Index of position j of frame f of a strand
// in the frame-major layout, relative to the strand's first word (n nucleotides, k residues per key, n >= 3k).
__device__ __forceinline__ uint32_t frame_major_index(uint32_t n, uint32_t k, uint32_t f, uint32_t j) {
    const uint32_t c0 = n / 3 - k + 1, c1 = (n - 1) / 3 - k + 1;  // positions of frames 0 and 1
    return (f > 0 ? c0 : 0u) + (f > 1 ? c1 : 0u) + j;
}

}  // namespace umgap
