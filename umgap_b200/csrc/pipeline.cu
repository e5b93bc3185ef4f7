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
#include <atomic>
#include <map>
#include <mutex>
#include <thread>

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

void make_code_lut_public(const umgap_index* idx, int table, int methionine, uint8_t* out65);

static void make_code_lut(const umgap_index* idx, int table, int methionine, CodonLut& lut) {
    CodonLut a;
    make_ascii_lut(table, methionine, a);
    for (int i = 0; i < 65; ++i) lut.v[i] = idx->code_of_byte[a.v[i]];
}

void make_code_lut_public(const umgap_index* idx, int table, int methionine, uint8_t* out65) {
    CodonLut l{};
    make_code_lut(idx, table, methionine, l);
    memcpy(out65, l.v, 65);
}

// nucleotide byte -> 0..3 in codon order T,C,A,G (translation.rs:20); anything else, lowercase
// included, is N = 4 (dna/mod.rs:34-44).  Complement is code ^ 2 (dna/mod.rs:23-31).
__device__ __forceinline__ uint32_t nt_code(uint8_t c) {
    return c == 'T' ? 0u : c == 'C' ? 1u : c == 'A' ? 2u : c == 'G' ? 3u : 4u;
}

// Where a batch's nucleotides come from: the bytes as received, or the packed form of
// umgap_classify_reads_packed -- codes[x / 16] holds nucleotide x at bits 2 (x % 16) (A, C, G, T = 0, 1, 2, 3),
// nmask[x / 16] bit x % 16 is up when the byte was none of those letters (N, dna/mod.rs:34-44).
struct NtSrc {
    const uint8_t* bytes = nullptr;
    const uint32_t* codes = nullptr;
    const uint16_t* nmask = nullptr;
};

#ifndef UMGAP_K1_BLOCKS
#define UMGAP_K1_BLOCKS 3
#endif
constexpr int kTile = 128;          // k-mer start positions per warp pass (4 per lane)
constexpr int kLookupWarps = 8;     // warps per CTA
constexpr int kQueue = 2 * kTile;   // pending re-probes of one tile (both strands)
constexpr int kQueueWide = 2;       // queue entries re-probed per lane and pass

// Per-warp shared-memory scratch of the lookup: nucleotide codes, the two codon-start residue
// arrays (F = forward codon at x, R = codon of the reverse strand whose lowest forward
// coordinate is x) and the queue of lookups that must probe another sector.
template <int K>
struct LookupSmem {
    static constexpr int W = kTile + 3 * (K - 1);  // codon starts needed per tile
    uint8_t nt[W + 4];
    uint8_t f[W + 4];
    uint8_t r[W + 4];
    uint64_t q[kQueue];  // hash | next distance << 45 | level << 48 | strand << 50 | tile position << 51
};

// All k-mer lookups of one read by one warp.  Per tile of 128 start positions: nucleotide codes
// to shared memory, every codon start translated once per strand through the 65-entry LUT
// (codon -> 5-bit index-alphabet code), then per strand each lane packs 4 keys, issues their 4
// sector loads back to back and resolves them branch-free; lookups that ended on a flagged
// sector are compacted into the queue and re-probed 64 at a time (2 loads in flight per
// lane), so the warp stays converged.  out: forward k-mer starting at p -> out[p]; reverse-strand
// k-mer starting at reverse coordinate q -> out[n + q]  (frame f record = entries f-1, f+2, ...).
// Returns (warp-uniform) the read's frame mask: bit strand*3 + frame is set when that frame record
// holds at least one non-zero taxon id -- the classify kernel runs seedextend on those only.
// region_lo / region_hi: only k-mers whose hash prefix (h >> 13) lies in [region_lo, region_hi) are
// looked up and written in this pass (tables beyond the GPU's address-translation reach are probed
// one region at a time, see launch_translate_lookup); 0 / 2^32 selects everything.
// Two layouts of a read's ids (2n words; a strand's k-mers in words [strand * n, strand * n + npos)):
//   position-major  k-mer at coordinate y -> strand * n + y; frame f's record is the stride-3 subsequence from f;
//   frame-major     the three frame records of a strand one after the other, each contiguous (behind the sampled
//                   lookup kernel, which writes whole frame records: fewer partially written sectors).
// Index of position j of frame f of a strand in the frame-major layout, relative to the strand's first word.
// (frame_major_index: table.cuh)

template <int K, class TV, bool REGION, bool PACKED>
__device__ __forceinline__ uint32_t lookup_read(const TV& t, const uint8_t* s_lut, LookupSmem<K>& sm,
                                                const NtSrc& nv, uint64_t off, uint32_t n, uint32_t* out, int lane,
                                                uint64_t region_lo, uint64_t region_hi, bool frame_major = false) {
    constexpr int W = LookupSmem<K>::W;
    const unsigned lt_mask = (1u << lane) - 1;
    const uint32_t npos = n - 3u * K + 1;
    uint32_t hitbits = 0;
    for (uint32_t w0 = 0; w0 < npos; w0 += kTile) {
        for (int i = lane; i < W + 2; i += 32) {
            const uint32_t x = w0 + i;
            if (PACKED) {  // A,C,G,T = 0..3 -> codon order T,C,A,G (nibble k of 0x0312)
                const uint64_t g = off + x;
                const uint32_t c = x < n ? (__ldg(nv.codes + (g >> 4)) >> (2 * (g & 15))) & 3u : 0u;
                const bool isn = x >= n || ((__ldg(nv.nmask + (g >> 4)) >> (g & 15)) & 1u);
                sm.nt[i] = isn ? (uint8_t)4 : (uint8_t)((0x0312u >> (4 * c)) & 3u);
            } else {
                sm.nt[i] = x < n ? (uint8_t)nt_code(nv.bytes[off + x]) : (uint8_t)4;
            }
        }
        __syncwarp();
        for (int i = lane; i < W; i += 32) {
            const uint32_t a = sm.nt[i], b = sm.nt[i + 1], c = sm.nt[i + 2];
            const bool has_n = ((a | b | c) & 4u) != 0;
            sm.f[i] = s_lut[has_n ? 64 : 16 * a + 4 * b + c];
            sm.r[i] = s_lut[has_n ? 64 : 16 * (c ^ 2) + 4 * (b ^ 2) + (a ^ 2)];
        }
        __syncwarp();
        uint32_t qn = 0;
#pragma unroll 1
        for (int strand = 0; strand < 2; ++strand) {
            const uint8_t* codes = strand ? sm.r : sm.f;
            uint64_t h[4];
            ulonglong4 sec[4];
            bool valid[4];
            bool elsewhere[REGION ? 4 : 1];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int pl = lane + 32 * u;
                uint64_t key = 0;
                uint32_t bad = 0;
#pragma unroll
                for (int i = 0; i < K; ++i) {
                    // forward: residues at pl, pl+3, ...; reverse: the same codon starts read downwards
                    const uint32_t c = codes[pl + 3 * (strand ? K - 1 - i : i)];
                    bad |= c;
                    key = (key << 5) | (c & 31u);
                }
                h[u] = mix45(key);
                valid[u] = (w0 + pl < npos) && !(bad & 0x80u);
                if (REGION) {
                    // another pass answers this position: valid k-mers of other regions, and -- except in
                    // the first pass -- the k-mers that cannot be keys (a miss in every pass)
                    const uint64_t prefix = h[u] >> 13;
                    const bool mine = valid[u] ? (prefix >= region_lo && prefix < region_hi) : region_lo == 0;
                    elsewhere[u] = !mine;
                    valid[u] = valid[u] && mine;
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (valid[u]) sec[u] = load_sector(sector_addr(t, h[u], 0, 0));
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t p = w0 + lane + 32 * u;
                const uint32_t y = strand ? npos - 1 - p : p;
                const uint32_t pos = (strand ? n : 0u) + (frame_major ? frame_major_index(n, K, y % 3, y / 3) : y);
                bool more = false;
                uint32_t v = kNoValue;
                if (valid[u]) v = probe_sector_data(sec[u], (uint32_t)h[u] & kTagMask, more);
                if (p < npos && !more && !(REGION && elsewhere[REGION ? u : 0])) out[pos] = v;
                // frame of a k-mer: forward start p -> p % 3, reverse coordinate q = npos-1-p -> q % 3
                if (!more && v != kNoValue && v != 0) hitbits |= 1u << (strand * 3 + (strand ? npos - 1 - p : p) % 3);
                const unsigned m = __ballot_sync(0xffffffffu, more);
                if (more)
                    sm.q[qn + __popc(m & lt_mask)] = h[u] | (1ull << 45) | ((uint64_t)strand << 50) | ((uint64_t)(lane + 32 * u) << 51);
                qn += __popc(m);
            }
        }
        __syncwarp();
        // re-probe the flagged ones, densely packed: distance d, then d+1, ... then the next level
        while (qn) {
            uint32_t qnext = 0;
            for (uint32_t c = 0; c < qn; c += 32 * kQueueWide) {
                uint64_t hq[kQueueWide];
                ulonglong4 s2[kQueueWide];
#pragma unroll
                for (int j = 0; j < kQueueWide; ++j) {
                    const uint32_t i = c + 32 * j + lane;
                    hq[j] = i < qn ? sm.q[i] : ~0ull;
                }
#pragma unroll
                for (int j = 0; j < kQueueWide; ++j)
                    if (hq[j] != ~0ull)
                        s2[j] = load_sector(sector_addr(t, hq[j] & kKeyMask, (uint32_t)(hq[j] >> 48) & 3u, (uint32_t)(hq[j] >> 45) & 7u));
                __syncwarp();
#pragma unroll
                for (int j = 0; j < kQueueWide; ++j) {
                    bool more = false;
                    uint64_t next = 0;
                    if (hq[j] != ~0ull) {
                        uint32_t d = (uint32_t)(hq[j] >> 45) & 7u;
                        uint32_t lv = (uint32_t)(hq[j] >> 48) & 3u;
                        const uint64_t hh = hq[j] & kKeyMask;
                        const uint32_t v = probe_sector_data(s2[j], (d << 28) | ((uint32_t)hh & kTagMask), more);
                        if (more && ++d == (uint32_t)kMaxDisp) {
                            d = 0;
                            if (++lv == num_levels(t, hh)) more = false;  // v is kNoValue here
                        }
                        if (!more) {
                            const uint32_t p = w0 + ((uint32_t)(hq[j] >> 51) & 127u);
                            const uint32_t rev = (uint32_t)(hq[j] >> 50) & 1u;
                            const uint32_t y = rev ? npos - 1 - p : p;
                            out[(rev ? n : 0u) + (frame_major ? frame_major_index(n, K, y % 3, y / 3) : y)] = v;
                            if (v != kNoValue && v != 0) hitbits |= 1u << (rev * 3 + (rev ? npos - 1 - p : p) % 3);
                        }
                        next = (hq[j] & ~(0x1Full << 45)) | ((uint64_t)d << 45) | ((uint64_t)lv << 48);
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, more);
                    if (more) sm.q[qnext + __popc(m & lt_mask)] = next;
                    qnext += __popc(m);
                }
                __syncwarp();
            }
            qn = qnext;
        }
    }
    return __reduce_or_sync(0xffffffffu, hitbits);
}

// Work list of the plain kernel behind the sampled one: the reads (absolute indices) at list[base .. base + *count),
// base = first read of group g_lo when a group table is given, else 0.
struct ReadList {
    const uint32_t* list = nullptr;
    const uint32_t* count = nullptr;
    const uint64_t* group_off = nullptr;
    uint64_t g_lo = 0;
};

// Lookup kernel: one warp per read, ids to global memory, frame hit masks to frame_hits.
template <int K, class TV, bool REGION, bool PACKED>
__global__ void __launch_bounds__(kLookupWarps * 32, UMGAP_K1_BLOCKS)
translate_lookup_kernel(const __grid_constant__ TV t, const __grid_constant__ CodonLut lut, const NtSrc nv,
                        const uint64_t* __restrict__ read_off, uint64_t r_begin, uint64_t r_end,
                        uint32_t* __restrict__ ids, uint8_t* __restrict__ frame_hits, uint64_t region_lo,
                        uint64_t region_hi, ReadList rl /* rl.list set: only the reads of that list */) {
    __shared__ uint8_t s_lut[72];
    __shared__ LookupSmem<K> s_sm[kLookupWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < 72) s_lut[threadIdx.x] = lut.v[threadIdx.x];
    __syncthreads();
    const uint64_t nwarps = (uint64_t)gridDim.x * kLookupWarps;
    const uint32_t* list = nullptr;
    if (rl.list) {  // the reads the sampled kernel left over (longer than one of its batches)
        list = rl.list + (rl.group_off ? rl.group_off[rl.g_lo] : 0);
        r_begin = 0;
        r_end = *rl.count;
    }
    for (uint64_t i = r_begin + (uint64_t)blockIdx.x * kLookupWarps + warp; i < r_end; i += nwarps) {
        const uint64_t r = list ? list[i] : i;
        const uint64_t off = read_off[r];
        const uint32_t n = (uint32_t)(read_off[r + 1] - off);
        uint32_t mask = 0;  // a read none of whose frames reaches K residues has no records at all
        if (n >= 3u * K)
            mask = lookup_read<K, TV, REGION, PACKED>(t, s_lut, s_sm[warp], nv, off, n, ids + 2 * off, lane, region_lo, region_hi, list != nullptr);
        if (frame_hits && lane == 0) frame_hits[r] = (uint8_t)(!REGION || region_lo == 0 ? mask : (mask | frame_hits[r]));
    }
}

// ---- sampled lookups: seedextend makes most lookups provably irrelevant ----------------------------
// With `prot2kmer2lca -o | seedextend -s S` (S >= 2) a frame record contributes ids only if it holds
// a run of S equal non-zero ids at S CONSECUTIVE k-mer positions (seedextend.rs:137-149: same_max
// counts consecutive list elements, and with -o the list has one element per position).  Any S
// consecutive positions contain one whose index in the frame is a multiple of s = min(S, 4).  So the
// kernel first probes only those positions (1/s of the lookups); a frame none of whose sampled k-mers
// returned a non-zero taxon cannot hold a seed, its record comes out empty whatever the other
// positions hold, and they are never looked up.  The frames with a sampled hit -- the one or two true
// reading frames of a read from the index, plus the odd stray hit -- get their remaining positions
// probed in a second phase of the same warp, and only these frames' ids are written.  Bit-identical
// output, about 0.4 of the HBM line fills.
//
// Work layout: one warp takes a batch of up to kSReads = 5 consecutive reads (<= kSSpan = 1280 nucleotides;
// 30 frame records).  It translates the batch itself, 16 nucleotides per lane and step, four bytes per
// instruction, into two residue-code arrays in shared memory (forward codon starting at x, reverse-strand
// codon whose lowest forward coordinate is x).  Then every LANE walks one frame record: the 9-residue key
// rolls from one position to the next (one shared-memory byte per residue), two lookups in flight per
// lane.  The answers of a record's first kSValRows sampled positions stay in shared memory for the second
// phase.  A read longer than kSSpan is queued for the plain kernel.  (A streaming pre-pass kernel that
// left the residue codes in HBM was measured: 5.86 vs 5.49 ms per step -- 1.4 GB of extra DRAM traffic
// and a launch per slice.)
#ifndef UMGAP_S_BLOCKS
#define UMGAP_S_BLOCKS 7
#endif
constexpr int kSReads = 5;
constexpr int kSSpan = 256 * kSReads;  // five reads of up to 256 nt fill 30 lanes; longer reads leave lanes idle
#ifndef UMGAP_S_UNROLL
#define UMGAP_S_UNROLL 2
#endif
constexpr int kSU = UMGAP_S_UNROLL;     // lookups in flight per lane
constexpr int kSQueue = 64 * kSU;
#ifndef UMGAP_S_WARPS
#define UMGAP_S_WARPS 4
#endif
constexpr int kSWarps = UMGAP_S_WARPS;
constexpr int kSBlocks = UMGAP_S_BLOCKS;  // CTAs per SM the launch bounds ask for
constexpr int kSItems = 32 + 6 * kSReads;
constexpr int kSValRows = 16;
static_assert(6 * kSReads <= 32, "one lane per frame record of a batch");

// Residue codes of the 16 codon starts of one chunk, both strands: w[0..4] = its 16 nucleotide bytes plus the 4 that
// follow (two are needed).  Four bytes at a time: code = ((x >> 1) ^ (x >> 2)) & 3 maps A,C,G,T to 0,1,2,3; a byte is
// one of those letters iff the letter of its code equals it (anything else, lowercase included, is N: dna/mod.rs:34-44),
// N sets bit 2 of the byte's code.  The codon index rolls from one start to the next; pair[] holds the forward and the
// reverse-strand residue of each codon (see fill_pair_lut).
__device__ __forceinline__ void translate_chunk16(uint32_t (&w)[5], const uint16_t* pair, uint32_t (&fo)[4], uint32_t (&ro)[4]) {
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const uint32_t x = w[i];
        const uint32_t t = ((x >> 1) ^ (x >> 2)) & 0x03030303u;
        const uint32_t sel = (t & 0xFu) | ((t >> 4) & 0xF0u) | ((t >> 8) & 0xF00u) | ((t >> 12) & 0xF000u);
        const uint32_t diff = __byte_perm(0x54474341u, 0u, sel) ^ x;  // "ACGT"[code] against the byte
        const uint32_t nz = (((diff & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | diff) & 0x80808080u;  // 0x80 in every differing byte
        w[i] = t | (nz >> 5);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) fo[i] = ro[i] = 0;
    uint32_t idx = ((w[0] & 3u) << 2) | ((w[0] >> 8) & 3u);            // codon index so far: codes 0 and 1
    uint32_t nm = ((w[0] >> 2) & 1u) << 1 | ((w[0] >> 10) & 1u);        // their N flags
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const uint32_t c = (w[(i + 2) >> 2] >> (8 * ((i + 2) & 3))) & 7u;
        idx = ((idx << 2) | (c & 3u)) & 63u;
        nm = ((nm << 1) | (c >> 2)) & 7u;
        const uint32_t pr = pair[nm ? 64u : idx];
        fo[i >> 2] |= (pr & 0xFFu) << (8 * (i & 3));
        ro[i >> 2] |= (pr >> 8) << (8 * (i & 3));
    }
}

// The same from the packed form: codes = 20 nucleotides at 2 bits each (the chunk's 16 and the 4 that follow),
// nbits = their N flags.
__device__ __forceinline__ void translate_chunk16_packed(uint64_t codes, uint32_t nbits, const uint16_t* pair, uint32_t (&fo)[4],
                                                         uint32_t (&ro)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) fo[i] = ro[i] = 0;
    uint32_t idx = (((uint32_t)codes & 3u) << 2) | (((uint32_t)codes >> 2) & 3u);
    uint32_t nm = ((nbits & 1u) << 1) | ((nbits >> 1) & 1u);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const uint32_t c = (uint32_t)(codes >> (2 * (i + 2))) & 3u;
        idx = ((idx << 2) | c) & 63u;
        nm = ((nm << 1) | ((nbits >> (i + 2)) & 1u)) & 7u;
        const uint32_t pr = pair[nm ? 64u : idx];
        fo[i >> 2] |= (pr & 0xFFu) << (8 * (i & 3));
        ro[i >> 2] |= (pr >> 8) << (8 * (i & 3));
    }
}

// pair[i] = forward residue | reverse residue << 8 of the codon with index i = 16 a + 4 b + c in A,C,G,T order
// (complement = 3 - code); [64] = a codon holding an N.  lut.v is in T,C,A,G order (translation.rs:20).  Threads
// 0..64 of the CTA fill it; the caller synchronises.
__device__ __forceinline__ void fill_pair_lut(const CodonLut& lut, uint16_t* pair) {
    for (uint32_t i = threadIdx.x; i < 65; i += blockDim.x) {  // any block size
        if (i == 64) {
            pair[64] = (uint16_t)(lut.v[64] | (uint32_t)lut.v[64] << 8);
        } else {
            const uint32_t tcag = 0x0312u;  // nibble k = T,C,A,G index of A,C,G,T code k: A->2, C->1, G->3, T->0
            const uint32_t a = (tcag >> (4 * (i >> 4))) & 3u, bb = (tcag >> (4 * ((i >> 2) & 3u))) & 3u, c = (tcag >> (4 * (i & 3u))) & 3u;
            pair[i] = (uint16_t)(lut.v[16 * a + 4 * bb + c] | (uint32_t)lut.v[16 * (c ^ 2) + 4 * (bb ^ 2) + (a ^ 2)] << 8);
        }
    }
}

struct SampledSmem {
    uint4 f[(kSSpan + 32) / 16];  // code spans as copied in 16-byte chunks; the batch starts at byte `mis`
    uint4 r[(kSSpan + 32) / 16];
    uint64_t q[kSQueue];          // hash | distance << 45 | level << 48 | tag << 50
    uint32_t roff[kSReads + 1];   // read starts relative to the batch
    uint32_t mask[kSReads];       // frame hit masks
    // answer of sampled position s < kSValRows of the record walked by lane l: val[s][l]; the routed modes, which keep no
    // answers here, order the queue by owner in this space (route_flush)
    alignas(8) uint32_t val[kSValRows][32];
    uint16_t item[kSItems];       // second phase: frame record | segment << 5
};
constexpr uint32_t kSRevOff = sizeof(uint4) * ((kSSpan + 32) / 16);  // byte distance from f[] to r[]
static_assert(kSQueue * 12 <= kSValRows * 32 * 4, "route_flush orders a full queue (hash + position per entry) in val[]");

// Re-probes the queued lookups until all are answered; done(tag, value) receives each answer.
template <class TV, class Done>
__device__ __forceinline__ void drain_queue(const TV& t, uint64_t* q, uint32_t qn, int lane, Done done) {
    const unsigned lt_mask = (1u << lane) - 1;
    __syncwarp();
    while (qn) {
        uint32_t qnext = 0;
#pragma unroll 1
        for (uint32_t c = 0; c < qn; c += 32) {
            const uint32_t i = c + lane;
            const uint64_t hq = i < qn ? q[i] : ~0ull;
            bool more = false;
            uint64_t next = 0;
            if (hq != ~0ull) {
                uint32_t d = (uint32_t)(hq >> 45) & 7u, lv = (uint32_t)(hq >> 48) & 3u;
                const uint64_t hh = hq & kKeyMask;
                const ulonglong4 s2 = load_sector(sector_addr(t, hh, lv, d));
                const uint32_t v = probe_sector_data(s2, (d << 28) | ((uint32_t)hh & kTagMask), more);
                if (more && ++d == (uint32_t)kMaxDisp) {
                    d = 0;
                    if (++lv == num_levels(t, hh)) more = false;  // v is kNoValue here
                }
                if (!more) done((uint32_t)(hq >> 50), v);
                next = (hq & ~(0x1Full << 45)) | ((uint64_t)d << 45) | ((uint64_t)lv << 48);
            }
            __syncwarp();
            const unsigned m = __ballot_sync(0xffffffffu, more);
            if (more) q[qnext + __popc(m & lt_mask)] = next;
            qnext += __popc(m);
            __syncwarp();
        }
        qn = qnext;
    }
}

// One probe of every queued lookup, two sector loads in flight per lane; the lookups that need another probe are
// compacted to the front of the queue, their number is returned.  The region passes (MODE 1 / 2 of the sampled
// kernel) send ALL their lookups through the queue, first probes included (distance 0, level 0): a pass probes only
// 1/R of the positions a lane walks, and probing from the walk would spend a DRAM latency per step on a few lanes.
template <class TV, class Done>
__device__ __forceinline__ uint32_t drain_round(const TV& t, uint64_t* q, uint32_t qn, int lane, Done done) {
    const unsigned lt_mask = (1u << lane) - 1;
    uint32_t qnext = 0;
    __syncwarp();
#pragma unroll 1
    for (uint32_t c = 0; c < qn; c += 64) {
        uint64_t hq[2];
        ulonglong4 s2[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const uint32_t i = c + 32 * j + lane;
            hq[j] = i < qn ? q[i] : ~0ull;
        }
#pragma unroll
        for (int j = 0; j < 2; ++j)
            if (hq[j] != ~0ull)
                s2[j] = load_sector(sector_addr(t, hq[j] & kKeyMask, (uint32_t)(hq[j] >> 48) & 3u, (uint32_t)(hq[j] >> 45) & 7u));
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            bool more = false;
            uint64_t next = 0;
            if (hq[j] != ~0ull) {
                uint32_t d = (uint32_t)(hq[j] >> 45) & 7u, lv = (uint32_t)(hq[j] >> 48) & 3u;
                const uint64_t hh = hq[j] & kKeyMask;
                const uint32_t v = probe_sector_data(s2[j], (d << 28) | ((uint32_t)hh & kTagMask), more);
                if (more && ++d == (uint32_t)kMaxDisp) {
                    d = 0;
                    if (++lv == num_levels(t, hh)) more = false;  // v is kNoValue here
                }
                if (!more) done((uint32_t)(hq[j] >> 50), v);
                next = (hq[j] & ~(0x1Full << 45)) | ((uint64_t)d << 45) | ((uint64_t)lv << 48);
            }
            const unsigned m = __ballot_sync(0xffffffffu, more);
            if (more) q[qnext + __popc(m & lt_mask)] = next;
            qnext += __popc(m);
        }
        __syncwarp();
    }
    return qnext;
}

// Routed (key-range-sharded) mode: where the sampled kernel's lookups go instead of the local table -- the bucket of
// the shard that owns the hash (route.cu; the ranks swap the buckets by NCCL and answer them from their own shard).
struct RouteSink {
    BucketPtrs h;                           // per owner: its bucket of `cap` hashes (local memory, or the owner's inbox over NVLink)
    uint32_t* send_pos = nullptr;           // [nshards][cap] where the answer belongs (phase 1: read * 8 + frame, phase 2: ids index)
    unsigned long long* cursors = nullptr;  // [nshards] bucket fills, then [nshards] overflow flags
    uint64_t cap = 0;
    uint32_t nshards = 1;
};

// Appends the queued lookups (hash | tag << 50) to their owners' buckets.  One atomic per owner and flush: lane o
// counts the queue's entries bound for shard o and claims that many consecutive slots of bucket o (a claim per
// 32-entry chunk was measured: the two or eight cursors are same-address atomics, 10 M of them per 1 M pairs, and the
// pack kernels ran at a third of their speed).  The entries are first ordered by owner in shared memory (`scratch`:
// kSQueue hashes + kSQueue positions), so that the stores of a flush form one contiguous run per owner -- buckets that
// live in another GPU's memory (exchange.cu) then receive a few long NVLink writes per flush instead of a 32-byte
// piece per owner and chunk (eight owners: pack kernels 11 ms -> see profiles/README.md).  pos(tag) = the send_pos value.
template <class Pos>
__device__ __forceinline__ void route_flush(const RouteSink& rs, const uint64_t* q, uint32_t qn, int lane, Pos pos, uint32_t* scratch) {
    const unsigned lt_mask = (1u << lane) - 1;
    uint64_t* sh = reinterpret_cast<uint64_t*>(scratch);
    uint32_t* sp = scratch + 2 * kSQueue;
    __syncwarp();
    uint32_t mine = 0;  // lane o: entries bound for shard o
#pragma unroll 1
    for (uint32_t c = 0; c < qn; c += 32) {
        const bool active = c + lane < qn;
        uint32_t local32;
        const uint32_t owner = active ? shard_split(q[c + lane] & kKeyMask, rs.nshards, local32) : 0xFFFFFFFFu;
        for (uint32_t o = 0; o < rs.nshards; ++o) {
            const unsigned m = __ballot_sync(0xffffffffu, owner == o);
            if ((uint32_t)lane == o) mine += __popc(m);
        }
    }
    const uint32_t start = warp_incl_scan(mine, lane) - mine;  // lane o: where owner o's run begins in the ordered queue
    unsigned long long base = 0;  // lane o: first claimed slot of bucket o
    if ((uint32_t)lane < rs.nshards && mine) base = atomicAdd(&rs.cursors[lane], (unsigned long long)mine);
    uint32_t fill = start;
#pragma unroll 1
    for (uint32_t c = 0; c < qn; c += 32) {
        const bool active = c + lane < qn;
        const uint64_t e = active ? q[c + lane] : 0ull;
        const uint64_t h = e & kKeyMask;
        uint32_t local32;
        const uint32_t owner = active ? shard_split(h, rs.nshards, local32) : 0xFFFFFFFFu;
        const uint32_t at = __shfl_sync(0xffffffffu, fill, active ? (int)owner : 0);
        uint32_t before = 0;  // entries of this chunk bound for the same shard in lower lanes
        for (uint32_t o = 0; o < rs.nshards; ++o) {
            const unsigned m = __ballot_sync(0xffffffffu, owner == o);
            if (owner == o) before = __popc(m & lt_mask);
            if ((uint32_t)lane == o) fill += __popc(m);
        }
        if (active) {
            sh[at + before] = h;
            sp[at + before] = pos((uint32_t)(e >> 50));
        }
    }
    __syncwarp();
#pragma unroll 1
    for (uint32_t c = 0; c < qn; c += 32) {
        const uint32_t i = c + lane;
        const bool active = i < qn;
        const uint64_t h = active ? sh[i] : 0ull;
        uint32_t local32;
        const uint32_t owner = active ? shard_split(h, rs.nshards, local32) : 0u;
        const uint32_t st = __shfl_sync(0xffffffffu, start, (int)owner);
        const unsigned long long b0 = __shfl_sync(0xffffffffu, base, (int)owner);
        if (active) {
            const unsigned long long at = b0 + (i - st);
            if (at < rs.cap) {
                rs.h.p[owner][at] = h;
                rs.send_pos[(uint64_t)owner * rs.cap + at] = sp[i];
            } else {
                rs.cursors[rs.nshards + owner] = 1;  // bucket overflow: raised by umgap_exchange_status / RoutedClassifier.overflowed
            }
        }
    }
    __syncwarp();
}

// Makes room in (all = false) or empties (all = true) the per-warp queue of the sampled kernel, by its mode.
template <int MODE, class TV, class Done, class Pos>
__device__ __forceinline__ uint32_t service_queue(const TV& t, const RouteSink& rs, uint64_t* q, uint32_t qn, int lane, bool all,
                                                  uint32_t room, Done done, Pos pos, uint32_t* scratch) {
    if (MODE == 0) {
        drain_queue(t, q, qn, lane, done);
        return 0;
    }
    if (MODE <= 2) {
        if (all) {
            while (qn) qn = drain_round(t, q, qn, lane, done);
        } else {
            do qn = drain_round(t, q, qn, lane, done);
            while (qn + room > (uint32_t)kSQueue);
        }
        return qn;
    }
    route_flush(rs, q, qn, lane, pos, scratch);
    return 0;
}

// Geometry of frame record rec (= read-in-batch * 6 + frame) from the batch's read offsets: number of
// k-mer positions cntf, shared-memory byte address a0 of the first residue of position 0 (relative to
// the batch's first forward code), address step per position dir (+3 forward, -3 reverse; residue kk
// of position j is at a0 + dir * (j + kk)), and the index of position 0 in the batch's ids.
template <int K>
__device__ __forceinline__ uint32_t record_geometry(const SampledSmem& sm, uint32_t rec, uint32_t& a0, int& dir, uint32_t& o0) {
    const uint32_t i = rec / 6, fr = rec % 6, sd = fr >= 3 ? 1u : 0u, f = fr - 3 * sd;
    const uint32_t ro = sm.roff[i], n = sm.roff[i + 1] - ro;
    if (n < 3u * K) return 0;
    const uint32_t npos = n - 3u * K + 1;
    dir = sd ? -3 : 3;
    a0 = sd ? kSRevOff + ro + (npos - 1 - f) + 3u * (K - 1) : ro + f;
    o0 = 2 * ro + sd * n + frame_major_index(n, K, f, 0);
    return (npos - f + 2) / 3;  // positions y = f + 3j < npos
}

// MODE 0: both phases in one launch (tables within the address-translation reach).  Tables probed one hash-prefix
// region [region_lo, region_hi) per pass (launch_translate_lookup) run the phases as separate launches: MODE 1 =
// phase 1 of one region (only the frame masks leave the kernel, OR-ed over the passes), MODE 2 = phase 2 of one
// region (every position of the live frames whose hash prefix lies in the region, sampled ones included; batches
// without a live frame are not even translated).  MODE 3 / 4: the two phases of the routed mode -- the walk is the same,
// the lookups go to the owners' buckets (RouteSink) instead of the table; phase 1's answers come back as frame masks
// (route_scatter_hits_kernel), phase 2 sends every position of the live frames.
template <int K, class TV, int STRIDE, int MODE, bool PACKED>
__global__ void __launch_bounds__(kSWarps * 32, kSBlocks)
lookup_sampled_kernel(const __grid_constant__ TV t, const __grid_constant__ CodonLut lut, const NtSrc nv,
                      uint64_t total_nt, const uint64_t* __restrict__ read_off, uint32_t nreads, uint32_t* __restrict__ ids,
                      uint8_t* __restrict__ frame_hits, const uint64_t* __restrict__ group_off, uint64_t g_lo, uint64_t g_hi,
                      uint32_t* __restrict__ long_list, uint32_t* __restrict__ long_count, uint32_t* __restrict__ unit_count,
                      uint64_t region_lo, uint64_t region_hi, const __grid_constant__ RouteSink rs) {
    constexpr bool kPhase2Only = MODE == 2 || MODE == 4;
    __shared__ SampledSmem s_sm[kSWarps];
    __shared__ uint16_t s_pair[65];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1;
    fill_pair_lut(lut, s_pair);
    __syncthreads();
    SampledSmem& sm = s_sm[warp];
    // the reads of groups [g_lo, g_hi) when a group table is given (slices of the device path), else all reads
    const uint32_t r_begin = group_off ? (uint32_t)group_off[g_lo] : 0u;
    const uint32_t r_end = group_off ? (uint32_t)group_off[g_hi] : nreads;
    const uint32_t nunits = (r_end - r_begin + kSReads - 1) / kSReads;
    // units are handed out through a counter: their cost varies (reads with and without live frames), and a launch
    // holds only a few units per resident warp
#pragma unroll 1
    for (;;) {
        uint32_t unit = 0;
        if (lane == 0) unit = atomicAdd(unit_count, 1u);
        unit = __shfl_sync(0xffffffffu, unit, 0);
        if (unit >= nunits) break;
        uint32_t cur = r_begin + unit * kSReads;
        const uint32_t end = min(cur + (uint32_t)kSReads, r_end);
#pragma unroll 1
        while (cur < end) {
            // ---- batch geometry: as many of the unit's remaining reads as fit kSSpan nucleotides
            const uint32_t left = end - cur;
            const uint64_t my_off = read_off[cur + min((uint32_t)lane, left)];
            const uint64_t off0 = __shfl_sync(0xffffffffu, my_off, 0);
            const unsigned fits = __ballot_sync(0xffffffffu, lane >= 1 && (uint32_t)lane <= left && my_off - off0 <= (uint64_t)kSSpan);
            const uint32_t nb = (uint32_t)__popc(fits);
            if (nb == 0) {  // a single read longer than the batch span: queued for the plain kernel (launched next)
                if ((MODE == 0 || MODE == 3 || (MODE == 1 && region_lo == 0)) && lane == 0) long_list[r_begin + atomicAdd(long_count, 1u)] = cur;
                if (MODE == 3 && lane == 0)  // routed: the plain pack kernel sends every position of it (word-wise OR: another
                                             // stream may be OR-ing a neighbouring read's mask into the same word)
                    atomicOr(reinterpret_cast<uint32_t*>(frame_hits) + (cur >> 2), 0x3Fu << (8 * (cur & 3u)));
                cur += 1;
                continue;
            }
            const uint32_t rel = (uint32_t)(my_off - off0);
            const uint32_t span = __shfl_sync(0xffffffffu, rel, nb);
            if ((uint32_t)lane <= nb) sm.roff[lane] = rel;
            if ((uint32_t)lane < nb) sm.mask[lane] = kPhase2Only ? (uint32_t)frame_hits[cur + lane] : 0u;
            if (kPhase2Only) {  // nothing to do for a batch without a live frame
                __syncwarp();
                const uint32_t live = (uint32_t)lane < 6 * nb ? (sm.mask[lane / 6] >> (lane % 6)) & 1u : 0u;
                if (!__any_sync(0xffffffffu, live)) {
                    __syncwarp();
                    cur += nb;
                    continue;
                }
            }
            // ---- stage both code spans, 16 bytes per lane and load
            // ---- stage: 16 nucleotides per lane and step (plus the 4 bytes that follow them: two are needed), translated
            //      in registers (translate.rs:114-133, dna/mod.rs:23-103, dna/translation.rs:125-144), both residue-code
            //      chunks to shared memory; the batch starts at byte `mis` of the first chunk
            const uint32_t mis = PACKED ? (uint32_t)(off0 & 15u) : (uint32_t)((uintptr_t)(nv.bytes + off0) & 15u);
            {
                const uint64_t x_first = off0 - mis;  // the byte array is 16-byte aligned, so is this
                const uint32_t nch = (mis + span + 15) / 16;
                for (uint32_t c = lane; c < nch; c += 32) {
                    const uint64_t x0 = x_first + 16ull * c;
                    uint32_t fo[4], ro[4];
                    if (PACKED) {  // one word of codes and one of N flags per chunk, plus the low bits of the next
                        const uint64_t wi = x0 >> 4;
                        uint64_t codes = __ldg(nv.codes + wi);
                        uint32_t nbits = __ldg(nv.nmask + wi);
                        if (x0 + 16 < total_nt) {
                            codes |= (uint64_t)(__ldg(nv.codes + wi + 1) & 0xFFu) << 32;
                            nbits |= ((uint32_t)__ldg(nv.nmask + wi + 1) & 0xFu) << 16;
                        } else {
                            nbits |= 0xFu << 16;  // N past the end
                        }
                        translate_chunk16_packed(codes, nbits, s_pair, fo, ro);
                        sm.f[c] = make_uint4(fo[0], fo[1], fo[2], fo[3]);
                        sm.r[c] = make_uint4(ro[0], ro[1], ro[2], ro[3]);
                        continue;
                    }
                    const uint8_t* nt = nv.bytes;
                    uint32_t w[5];
                    const uint4 v = __ldg(reinterpret_cast<const uint4*>(nt + x0));  // the chunk holds a nucleotide of the batch
                    w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
                    w[4] = x0 + 20 <= total_nt ? __ldg(reinterpret_cast<const uint32_t*>(nt + x0 + 16)) : 0x4E4E4E4Eu;  // 'N's past the end
                    translate_chunk16(w, s_pair, fo, ro);
                    sm.f[c] = make_uint4(fo[0], fo[1], fo[2], fo[3]);
                    sm.r[c] = make_uint4(ro[0], ro[1], ro[2], ro[3]);
                }
            }
            __syncwarp();
            const uint8_t* cb = reinterpret_cast<const uint8_t*>(sm.f) + mis;
            uint32_t* const out0 = ids + 2 * off0;
            uint32_t qn = 0;
            // ---- phase 1: lane `rec` walks the sampled positions (j = 0, STRIDE, ...) of frame record `rec`
            const uint32_t rec = lane;
            auto done1 = [&](uint32_t tag, uint32_t v) {  // tag = record | sampled index << 5
                const uint32_t rc = tag & 31u, sx = tag >> 5;
                if (MODE == 0 && sx < (uint32_t)kSValRows) sm.val[sx][rc] = v;
                if (v != kNoValue && v != 0) atomicOr(&sm.mask[rc / 6], 1u << (rc % 6));
            };
            const uint32_t cur0 = cur;
            auto pos1 = [&](uint32_t tag) { return (cur0 + tag / 6u) * 8u + tag % 6u; };       // routed phase 1: read * 8 + frame
            auto pos2 = [&](uint32_t tag) { return (uint32_t)(2 * off0) + tag; };               // routed phase 2: index into ids
            if (!kPhase2Only) {
                uint32_t a = 0, o0 = 0, cs = 0;
                int dir = 3;
                if (rec < 6 * nb) cs = (record_geometry<K>(sm, rec, a, dir, o0) + STRIDE - 1) / STRIDE;
                uint64_t key = 0;
                uint32_t bad = 0;
                if (cs) {
#pragma unroll
                    for (int kk = 0; kk < K; ++kk) {
                        const uint32_t c = cb[a + dir * kk];
                        key = (key << 5) | (c & 31u);
                        bad = (bad << 1) | (c >> 7);
                    }
                }
                const uint32_t max_cs = __reduce_max_sync(0xffffffffu, cs);
#pragma unroll 1
                for (uint32_t s0 = 0; s0 < max_cs; s0 += kSU) {
                    if (qn + 32 * kSU > (uint32_t)kSQueue) qn = service_queue<MODE>(t, rs, sm.q, qn, lane, false, 32 * kSU, done1, pos1, &sm.val[0][0]);
                    uint64_t h[kSU];
                    ulonglong4 sec[kSU];
                    bool valid[kSU];
#pragma unroll
                    for (int u = 0; u < kSU; ++u) {
                        h[u] = mix45(key);
                        valid[u] = s0 + u < cs && bad == 0;
                        if (MODE == 1) valid[u] = valid[u] && (h[u] >> 13) >= region_lo && (h[u] >> 13) < region_hi;
                        if (MODE == 0 && s0 + u < cs && bad != 0 && s0 + u < (uint32_t)kSValRows) sm.val[s0 + u][lane] = kNoValue;
                        if (s0 + u + 1 < cs) {  // roll on to the next sampled position
#pragma unroll
                            for (int m = 0; m < STRIDE; ++m) {
                                const uint32_t c = cb[a + dir * (K + m)];
                                key = ((key << 5) | (c & 31u)) & kKeyMask;
                                bad = ((bad << 1) | (c >> 7)) & ((1u << K) - 1);
                            }
                            a += dir * STRIDE;
                        }
                    }
                    if (MODE == 1 || MODE == 3) {  // region pass: the in-region lookups are probed from the queue, densely; routed: all go to the buckets
#pragma unroll
                        for (int u = 0; u < kSU; ++u) {
                            const unsigned m = __ballot_sync(0xffffffffu, valid[u]);
                            if (valid[u]) sm.q[qn + __popc(m & lt_mask)] = h[u] | ((uint64_t)rec << 50);
                            qn += __popc(m);
                        }
                    } else {
#pragma unroll
                        for (int u = 0; u < kSU; ++u)
                            if (valid[u]) sec[u] = load_sector(sector_addr(t, h[u], 0, 0));
#pragma unroll
                        for (int u = 0; u < kSU; ++u) {
                            const uint32_t tag = rec | ((s0 + u) << 5);
                            bool more = false;
                            if (valid[u]) {
                                const uint32_t v = probe_sector_data(sec[u], (uint32_t)h[u] & kTagMask, more);
                                if (!more) done1(tag, v);
                            }
                            const unsigned m = __ballot_sync(0xffffffffu, more);
                            if (more) sm.q[qn + __popc(m & lt_mask)] = h[u] | (1ull << 45) | ((uint64_t)tag << 50);
                            qn += __popc(m);
                        }
                    }
                }
            }
            qn = service_queue<MODE>(t, rs, sm.q, qn, lane, true, 0, done1, pos1, &sm.val[0][0]);
            __syncwarp();
            if (MODE == 0 && (uint32_t)lane < nb) frame_hits[cur + lane] = (uint8_t)sm.mask[lane];
            if (MODE == 1) {  // the masks accumulate over the region passes; phase 2 is another launch
                if ((uint32_t)lane < nb) frame_hits[cur + lane] = (uint8_t)(region_lo == 0 ? sm.mask[lane] : (sm.mask[lane] | frame_hits[cur + lane]));
                __syncwarp();
                cur += nb;
                continue;
            }
            if (MODE == 3) {  // the masks come back from the owners of the hashes
                __syncwarp();
                cur += nb;
                continue;
            }
            // ---- phase 2: every position of the frames with a sampled hit, in segments of `seg` positions per lane;
            //      sampled positions whose answer is still in val[] are copied, the others probed
            uint32_t cntf = 0;
            {
                uint32_t a, o0;
                int dir;
                if (rec < 6 * nb && (sm.mask[rec / 6] >> (rec % 6) & 1u)) cntf = record_geometry<K>(sm, rec, a, dir, o0);
            }
            const uint32_t total = __reduce_add_sync(0xffffffffu, cntf);
            if (total) {
                const uint32_t seg = max(4u, (total + 27) / 28);
                const uint32_t nseg = (cntf + seg - 1) / seg;
                const uint32_t incl = warp_incl_scan(nseg, lane);
                for (uint32_t g = 0; g < nseg; ++g) sm.item[incl - nseg + g] = (uint16_t)(rec | (g << 5));
                const uint32_t nitems = __shfl_sync(0xffffffffu, incl, 31);
                __syncwarp();
                auto done2 = [&](uint32_t o, uint32_t v) { out0[o] = v; };
#pragma unroll 1
                for (uint32_t it0 = 0; it0 < nitems; it0 += 32) {
                    uint32_t a = 0, o = 0, cnt = 0, j = 0, rc = 0;
                    int dir = 3;
                    if (it0 + lane < nitems) {
                        const uint32_t item = sm.item[it0 + lane];
                        uint32_t a0, o0;
                        rc = item & 31u;
                        const uint32_t n_pos = record_geometry<K>(sm, rc, a0, dir, o0);
                        j = (item >> 5) * seg;
                        cnt = min(seg, n_pos - j);
                        a = a0 + dir * j;
                        o = o0 + j;
                    }
                    uint64_t key = 0;
                    uint32_t bad = 0;
                    if (cnt) {
#pragma unroll
                        for (int kk = 0; kk < K; ++kk) {
                            const uint32_t c = cb[a + dir * kk];
                            key = (key << 5) | (c & 31u);
                            bad = (bad << 1) | (c >> 7);
                        }
                    }
                    const uint32_t max_cnt = __reduce_max_sync(0xffffffffu, cnt);
#pragma unroll 1
                    for (uint32_t s0 = 0; s0 < max_cnt; s0 += kSU) {
                        if (qn + 32 * kSU > (uint32_t)kSQueue) qn = service_queue<MODE>(t, rs, sm.q, qn, lane, false, 32 * kSU, done2, pos2, &sm.val[0][0]);
                        uint64_t h[kSU];
                        ulonglong4 sec[kSU];
                        uint32_t v[kSU];
                        bool valid[kSU];
                        bool mine[kSU];  // MODE 2: this pass answers the position
#pragma unroll
                        for (int u = 0; u < kSU; ++u) {
                            const bool act = s0 + u < cnt;
                            const uint32_t jj = j + s0 + u, sx = jj / STRIDE;
                            const bool cached = MODE == 0 && act && jj % STRIDE == 0 && sx < (uint32_t)kSValRows;
                            v[u] = cached ? sm.val[sx][rc] : kNoValue;
                            h[u] = mix45(key);
                            valid[u] = act && !cached && bad == 0;
                            mine[u] = act;
                            if (MODE == 2) {  // a k-mer that cannot be a key is a miss in every pass: the first one writes it
                                mine[u] = act && (bad == 0 ? ((h[u] >> 13) >= region_lo && (h[u] >> 13) < region_hi) : region_lo == 0);
                                valid[u] = valid[u] && mine[u];
                            }
                            if (s0 + u + 1 < cnt) {
                                const uint32_t c = cb[a + dir * K];
                                key = ((key << 5) | (c & 31u)) & kKeyMask;
                                bad = ((bad << 1) | (c >> 7)) & ((1u << K) - 1);
                                a += dir;
                            }
                        }
                        if (kPhase2Only) {  // region pass: in-region lookups through the queue (routed: all, to the buckets); a k-mer that cannot be a key is a miss
#pragma unroll
                            for (int u = 0; u < kSU; ++u) {
                                const uint32_t oi = o + s0 + u;
                                if (mine[u] && !valid[u]) out0[oi] = kNoValue;
                                const unsigned m = __ballot_sync(0xffffffffu, valid[u]);
                                if (valid[u]) sm.q[qn + __popc(m & lt_mask)] = h[u] | ((uint64_t)oi << 50);
                                qn += __popc(m);
                            }
                        } else {
#pragma unroll
                            for (int u = 0; u < kSU; ++u)
                                if (valid[u]) sec[u] = load_sector(sector_addr(t, h[u], 0, 0));
#pragma unroll
                            for (int u = 0; u < kSU; ++u) {
                                bool more = false;
                                if (valid[u]) v[u] = probe_sector_data(sec[u], (uint32_t)h[u] & kTagMask, more);
                                const uint32_t oi = o + s0 + u;
                                if (mine[u] && !more) out0[oi] = v[u];
                                const unsigned m = __ballot_sync(0xffffffffu, more);
                                if (more) sm.q[qn + __popc(m & lt_mask)] = h[u] | (1ull << 45) | ((uint64_t)oi << 50);
                                qn += __popc(m);
                            }
                        }
                    }
                }
                qn = service_queue<MODE>(t, rs, sm.q, qn, lane, true, 0, done2, pos2, &sm.val[0][0]);
            }
            __syncwarp();
            cur += nb;
        }
    }
}

// ---- classify: seedextend + uniq join + aggregate, one warp per group ----------------------------
struct ClassifyParams {
    int frame_major;  // layout of the ids (see frame_major_index): 1 behind the sampled lookup kernel
    int k;
    int one_on_one;
    int seedextend;
    uint32_t min_seed, max_gap;
    AggParams agg;
};

constexpr int kAggWarps = 4;
constexpr int kAggSlots = 8;        // groups whose live frame records are pooled by one warp
constexpr uint32_t kAggTable = 64;  // slots of a group's (taxon -> occurrences) table in shared memory
constexpr uint32_t kAggFull = 48;   // distinct taxa beyond which a group is redone through global scratch

struct DevError {  // first error raised by a kernel
    unsigned int flag;
    unsigned int taxon;
};

// Where seedextend's kept ids go.  The aggregators only need the multiset of non-zero ids
// (taxa2agg.rs:169 drops zeros, agg/mod.rs:27-36 counts), so a group's kept ids are counted on the fly:
// TableSink -- open-addressing table of kAggTable (taxon, occurrences) slots in shared memory, filled with
//              atomics by the lanes that run the group's frame records; `distinct` counts the slots taken.
// ListSink  -- append-only (taxon, occurrences) list of unbounded size in global scratch (huge groups).
struct TableSink {
    uint32_t* keys;      // 0 = free
    uint32_t* cnts;
    uint32_t* distinct;
    __device__ __forceinline__ void add(uint32_t id, uint32_t len) const {
        uint32_t slot = (id * 0x9E3779B1u) >> 26;
        for (uint32_t probe = 0; probe < kAggTable; ++probe) {
            const uint32_t prev = atomicCAS(&keys[slot], 0u, id);
            if (prev == 0u) atomicAdd(distinct, 1u);
            if (prev == 0u || prev == id) {
                atomicAdd(&cnts[slot], len);
                return;
            }
            slot = (slot + 1) & (kAggTable - 1);
        }
        atomicAdd(distinct, kAggTable);  // table full: the group takes the list path
    }
};
struct ListSink {
    uint32_t* A;
    uint32_t* C;
    uint32_t* counter;
    __device__ __forceinline__ void add(uint32_t id, uint32_t len) const {
        const uint32_t at = atomicAdd(counter, 1u);
        A[at] = id;
        C[at] = len;
    }
};

// seedextend of one frame record (fr = 0,1,2 forward; 3,4,5 reverse) of a read with n nucleotides whose
// ids start at read_ids; the kept ids leave as run-length pairs through sink.add(id, occurrences).
// Returns false when the record does not exist (prot2kmer2lca.rs:172).  read_priv: the read's 4n words
// of private scratch.
template <class Sink>
__device__ __forceinline__ bool seedextend_frame(const ClassifyParams& cp, const uint32_t* read_ids, uint32_t n,
                                                 uint32_t fr, const Sink& sink, uint32_t* read_priv) {
    const uint32_t f = fr % 3;
    const uint32_t plen = n >= f ? (n - f) / 3 : 0;  // peptide length of the frame
    if (plen < (uint32_t)cp.k) return false;
    const uint32_t cnt = plen - cp.k + 1;
    const uint32_t st = cp.frame_major ? 1u : 3u;  // distance of consecutive positions of the record
    const uint32_t* base = read_ids + (fr >= 3 ? n : 0) + (cp.frame_major ? frame_major_index(n, (uint32_t)cp.k, f, 0) : f);
    // private slice of this record: pair i at words 6i, 6i+1 past 2*(strand*n + f), inside the read's 4n words
    uint32_t* priv = read_priv + 2 * ((fr >= 3 ? n : 0) + f);
    if (cp.seedextend && cp.one_on_one && cp.min_seed >= 2 && cnt <= 63) {
        // The seedextend machine (seedextend.rs:101-149) in closed form on bit masks, one record per lane, no
        // per-element state: a gap of more than max_gap zeros ends a range (:116-127), shorter gaps stay inside it,
        // so the ranges are the stretches between long gaps; a range is selected iff it holds min_seed equal
        // consecutive non-zero ids (:137-149; any zero breaks a run).  The one irregularity of the machine: a record
        // that BEGINS with 1..max_gap zeros loses its first non-zero id (:130-134 moves the range start past t[end]
        // without making it `last`), after which everything restarts behind that id.  (Checked against the
        // line-by-line restatement on random lists: tests/test_host_cpu.py::test_seedextend_closed_form.)
        // nz = id != 0, eq = id equals its predecessor; the kept ids leave as (id, run length).
        uint32_t nz_lo = 0, eq_lo = 0, nz_hi = 0, eq_hi = 0, prev = kNoValue;
        const uint32_t c_lo = cnt < 32 ? cnt : 32;
#pragma unroll 4
        for (uint32_t i = 0; i < c_lo; ++i) {
            uint32_t v = base[st * i];
            v = v == kNoValue ? 0u : v;
            nz_lo |= (uint32_t)(v != 0) << i;
            eq_lo |= (uint32_t)(v == prev) << i;
            prev = v;
        }
#pragma unroll 4
        for (uint32_t i = 32; i < cnt; ++i) {
            uint32_t v = base[st * i];
            v = v == kNoValue ? 0u : v;
            nz_hi |= (uint32_t)(v != 0) << (i - 32);
            eq_hi |= (uint32_t)(v == prev) << (i - 32);
            prev = v;
        }
        uint64_t nz = (uint64_t)nz_hi << 32 | nz_lo, eq = (uint64_t)eq_hi << 32 | eq_lo;
        if (!nz) return true;
        uint64_t inr = nz;  // positions inside a range
        if (cp.max_gap) {
            const uint32_t G = cp.max_gap < 63 ? cp.max_gap : 63;
            uint64_t valid = (1ull << cnt) - 1;
            const uint32_t z = (uint32_t)__ffsll((long long)nz) - 1;  // leading zeros
            if (z >= 1 && z <= G) {  // the record starts with a short gap: positions 0..z vanish, a fresh run starts at z + 1
                const uint64_t above = ~((2ull << z) - 1);
                nz &= above;
                valid &= above;
                eq &= above & ~(1ull << (z + 1));
            }
            const uint64_t zero = ~nz & valid;
            uint64_t lng = zero;  // first positions of G + 1 consecutive zeros ...
            for (uint32_t g = 1; g <= G && lng; ++g) lng &= zero >> g;
            uint64_t cover = lng;  // ... and every position of such a gap
            for (uint32_t g = 1; g <= G; ++g) cover |= lng << g;
            inr = valid & ~cover;
        }
        uint64_t seed = nz & eq;
        for (uint32_t k = 1; k + 1 < cp.min_seed && seed; ++k) seed &= eq << k;
        uint64_t kept = 0;
        while (seed) {  // the range around the lowest remaining seed
            const uint32_t sp = (uint32_t)__ffsll((long long)seed) - 1;
            const uint64_t up = ((inr + (1ull << sp)) ^ inr) & inr;          // from the seed to the range's last position
            const uint64_t out_below = ~inr & ((1ull << sp) - 1);
            const uint32_t first = out_below ? 64u - (uint32_t)__clzll((long long)out_below) : 0u;
            const uint64_t range = up | (((1ull << sp) - 1) & ~((1ull << first) - 1));
            kept |= range;
            seed &= ~range;
        }
        const uint64_t ends = ~(kept & nz & eq);  // where a run of equal kept ids stops (bit 63 included)
        uint64_t heads = kept & nz & ~eq;
        while (heads) {
            const uint32_t hp = (uint32_t)__ffsll((long long)heads) - 1;
            heads &= heads - 1;
            sink.add(base[st * hp], (uint32_t)__ffsll((long long)(ends >> (hp + 1))));
        }
        return true;
    }
    if (cp.seedextend && cp.one_on_one) {
        // Single pass of the seedextend machine (seedextend.rs:101-149) that builds the run-length
        // list of the CURRENT range tentatively in the record's private slice and commits it when the
        // range is selected / rolls it back when the range is abandoned -- no second walk over the
        // selected ranges.  Zeros never break a run (they are dropped before counting).
        uint32_t k = 0, committed = 0, rid = 0, rlen = 0;
        auto close_run = [&]() {
            if (rlen) {
                priv[6 * k] = rid;
                priv[6 * k + 1] = rlen;
                ++k;
            }
            rid = 0;
            rlen = 0;
        };
        auto feed = [&](uint32_t v) {
            if (v == 0) return;
            if (v == rid) {
                ++rlen;
            } else {
                close_run();
                rid = v;
                rlen = 1;
            }
        };
        auto at = [&](uint32_t i) -> uint32_t {
            const uint32_t v = base[st * i];
            return v == kNoValue ? 0u : v;
        };
        // The four cases of the reference loop body are evaluated as predicates so that the lanes of
        // a warp (one record each, different data) do not serialise on divergent branches; only the
        // rare stores of closed runs branch.  The next element is fetched one iteration ahead.
        uint32_t last = at(0), start = 0, same = 1, smax = 1;
        feed(last);
        uint32_t nxt = cnt > 1 ? at(1) : 0u;
        for (uint32_t end = 1; end <= cnt; ++end) {
            const uint32_t cur = nxt;                        // the sentinel 0 (:99) when end == cnt
            nxt = end + 1 < cnt ? at(end + 1) : 0u;
            const bool differs = cur != last;
            const bool gap_close = differs && last == 0 && same > cp.max_gap;                // :116-127
            const bool lead_gap = differs && !gap_close && last == 0 && end - start == same;  // :130-134
            const bool other = differs && !gap_close && !lead_gap;                            // :137-142
            if (gap_close) {  // the range ends before the gap: keep its runs if it holds a seed
                close_run();
                if (smax >= cp.min_seed) committed = k; else k = committed;
            }
            smax = gap_close ? 1u : (other && last != 0 && same > smax) ? same : smax;
            start = gap_close ? end : lead_gap ? end + 1 : start;
            same = !differs ? same + 1 : lead_gap ? same : 1u;   // a leading gap leaves last / same stale
            last = lead_gap ? last : cur;
            if (!lead_gap && end < cnt) feed(cur);               // a leading gap also skips t[end]
        }
        close_run();
        if (smax >= cp.min_seed) committed = k;              // :144-149 (trailing zeros carry no ids)
        for (uint32_t i = 0; i < committed; ++i) sink.add(priv[6 * i], priv[6 * i + 1]);
        return true;
    }
    uint32_t run_id = 0, run_len = 0;
    auto push = [&](uint32_t v) {
        if (v == 0) return;
        if (v == run_id) {
            ++run_len;
        } else {
            if (run_len) sink.add(run_id, run_len);
            run_id = v;
            run_len = 1;
        }
    };
    if (cp.seedextend) {
        seedextend_stream(base, st, cnt, cp.one_on_one != 0, cp.min_seed, cp.max_gap, push);
    } else {
        for (uint32_t i = 0; i < cnt; ++i) {
            const uint32_t v = base[st * i];
            if (v != kNoValue) push(v);
        }
    }
    if (run_len) sink.add(run_id, run_len);
    return true;
}

// The same for record rec (= read-in-group * 6 + frame) of the group starting at read r0, ids in
// global memory.
template <class Sink>
__device__ __forceinline__ bool seedextend_record(const ClassifyParams& cp, const uint32_t* __restrict__ ids,
                                                  const uint64_t* __restrict__ read_off, uint64_t r0, uint32_t rec,
                                                  const Sink& sink, uint32_t* scratch) {
    // scratch holds 12 words per nucleotide: a group's slice is [12*off(r0), ...): first 4 words per
    // nucleotide of private record slices, then the four overflow lists of 2 words per nucleotide
    const uint64_t r = r0 + rec / 6;
    const uint64_t off = read_off[r], off0 = read_off[r0];
    return seedextend_frame(cp, ids + 2 * off, (uint32_t)(read_off[r + 1] - off), rec % 6, sink,
                            scratch + 12 * off0 + 4 * (off - off0));
}

// One warp per kAggSlots groups.  Of a pair's twelve frame records only the one or two that
// carry hits need the seedextend machine (the lookup kernel left a 6-bit mask per read), so the
// warp first pools the live records of all its groups and runs their machines side by side, one
// lane each, 32 at a time, counting the kept ids in the groups' shared-memory tables; every group's
// distinct taxa (a dozen, typically) are then sorted in registers and aggregated by the whole warp.
__global__ void __launch_bounds__(kAggWarps * 32)
classify_kernel(TaxView tv, ClassifyParams cp, const uint32_t* __restrict__ ids,
                const uint64_t* __restrict__ read_off, const uint64_t* __restrict__ group_off,
                uint64_t g_begin, uint64_t ngroups /* end of the group range */,
                const uint8_t* __restrict__ frame_hits /* may be null: every record is live */,
                uint32_t* __restrict__ scratch, uint32_t* __restrict__ taxon_out, DevError* err) {
    __shared__ uint32_t s_key[kAggWarps][kAggSlots][kAggTable];
    __shared__ uint32_t s_occ[kAggWarps][kAggSlots][kAggTable];
    __shared__ uint32_t s_a[kAggWarps][kAggTable + 1];   // per-warp scratch of the aggregation
    __shared__ uint32_t s_c[kAggWarps][kAggTable + 1];
    __shared__ uint32_t s_p[kAggWarps][kAggTable + 1];
    __shared__ uint32_t s_l[kAggWarps][kAggTable + 1];
    __shared__ uint32_t s_cnt[kAggWarps][kAggSlots];       // distinct taxa per group (list length on the list path)
    __shared__ uint64_t s_r0[kAggWarps][kAggSlots];
    __shared__ uint32_t s_nrec[kAggWarps][kAggSlots + 1];  // exclusive prefix of the groups' record counts
    __shared__ uint32_t s_live[kAggWarps][64];             // pooled live records: slot << 24 | record
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1;
    const uint64_t nwarps = (uint64_t)gridDim.x * kAggWarps;
    const uint64_t nunits = (ngroups - g_begin + kAggSlots - 1) / kAggSlots;
    for (uint64_t unit = (uint64_t)blockIdx.x * kAggWarps + warp; unit < nunits; unit += nwarps) {
        const uint64_t g_base = g_begin + unit * kAggSlots;
        const int ng = (int)((ngroups - g_base) < (uint64_t)kAggSlots ? (ngroups - g_base) : kAggSlots);
        // group geometry: lane sl < ng loads group sl; groups of more than 2^24/6 reads are clamped
        // here and handled by the list path below (their records are enumerated there again)
        uint64_t my_r0 = 0, my_n = 0;
        if (lane < ng) {
            my_r0 = group_off[g_base + lane];
            my_n = (group_off[g_base + lane + 1] - my_r0) * 6;
        }
        const uint32_t my_n32 = my_n > 0xFFFFFFu ? 0xFFFFFFu : (uint32_t)my_n;
        const uint32_t incl = warp_incl_scan(lane < ng ? my_n32 : 0u, lane);
        if (lane < kAggSlots) {
            s_r0[warp][lane] = my_r0;
            s_nrec[warp][lane] = incl - (lane < ng ? my_n32 : 0u);
            s_cnt[warp][lane] = 0;
        }
        for (uint32_t i = lane; i < kAggSlots * kAggTable; i += 32) {
            (&s_key[warp][0][0])[i] = 0;
            (&s_occ[warp][0][0])[i] = 0;
        }
        const uint32_t total_recs = __shfl_sync(0xffffffffu, incl, 31);
        if (lane == 0) s_nrec[warp][kAggSlots] = total_recs;
        __syncwarp();
        unsigned present = 0;  // bit sl: group sl produced at least one record (prot2kmer2lca.rs:172)
        uint32_t nlive = 0;
        auto run_pass = [&](uint32_t count) {  // seedextend machines of the first `count` pooled records
            if ((uint32_t)lane < count) {
                const uint32_t e = s_live[warp][lane];
                const uint32_t sl = e >> 24, rec = e & 0xFFFFFFu;
                // a group that already overflowed its table is redone below: skip its remaining records
                if (s_cnt[warp][sl] <= kAggFull)
                    seedextend_record(cp, ids, read_off, s_r0[warp][sl], rec,
                                      TableSink{s_key[warp][sl], s_occ[warp][sl], &s_cnt[warp][sl]}, scratch);
            }
            __syncwarp();
        };
        for (uint32_t base = 0; base < total_recs; base += 32) {
            const uint32_t e = base + lane;
            bool exists = false, live = false;
            uint32_t sl = 0, rec = 0;
            if (e < total_recs) {
                while (sl + 1 < (uint32_t)ng && e >= s_nrec[warp][sl + 1]) ++sl;
                rec = e - s_nrec[warp][sl];
                const uint64_t r = s_r0[warp][sl] + rec / 6;
                const uint32_t fr = rec % 6, f = fr % 3;
                const uint32_t n = (uint32_t)(read_off[r + 1] - read_off[r]);
                exists = n >= f && (n - f) / 3 >= (uint32_t)cp.k;
                live = exists && (!frame_hits || (frame_hits[r] >> fr & 1));
            }
            for (int q = 0; q < ng; ++q)
                if (__ballot_sync(0xffffffffu, exists && sl == (uint32_t)q)) present |= 1u << q;
            const unsigned m = __ballot_sync(0xffffffffu, live);
            if (live) s_live[warp][nlive + __popc(m & lt_mask)] = (sl << 24) | rec;
            nlive += __popc(m);
            __syncwarp();
            if (nlive >= 32) {
                run_pass(32);
                const uint32_t rest = nlive - 32;  // < 32: slide the remainder to the front
                const uint32_t moved = (uint32_t)lane < rest ? s_live[warp][32 + lane] : 0u;
                __syncwarp();
                if ((uint32_t)lane < rest) s_live[warp][lane] = moved;
                nlive = rest;
                __syncwarp();
            }
        }
        if (nlive) run_pass(nlive);
        for (int sl = 0; sl < ng; ++sl) {
            const uint64_t r0 = s_r0[warp][sl];
            const uint64_t nrec = (group_off[g_base + sl + 1] - r0) * 6;
            const uint32_t distinct = s_cnt[warp][sl];
            uint32_t res, bad = 0;
            if (distinct > kAggFull || nrec > 0xFFFFFFu) {
                // rare: more distinct taxa than the table takes (or a huge group) -> redo into the group's own
                // slice of the global scratch (after the private record slices: four lists of gsize words;
                // the number of runs is below gsize because a read yields fewer k-mers than 2x its length)
                const uint64_t r1 = r0 + nrec / 6;
                const uint64_t gsize = 2 * (read_off[r1] - read_off[r0]);
                const uint64_t gbase = 12 * read_off[r0] + 2 * gsize;
                uint32_t* A = scratch + gbase;
                uint32_t* C = A + gsize;
                __syncwarp();
                if (lane == 0) s_cnt[warp][sl] = 0;
                __syncwarp();
                bool any = false;
                for (uint64_t rec = lane; rec < nrec; rec += 32) {
                    if (frame_hits) {  // a frame without hits contributes nothing (and, after sampled lookups, has no ids)
                        const uint64_t r = r0 + rec / 6;
                        const uint32_t fr = (uint32_t)(rec % 6), f = fr % 3;
                        if (!(frame_hits[r] >> fr & 1)) {
                            const uint32_t n = (uint32_t)(read_off[r + 1] - read_off[r]);
                            any |= n >= f && (n - f) / 3 >= (uint32_t)cp.k;
                            continue;
                        }
                    }
                    any |= seedextend_record(cp, ids, read_off, r0, (uint32_t)rec, ListSink{A, C, &s_cnt[warp][sl]}, scratch);
                }
                if (__any_sync(0xffffffffu, any)) present |= 1u << sl;
                __threadfence_block();
                __syncwarp();
                res = !(present >> sl & 1) ? UMGAP_ABSENT
                                           : warp_aggregate<true>(tv, A, C, C + gsize, C + 2 * gsize, s_cnt[warp][sl], cp.agg, lane, &bad);
            } else if (!(present >> sl & 1)) {
                res = UMGAP_ABSENT;
            } else {
                // compact the table: slots lane and lane + 32
                const uint32_t k0 = s_key[warp][sl][lane], k1 = s_key[warp][sl][lane + 32];
                const unsigned m0 = __ballot_sync(0xffffffffu, k0 != 0), m1 = __ballot_sync(0xffffffffu, k1 != 0);
                const uint32_t n0 = (uint32_t)__popc(m0), n = n0 + (uint32_t)__popc(m1);
                if (n <= 32) {
                    __syncwarp();
                    if (k0) {
                        const uint32_t at = (uint32_t)__popc(m0 & lt_mask);
                        s_a[warp][at] = k0;
                        s_c[warp][at] = s_occ[warp][sl][lane];
                    }
                    if (k1) {
                        const uint32_t at = n0 + (uint32_t)__popc(m1 & lt_mask);
                        s_a[warp][at] = k1;
                        s_c[warp][at] = s_occ[warp][sl][lane + 32];
                    }
                    __syncwarp();
                    const uint32_t id = (uint32_t)lane < n ? s_a[warp][lane] : 0u, c = (uint32_t)lane < n ? s_c[warp][lane] : 0u;
                    __syncwarp();
                    res = warp_aggregate_distinct(tv, id, c, n, s_a[warp], s_p[warp], s_l[warp], cp.agg, lane, &bad);
                } else {
                    __syncwarp();
                    if (k0) {
                        const uint32_t at = (uint32_t)__popc(m0 & lt_mask);
                        s_a[warp][at] = k0;
                        s_c[warp][at] = s_occ[warp][sl][lane];
                    }
                    if (k1) {
                        const uint32_t at = n0 + (uint32_t)__popc(m1 & lt_mask);
                        s_a[warp][at] = k1;
                        s_c[warp][at] = s_occ[warp][sl][lane + 32];
                    }
                    __syncwarp();
                    res = warp_aggregate<true>(tv, s_a[warp], s_c[warp], s_p[warp], s_l[warp], n, cp.agg, lane, &bad);
                }
            }
            if (res == kAggUnknown) {
                const uint32_t bad_any = __reduce_max_sync(0xffffffffu, bad);
                if (lane == 0 && atomicCAS(&err->flag, 0u, 1u) == 0u) err->taxon = bad_any;
                res = UMGAP_ABSENT;
            }
            if (lane == 0) taxon_out[g_base + sl] = res;
            __syncwarp();
        }
    }
}

// Chunk-relative offset slices: an uploaded slice has the chunk's first nucleotide / first read subtracted; a slice the
// host found to be an arithmetic progression (all reads of one length, all groups of one size) was not uploaded at all
// and is written here (step != 0).  add_a: a packed chunk starts inside its first 16-nucleotide word.
__global__ void rebase_kernel(uint64_t* a, uint64_t na, uint64_t base_a, uint64_t step_a, uint64_t add_a, uint64_t* b, uint64_t nb,
                              uint64_t base_b, uint64_t step_b) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < na; i += stride)
        a[i] = (step_a ? i * step_a : a[i] - base_a) + add_a;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nb; i += stride) b[i] = step_b ? i * step_b : b[i] - base_b;
}

}  // namespace umgap

using namespace umgap;

// ---- workspace slots of an index handle --------------------------------------------------------
constexpr int kMaxBufs = 6;  // chunk streams of the host-buffer path, each with its own set of buffers
enum { WS_ERR = 0, WS_NT = 1, WS_ROFF = WS_NT + kMaxBufs, WS_GOFF = WS_ROFF + kMaxBufs, WS_OUT = WS_GOFF + kMaxBufs, WS_HITS = WS_OUT + kMaxBufs,
       WS_IDS = WS_HITS + kMaxBufs, WS_SCRATCH = WS_IDS + kMaxBufs, WS_LONG = WS_SCRATCH + kMaxBufs,
       WS_ERRQ = WS_LONG + kMaxBufs, WS_END = WS_ERRQ + 1 };
constexpr uint32_t kErrRing = 64;  // device error slots of the batches in flight (umgap_classify_reads*_async)
constexpr uint32_t kMaxPending = 32;
static_assert(WS_END <= Workspace::kSlots, "workspace slots");

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

// ---- optional per-launch timing (umgap_kernel_timing), shared with route.cu -----------------------
namespace umgap {
// Process-wide counters and the timing aid; several host threads may drive different GPUs at once
// (umgap_classify_reads_multi), so the counters are atomic and the event lists are guarded.
static std::atomic<uint64_t> g_launch_count{0};  // kernels launched by the fused path
static std::atomic<uint64_t> g_h2d_bytes{0}, g_d2h_bytes{0};  // bytes umgap_classify_reads moved over PCIe
static bool g_sampling = getenv("UMGAP_NO_SAMPLING") == nullptr;  // sampled lookups in front of seedextend (umgap_pipeline_sampling)
static int g_slices = [] {           // slices of the device-buffer entry point (umgap_pipeline_slices)
    const char* e = getenv("UMGAP_SLICES");
    const int v = e ? atoi(e) : 0;
    return v > 0 ? std::min(v, 64) : 1;
}();
bool g_timing = false;
std::vector<TimedLaunch> g_launches;
static std::mutex g_timing_mu;
static std::map<int, std::vector<cudaEvent_t>> g_event_pool;  // per device: an event records only on its own device's streams
static cudaEvent_t take_event(int dev) {
    {
        std::lock_guard<std::mutex> lk(g_timing_mu);
        std::vector<cudaEvent_t>& pool = g_event_pool[dev];
        if (!pool.empty()) {
            cudaEvent_t e = pool.back();
            pool.pop_back();
            return e;
        }
    }
    cudaEvent_t e;
    UMGAP_CUDA(cudaEventCreate(&e));
    return e;
}
LaunchTimer::LaunchTimer(int kind, cudaStream_t s, bool enabled) : st(s), on(g_timing && enabled) {
    if (!on) return;
    UMGAP_CUDA(cudaGetDevice(&t.dev));
    t.kind = kind;
    t.a = take_event(t.dev);
    t.b = take_event(t.dev);
    UMGAP_CUDA(cudaEventRecord(t.a, st));
}
void LaunchTimer::cancel() {
    if (!on) return;
    std::lock_guard<std::mutex> lk(g_timing_mu);
    g_event_pool[t.dev].push_back(t.a);
    g_event_pool[t.dev].push_back(t.b);
    on = false;
}
void LaunchTimer::stop() {
    if (!on) return;
    UMGAP_CUDA(cudaEventRecord(t.b, st));
    std::lock_guard<std::mutex> lk(g_timing_mu);
    g_launches.push_back(t);
}
}  // namespace umgap

// Level 0 of a table larger than this is probed one hash-prefix region per launch.  Random gathers keep the full
// line rate up to a 64 GiB footprint and lose half of it at 80 GiB (profiles/r01_randsector_sweep.log,
// r01_vmm_pages_probe.log); the batch buffers need their share of the reach, and every pass costs a translation
// of the batch, so the regions are as large as that allows (128 GB table: 2 regions 8.3 ms, 3 regions 10.4 ms).
constexpr uint64_t kDefaultRegionBytes = 60ull << 30;

// Lookup launch over reads [r_begin, r_end).
static void launch_translate_lookup(const umgap_index* idx, const umgap_pipeline_opts* o,
                                    const NtSrc& nv, const uint64_t* read_off_dev, uint64_t r_begin,
                                    uint64_t r_end, uint32_t* ids_dev, uint8_t* frame_hits_dev, cudaStream_t st,
                                    ReadList rl = ReadList(), bool timed = true) {
    if (r_end <= r_begin) return;
    CodonLut lut{};
    make_code_lut(idx, o->table, o->methionine, lut);
    const unsigned blocks = rl.list ? 148u : (unsigned)std::min<uint64_t>(ceil_div(r_end - r_begin, kLookupWarps), 148ull * 32);
    if (idx->nshards > 1 && !idx->attached)
        UMGAP_FAIL(UMGAP_ERR_INVALID, "sharded index: call umgap_index_attach_shards() before looking up");
    // Random probes over more than ~64 GiB collapse to a quarter of the line rate on B200 (address-
    // translation reach, profiles/r01_randsector_sweep.log) while any 64 GiB window runs at full rate:
    // a larger level 0 is probed one hash-prefix region at a time (the line index is monotone in the
    // prefix).  Every pass translates and packs all k-mers again, which costs less than the cliff.
    const uint64_t region_bytes = idx->region_bytes ? idx->region_bytes : kDefaultRegionBytes;
    const uint64_t table_bytes = (uint64_t)idx->level_nlines[0] * 128;
    const int nregions = idx->nshards > 1 ? 1 : (int)std::max<uint64_t>(1, ceil_div(table_bytes, region_bytes));
    for (int reg = 0; reg < nregions; ++reg) {
        const uint64_t lo = (1ull << 32) * reg / nregions, hi = (1ull << 32) * (reg + 1) / nregions;
        LaunchTimer timer(0, st, timed);
        switch (idx->k) {
#define UMGAP_TL(KK, VIEW, VIEWARG, REG)                                                                     \
    if (nv.codes)                                                                                            \
        translate_lookup_kernel<KK, VIEW, REG, true><<<blocks, kLookupWarps * 32, 0, st>>>(                   \
            VIEWARG, lut, nv, read_off_dev, r_begin, r_end, ids_dev, frame_hits_dev, lo, hi, rl);            \
    else                                                                                                     \
        translate_lookup_kernel<KK, VIEW, REG, false><<<blocks, kLookupWarps * 32, 0, st>>>(                  \
            VIEWARG, lut, nv, read_off_dev, r_begin, r_end, ids_dev, frame_hits_dev, lo, hi, rl)
#define UMGAP_CASE(KK)                                                                                       \
    case KK:                                                                                                 \
        if (idx->nshards > 1) {                                                                              \
            UMGAP_TL(KK, ShardedView, idx->sharded, false);                                                  \
        } else if (nregions > 1) {                                                                           \
            UMGAP_TL(KK, TableView, idx->view(), true);                                                      \
        } else {                                                                                             \
            UMGAP_TL(KK, TableView, idx->view(), false);                                                     \
        }                                                                                                    \
        break;
            UMGAP_CASE(1) UMGAP_CASE(2) UMGAP_CASE(3) UMGAP_CASE(4) UMGAP_CASE(5) UMGAP_CASE(6)
            UMGAP_CASE(7) UMGAP_CASE(8) UMGAP_CASE(9)
#undef UMGAP_CASE
#undef UMGAP_TL
            default:
                UMGAP_FAIL(UMGAP_ERR_INVALID, "unsupported k %d", idx->k);
        }
        UMGAP_CUDA(cudaGetLastError());
        ++g_launch_count;
        timer.stop();
    }
}

static void launch_classify(const umgap_index* idx, const umgap_taxonomy* tax,
                            const umgap_pipeline_opts* o, const uint32_t* ids_dev,
                            const uint64_t* read_off_dev, const uint64_t* group_off_dev, uint64_t g_begin,
                            uint64_t g_end, const uint8_t* frame_hits_dev, uint32_t* scratch_dev, uint32_t* out_dev,
                            DevError* err, cudaStream_t st, bool frame_major = false) {
    if (g_end <= g_begin) return;
    const unsigned blocks = (unsigned)std::min<uint64_t>(ceil_div(ceil_div(g_end - g_begin, kAggSlots), kAggWarps), 148ull * 64);
    LaunchTimer timer(1, st);
    ClassifyParams cp = make_params(idx, o);
    cp.frame_major = frame_major ? 1 : 0;
    classify_kernel<<<blocks, kAggWarps * 32, 0, st>>>(tax->view, cp, ids_dev,
                                                       read_off_dev, group_off_dev, g_begin, g_end,
                                                       frame_hits_dev, scratch_dev, out_dev, err);
    UMGAP_CUDA(cudaGetLastError());
    ++g_launch_count;
    timer.stop();
}

// The whole device-side classification of one batch: lookup kernel, then classify kernel.
// (Running the classify kernel of one slice concurrently with the lookup kernel of the next, on two
// streams with priorities, was measured and gives nothing: both kernels want the same registers
// and issue slots -- profiles/README.md.)
struct SampledPlan {  // the sampled lookup stage of one batch
    int stride = 0;       // 0: not applicable, use the plain kernel
    CodonLut lut{};
    uint64_t total_nt = 0;
    uint32_t* long_count = nullptr;  // 64 counters (one per slice), then the list of reads left to the plain kernel
    uint32_t* long_list = nullptr;
    int nregions = 1;     // > 1: level 0 exceeds the probe region, phases as separate launches per hash-prefix region
};

// reads_hint: number of reads the launch will find in its group range (sizes the grid only).
static void launch_sampled(const umgap_index* idx, const umgap_pipeline_opts* o, const SampledPlan& sp, const NtSrc& nv,
                           const uint64_t* read_off_dev, uint64_t nreads, uint64_t reads_hint, uint32_t* ids_dev,
                           uint8_t* frame_hits_dev, const uint64_t* group_off_dev, uint64_t g_lo, uint64_t g_hi, int slice,
                           cudaStream_t st) {
    static const unsigned grid_cap = [] {  // CTAs of one launch (UMGAP_S_GRID = CTAs per SM, for measurements)
        const char* e = getenv("UMGAP_S_GRID");
        const int v = e ? atoi(e) : 0;
        return 148u * (unsigned)(v > 0 ? v : kSBlocks);  // one resident wave: the units are handed out dynamically
    }();
    const unsigned blocks = (unsigned)std::min<uint64_t>(ceil_div(ceil_div(reads_hint, kSReads), kSWarps) + 1, grid_cap);
#define UMGAP_SAMPLED_P(S, MODE, COUNTER, LO, HI, PK)                                                                          \
    lookup_sampled_kernel<9, TableView, S, MODE, PK><<<blocks, kSWarps * 32, 0, st>>>(                                          \
        idx->view(), sp.lut, nv, sp.total_nt, read_off_dev, (uint32_t)nreads, ids_dev, frame_hits_dev, group_off_dev, g_lo,      \
        g_hi, sp.long_list, sp.long_count + slice, sp.long_count + 64 + (COUNTER), LO, HI, RouteSink())
#define UMGAP_SAMPLED(S, MODE, COUNTER, LO, HI)                                                                                \
    if (nv.codes) UMGAP_SAMPLED_P(S, MODE, COUNTER, LO, HI, true); else UMGAP_SAMPLED_P(S, MODE, COUNTER, LO, HI, false)
#define UMGAP_SAMPLED_STRIDES(MODE, COUNTER, LO, HI)                 \
    switch (sp.stride) {                                             \
        case 2: UMGAP_SAMPLED(2, MODE, COUNTER, LO, HI); break;      \
        case 3: UMGAP_SAMPLED(3, MODE, COUNTER, LO, HI); break;      \
        default: UMGAP_SAMPLED(4, MODE, COUNTER, LO, HI); break;     \
    }                                                                \
    UMGAP_CUDA(cudaGetLastError());                                  \
    ++g_launch_count
    if (sp.nregions <= 1) {
        UMGAP_SAMPLED_STRIDES(0, slice, 0ull, 1ull << 32);
    } else {
        // phase 1 of every region (frame masks complete after the last), then phase 2 of every region; a launch
        // confines its probes to one region of level 0 (address-translation reach, launch_translate_lookup)
        for (int ph = 1; ph <= 2; ++ph)
            for (int reg = 0; reg < sp.nregions; ++reg) {
                const uint64_t lo = (1ull << 32) * reg / sp.nregions, hi = (1ull << 32) * (reg + 1) / sp.nregions;
                const int counter = (ph - 1) * sp.nregions + reg;
                if (ph == 1) {
                    UMGAP_SAMPLED_STRIDES(1, counter, lo, hi);
                } else {
                    UMGAP_SAMPLED_STRIDES(2, counter, lo, hi);
                }
            }
    }
#undef UMGAP_SAMPLED_STRIDES
#undef UMGAP_SAMPLED
#undef UMGAP_SAMPLED_P
    // the reads the kernel queued (longer than a warp batch; rare): every position, plain kernel over the list
    ReadList rl;
    rl.list = sp.long_list;
    rl.count = sp.long_count + slice;
    rl.group_off = group_off_dev;
    rl.g_lo = g_lo;
    launch_translate_lookup(idx, o, nv, read_off_dev, 0, nreads, ids_dev, frame_hits_dev, st, rl, false);
}

// Sampled lookups (see lookup_sampled_kernel): valid only in front of seedextend with -o and S >= 2; only
// the frames flagged in frame_hits_dev have ids afterwards, which is all the classify kernel reads.
// Decides whether the stage applies and, if so, prepares what a batch needs once: the codon table and the
// work list through which reads longer than a warp batch reach the plain kernel.
static SampledPlan prepare_sampled(const umgap_index* idx, const umgap_pipeline_opts* o, const NtSrc& nv,
                                   const uint64_t* read_off_dev, uint64_t nreads, uint64_t total_nt, uint32_t* ids_dev,
                                   uint8_t* frame_hits_dev, cudaStream_t st, int buf) {
    SampledPlan sp;
    const bool disabled = !g_sampling;
    const uint64_t region_bytes = idx->region_bytes ? idx->region_bytes : kDefaultRegionBytes;
    const uint64_t nregions = std::max<uint64_t>(1, ceil_div((uint64_t)idx->level_nlines[0] * 128, region_bytes));
    if (disabled || !o->seedextend || !o->one_on_one || o->min_seed_size < 2 || idx->k != 9 || idx->nshards > 1 ||
        nregions > 32 || !frame_hits_dev || nreads >= (1ull << 31) || (!nv.codes && ((uintptr_t)nv.bytes & 15u) != 0))
        return sp;
    sp.stride = std::min(o->min_seed_size, 4);
    sp.nregions = (int)nregions;
    if (!nreads) return sp;
    CodonLut lut{};
    make_code_lut(idx, o->table, o->methionine, lut);
    sp.total_nt = total_nt;
    sp.lut = lut;
    // 64 long-read counters and 64 unit counters (one of each per slice), then the long-read list
    sp.long_count = (uint32_t*)idx->ws.get(WS_LONG + buf, (128 + nreads) * sizeof(uint32_t));
    sp.long_list = sp.long_count + 128;
    UMGAP_CUDA(cudaMemsetAsync(sp.long_count, 0, 128 * sizeof(uint32_t), st));
    return sp;
}

static void launch_pipeline(const umgap_index* idx, const umgap_taxonomy* tax, const umgap_pipeline_opts* o,
                            const NtSrc& nv, const uint64_t* read_off_dev, uint64_t nreads, uint64_t total_nt,
                            const uint64_t* group_off_dev, uint64_t ngroups, uint32_t* ids_dev, uint32_t* scratch_dev,
                            uint8_t* frame_hits_dev, uint32_t* out_dev, DevError* err, cudaStream_t st, int buf = 0,
                            bool sliced = false) {
    if (!ngroups) return;
    const int kSlices = g_slices;
    LaunchTimer timer(0, st);
    const SampledPlan sp = prepare_sampled(idx, o, nv, read_off_dev, nreads, total_nt, ids_dev, frame_hits_dev, st, buf);
    const bool slice_it = sliced && kSlices >= 2 && ngroups >= 4096u * (uint64_t)kSlices && nreads && sp.nregions == 1;
    if (!sp.stride) {
        timer.cancel();  // the plain launch brackets itself
        launch_translate_lookup(idx, o, nv, read_off_dev, 0, nreads, ids_dev, frame_hits_dev, st);
        launch_classify(idx, tax, o, ids_dev, read_off_dev, group_off_dev, 0, ngroups, frame_hits_dev, scratch_dev, out_dev, err, st);
        return;
    }
    if (!slice_it) {
        if (nreads) launch_sampled(idx, o, sp, nv, read_off_dev, nreads, nreads, ids_dev, frame_hits_dev, nullptr, 0, 0, 0, st);
        timer.stop();
        launch_classify(idx, tax, o, ids_dev, read_off_dev, group_off_dev, 0, ngroups, frame_hits_dev, scratch_dev, out_dev, err, st, true);
        return;
    }
    // Sliced: the groups are cut into kSlices ranges; the lookup and classify kernels of a slice follow each other on
    // one of two internal streams (even / odd slices), so the classify kernel of slice i runs beside the lookup kernel
    // of slice i + 1 and never more than two slices are in flight.  (One stream of back-to-back lookup kernels plus a
    // higher-priority stream of classify kernels was measured and is slower: 7.1 vs 6.6 ms, the lookup kernels run
    // ahead and the classify kernels pile up at the end.)  The lookup kernel takes its read range from the group table on the
    // device; every buffer is addressed by absolute read / nucleotide position, so the slices share them.
    timer.stop();  // the bracket of this mode covers the once-per-batch part only; the slices are bracketed one by one
    if (!idx->aux_fork) {
        for (int i = 0; i < 2; ++i) {
            UMGAP_CUDA(cudaStreamCreateWithFlags(&idx->aux_stream[i], cudaStreamNonBlocking));
            UMGAP_CUDA(cudaEventCreateWithFlags(&idx->aux_join[i], cudaEventDisableTiming));
        }
        UMGAP_CUDA(cudaEventCreateWithFlags(&idx->aux_fork, cudaEventDisableTiming));
    }
    UMGAP_CUDA(cudaEventRecord(idx->aux_fork, st));
    for (int i = 0; i < 2; ++i) UMGAP_CUDA(cudaStreamWaitEvent(idx->aux_stream[i], idx->aux_fork, 0));
    for (int sl = 0; sl < kSlices; ++sl) {
        const uint64_t g_lo = ngroups * sl / kSlices, g_hi = ngroups * (sl + 1) / kSlices;
        cudaStream_t s = idx->aux_stream[sl & 1];
        {
            LaunchTimer t2(0, s);
            launch_sampled(idx, o, sp, nv, read_off_dev, nreads, ceil_div(nreads, kSlices), ids_dev, frame_hits_dev, group_off_dev,
                           g_lo, g_hi, sl, s);
            t2.stop();
        }
        launch_classify(idx, tax, o, ids_dev, read_off_dev, group_off_dev, g_lo, g_hi, frame_hits_dev, scratch_dev, out_dev, err, s, true);
    }
    for (int i = 0; i < 2; ++i) {
        UMGAP_CUDA(cudaEventRecord(idx->aux_join[i], idx->aux_stream[i]));
        UMGAP_CUDA(cudaStreamWaitEvent(st, idx->aux_join[i], 0));
    }
}

// Common difference of off[0..n] when it is an arithmetic progression with a non-zero step, else 0.
static uint64_t uniform_step(const uint64_t* off, uint64_t n) {
    static const bool disabled = getenv("UMGAP_UPLOAD_OFFSETS") != nullptr;
    if (n == 0 || disabled) return 0;
    const uint64_t step = off[1] - off[0];
    if (!step || off[n] - off[0] != n * step) return 0;
    for (uint64_t i0 = 1; i0 < n; i0 += 4096) {  // blockwise so that the inner loop vectorises
        const uint64_t i1 = std::min(n, i0 + 4096);
        uint64_t bad = 0;
        for (uint64_t i = i0; i < i1; ++i) bad |= (off[i + 1] - off[i]) ^ step;
        if (bad) return 0;
    }
    return step;
}

static void raise_dev_error(const DevError& e) {
    if (e.flag) UMGAP_FAIL(UMGAP_ERR_UNKNOWN_TAXON, "Unknown Taxon ID: %u", e.taxon);
}

namespace umgap {
// The workspaces the pack and classify kernels of a batch of this size take from the handle, allocated now: a launch
// path that allocates synchronises the device, which the exchange step (exchange.cu) must never do inside a batch.
void pipeline_reserve(const umgap_index* idx, uint64_t nreads, uint64_t total_nt, int buf) {
    use_device(idx->device);
    idx->ws.get(WS_SCRATCH + buf, (12 * total_nt + 64) * sizeof(uint32_t));
    UMGAP_CUDA(cudaMemset(idx->ws.get(WS_ERR, sizeof(DevError)), 0, sizeof(DevError)));
    idx->ws.get(WS_LONG + buf, (128 + nreads) * sizeof(uint32_t));
}
// umgap_classify_ids[_masked]_dev with the workspace set `buf` (0 .. kMaxBufs - 1): batches in flight at the same time
// on one handle (the lanes of the exchange step) must not share the scratch of the groups with many distinct taxa.
void classify_ids_buf(const umgap_index* idx, const umgap_taxonomy* tax, const umgap_pipeline_opts* opts, const uint32_t* ids_dev,
                      const uint64_t* read_off_dev, uint64_t total_nt, const uint64_t* group_off_dev, uint64_t ngroups,
                      const uint8_t* frame_hits_dev, bool frame_major, uint32_t* taxon_out_dev, int buf, cudaStream_t st) {
    check_opts(idx, tax, opts);
    if (!tax || !ids_dev) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
    if (buf < 0 || buf >= kMaxBufs) UMGAP_FAIL(UMGAP_ERR_INVALID, "workspace set out of range");
    use_device(idx->device);
    uint32_t* scratch = (uint32_t*)idx->ws.get(WS_SCRATCH + buf, (12 * total_nt + 64) * sizeof(uint32_t));
    DevError* err = (DevError*)idx->ws.get(WS_ERR, sizeof(DevError));
    launch_classify(idx, tax, opts, ids_dev, read_off_dev, group_off_dev, 0, ngroups, frame_hits_dev, scratch, taxon_out_dev, err, st,
                    frame_major);
}
// Raises the error the classify kernel of the last *_dev call left behind (Unknown Taxon ID); the caller has synchronised.
void pipeline_take_error(const umgap_index* idx) {
    use_device(idx->device);
    DevError he;
    void* err = idx->ws.get(WS_ERR, sizeof(DevError));
    UMGAP_CUDA(cudaMemcpy(&he, err, sizeof he, cudaMemcpyDeviceToHost));
    UMGAP_CUDA(cudaMemset(err, 0, sizeof(DevError)));
    raise_dev_error(he);
}
}  // namespace umgap

extern "C" {

int umgap_index_set_probe_region(umgap_index* idx, uint64_t bytes) {
    return guarded([&] {
        if (!idx) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        idx->region_bytes = bytes;
    });
}

int umgap_pipeline_sampling(int enable) {
    const int before = g_sampling ? 1 : 0;
    if (enable >= 0) g_sampling = enable != 0;
    return before;
}

int umgap_pipeline_slices(int slices) {
    const int before = g_slices;
    if (slices > 0) g_slices = std::min(slices, 64);
    return before;
}

int umgap_transfer_bytes(uint64_t* h2d, uint64_t* d2h) {
    if (h2d) *h2d = g_h2d_bytes;
    if (d2h) *d2h = g_d2h_bytes;
    return UMGAP_OK;
}

int umgap_kernel_launch_count(uint64_t* launches) {
    if (launches) *launches = g_launch_count;
    return UMGAP_OK;
}

int umgap_kernel_timing(int enable) {
    g_timing = enable != 0;
    return UMGAP_OK;
}

int umgap_kernel_times_ex(double* ms_out, uint64_t* launches_out, int nkinds) {
    return guarded([&] {
        double ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        uint64_t cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        std::lock_guard<std::mutex> lk(g_timing_mu);
        for (TimedLaunch& t : g_launches) {
            UMGAP_CUDA(cudaEventSynchronize(t.b));
            float e = 0;
            UMGAP_CUDA(cudaEventElapsedTime(&e, t.a, t.b));
            ms[t.kind & 7] += e;
            cnt[t.kind & 7]++;
            g_event_pool[t.dev].push_back(t.a);
            g_event_pool[t.dev].push_back(t.b);
        }
        g_launches.clear();
        for (int i = 0; i < nkinds && i < 8; ++i) {
            if (ms_out) ms_out[i] = ms[i];
            if (launches_out) launches_out[i] = cnt[i];
        }
    });
}

int umgap_kernel_times(double* lookup_ms, uint64_t* lookup_launches, double* classify_ms,
                       uint64_t* classify_launches) {
    double ms[8];
    uint64_t cnt[8];
    const int rc = umgap_kernel_times_ex(ms, cnt, 8);
    if (rc != UMGAP_OK) return rc;
    if (lookup_ms) *lookup_ms = ms[0];
    if (lookup_launches) *lookup_launches = cnt[0];
    if (classify_ms) *classify_ms = ms[1];
    if (classify_launches) *classify_launches = cnt[1];
    return UMGAP_OK;
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
        NtSrc nv;
        nv.bytes = nt_dev;
        launch_translate_lookup(idx, opts, nv, read_off_dev, 0, nreads, ids_dev, nullptr, (cudaStream_t)stream);
    });
}

int umgap_classify_ids_dev(const umgap_index* idx, const umgap_taxonomy* tax, const umgap_pipeline_opts* opts,
                           const uint32_t* ids_dev, const uint64_t* read_off_dev, uint64_t total_nt,
                           const uint64_t* group_off_dev, uint64_t ngroups, uint32_t* taxon_out_dev, void* stream) {
    return guarded([&] {
        check_opts(idx, tax, opts);
        if (!tax || !ids_dev) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        use_device(idx->device);
        cudaStream_t st = (cudaStream_t)stream;
        uint32_t* scratch = (uint32_t*)idx->ws.get(WS_SCRATCH, (12 * total_nt + 64) * sizeof(uint32_t));
        DevError* err = (DevError*)idx->ws.get(WS_ERR, sizeof(DevError));
        UMGAP_CUDA(cudaMemsetAsync(err, 0, sizeof(DevError), st));
        launch_classify(idx, tax, opts, ids_dev, read_off_dev, group_off_dev, 0, ngroups, nullptr, scratch, taxon_out_dev, err, st);
    });
}

int umgap_classify_ids_masked_dev(const umgap_index* idx, const umgap_taxonomy* tax, const umgap_pipeline_opts* opts,
                                  const uint32_t* ids_dev, const uint64_t* read_off_dev, uint64_t total_nt,
                                  const uint64_t* group_off_dev, uint64_t ngroups, const uint8_t* frame_hits_dev,
                                  int frame_major, uint32_t* taxon_out_dev, void* stream) {
    return guarded([&] {
        check_opts(idx, tax, opts);
        if (!tax || !ids_dev) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        use_device(idx->device);
        cudaStream_t st = (cudaStream_t)stream;
        uint32_t* scratch = (uint32_t*)idx->ws.get(WS_SCRATCH, (12 * total_nt + 64) * sizeof(uint32_t));
        DevError* err = (DevError*)idx->ws.get(WS_ERR, sizeof(DevError));
        UMGAP_CUDA(cudaMemsetAsync(err, 0, sizeof(DevError), st));
        launch_classify(idx, tax, opts, ids_dev, read_off_dev, group_off_dev, 0, ngroups, frame_hits_dev, scratch, taxon_out_dev, err, st,
                        frame_major != 0);
    });
}

extern "C++" {
namespace umgap {
void launch_route_pack_list(const umgap_index* idx, const umgap_pipeline_opts* opts, const uint8_t* nt_dev,
                            const uint64_t* read_off_dev, uint64_t cap, const BucketPtrs& hp, uint32_t* send_pos_dev,
                            uint64_t* cursors_dev, uint32_t* ids_dev, const uint32_t* list, const uint32_t* list_count,
                            const uint64_t* group_off_dev, uint64_t g_lo, cudaStream_t st);  // route.cu
}
}

int umgap_route_sampled_applies(const umgap_index* idx, const umgap_pipeline_opts* o) {
    return idx && o && g_sampling && o->seedextend && o->one_on_one && o->min_seed_size >= 2 && idx->k == 9 ? 1 : 0;
}

extern "C++" void umgap::route_pack_sampled(const umgap_index* idx, const umgap_pipeline_opts* opts, int phase, const uint8_t* nt_dev,
                                 const uint64_t* read_off_dev, uint64_t nreads, uint64_t total_nt, uint64_t cap,
                                 const BucketPtrs& hp, uint32_t* send_pos_dev, uint64_t* cursors_dev,
                                 uint8_t* frame_hits_dev, uint32_t* ids_dev, const uint64_t* group_off_dev, uint64_t g_lo,
                                 uint64_t g_hi, int slot, cudaStream_t st) {
    {
        if (!idx || !opts || !send_pos_dev || !cursors_dev || !ids_dev || !frame_hits_dev)
            UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        if (idx->nshards > kMaxShards) UMGAP_FAIL(UMGAP_ERR_INVALID, "the exchange step supports at most %d shards", kMaxShards);
        if (!umgap_route_sampled_applies(idx, opts))
            UMGAP_FAIL(UMGAP_ERR_INVALID, "sampled routing needs k = 9, -o and seedextend -s >= 2 (umgap_route_sampled_applies)");
        if (phase != 1 && phase != 2) UMGAP_FAIL(UMGAP_ERR_INVALID, "phase must be 1 or 2");
        if (slot < 0 || slot >= kMaxBufs) UMGAP_FAIL(UMGAP_ERR_INVALID, "slot must be in [0, %d)", kMaxBufs);
        if (2 * total_nt >= (1ull << 32) || nreads >= (1ull << 28)) UMGAP_FAIL(UMGAP_ERR_INVALID, "batch too large for 32-bit positions");
        if (((uintptr_t)nt_dev & 15u) || ((uintptr_t)frame_hits_dev & 3u)) UMGAP_FAIL(UMGAP_ERR_INVALID, "nt_dev must be 16-byte, frame_hits_dev 4-byte aligned");
        use_device(idx->device);
        UMGAP_CUDA(cudaMemsetAsync(cursors_dev, 0, 2 * (size_t)idx->nshards * sizeof(uint64_t), st));
        // 64 long-read counters, 64 unit counters, then the list of the reads longer than a warp batch (filled by phase 1)
        uint32_t* counters = (uint32_t*)idx->ws.get(WS_LONG + slot, (128 + nreads) * sizeof(uint32_t));
        if (phase == 1) {
            UMGAP_CUDA(cudaMemsetAsync(counters, 0, 128 * sizeof(uint32_t), st));
            // a group range shares frame_hits_dev with the other ranges of the batch: the caller cleared it
            if (!group_off_dev) UMGAP_CUDA(cudaMemsetAsync(frame_hits_dev, 0, (nreads + 3) / 4 * 4, st));
        } else {
            UMGAP_CUDA(cudaMemsetAsync(counters + 64, 0, 64 * sizeof(uint32_t), st));
        }
        if (!nreads) return;
        CodonLut lut{};
        make_code_lut(idx, opts->table, opts->methionine, lut);
        RouteSink rs;
        rs.h = hp;
        rs.send_pos = send_pos_dev;
        rs.cursors = reinterpret_cast<unsigned long long*>(cursors_dev);
        rs.cap = cap;
        rs.nshards = (uint32_t)idx->nshards;
        const unsigned blocks = (unsigned)std::min<uint64_t>(ceil_div(ceil_div(nreads, kSReads), kSWarps) + 1, 148ull * kSBlocks);
        const int stride = std::min(opts->min_seed_size, 4);
        NtSrc nv;
        nv.bytes = nt_dev;
#define UMGAP_ROUTE_SAMPLED(S, MODE)                                                                                          \
    lookup_sampled_kernel<9, TableView, S, MODE, false><<<blocks, kSWarps * 32, 0, st>>>(                                       \
        idx->view(), lut, nv, total_nt, read_off_dev, (uint32_t)nreads, ids_dev, frame_hits_dev, group_off_dev, g_lo, g_hi, counters + 128, \
        counters, counters + 64, 0ull, 1ull << 32, rs)
        if (phase == 1) {
            switch (stride) {
                case 2: UMGAP_ROUTE_SAMPLED(2, 3); break;
                case 3: UMGAP_ROUTE_SAMPLED(3, 3); break;
                default: UMGAP_ROUTE_SAMPLED(4, 3); break;
            }
        } else {
            switch (stride) {
                case 2: UMGAP_ROUTE_SAMPLED(2, 4); break;
                case 3: UMGAP_ROUTE_SAMPLED(3, 4); break;
                default: UMGAP_ROUTE_SAMPLED(4, 4); break;
            }
        }
#undef UMGAP_ROUTE_SAMPLED
        UMGAP_CUDA(cudaGetLastError());
        ++g_launch_count;
        if (phase == 2) {  // every position of the reads longer than a warp batch
            launch_route_pack_list(idx, opts, nt_dev, read_off_dev, cap, hp, send_pos_dev, cursors_dev, ids_dev, counters + 128,
                                   counters, group_off_dev, g_lo, st);
            ++g_launch_count;
        }
    }
}

int umgap_route_pack_sampled_dev(const umgap_index* idx, const umgap_pipeline_opts* opts, int phase, const uint8_t* nt_dev,
                                 const uint64_t* read_off_dev, uint64_t nreads, uint64_t total_nt, uint64_t cap,
                                 uint64_t* send_h_dev, uint32_t* send_pos_dev, uint64_t* cursors_dev,
                                 uint8_t* frame_hits_dev, uint32_t* ids_dev, const uint64_t* group_off_dev, uint64_t g_lo,
                                 uint64_t g_hi, int slot, void* stream) {
    return guarded([&] {
        if (!idx || !send_h_dev) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        if (idx->nshards > kMaxShards) UMGAP_FAIL(UMGAP_ERR_INVALID, "the exchange step supports at most %d shards", kMaxShards);
        BucketPtrs hp{};
        for (int o = 0; o < idx->nshards; ++o) hp.p[o] = send_h_dev + (uint64_t)o * cap;  // the buckets of this rank, sent by the host
        route_pack_sampled(idx, opts, phase, nt_dev, read_off_dev, nreads, total_nt, cap, hp, send_pos_dev, cursors_dev, frame_hits_dev,
                           ids_dev, group_off_dev, g_lo, g_hi, slot, (cudaStream_t)stream);
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
        uint32_t* scratch = (uint32_t*)idx->ws.get(WS_SCRATCH, (12 * total_nt + 64) * sizeof(uint32_t));
        DevError* err = (DevError*)idx->ws.get(WS_ERR, sizeof(DevError));
        UMGAP_CUDA(cudaMemsetAsync(err, 0, sizeof(DevError), st));
        uint8_t* hits = (uint8_t*)idx->ws.get(WS_HITS, nreads + 64);
        NtSrc nv;
        nv.bytes = nt_dev;
        launch_pipeline(idx, tax, opts, nv, read_off_dev, nreads, total_nt, group_off_dev, ngroups, ids, scratch, hits,
                        taxon_out_dev, err, st, 0, true);
    });
}

extern "C++" {
namespace {
// N flags of a packed chunk that travel as (word index << 16 | flags) entries when few words hold an N.
__global__ void nmask_scatter_kernel(uint16_t* nmask, const uint64_t* entries, uint64_t n, uint64_t first_word) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) nmask[(entries[i] >> 16) - first_word] = (uint16_t)(entries[i] & 0xFFFFu);
}

struct HostReads {  // one of the two host forms of a batch's nucleotides
    const uint8_t* nt = nullptr;       // bytes as received
    const uint32_t* codes = nullptr;   // packed: 2 bits per nucleotide ...
    const uint64_t* n_entries = nullptr;  // ... and the words that hold an N: word index << 16 | flags, ascending
    uint64_t n_count = 0;
};

}  // namespace

}  // extern "C++"
// A batch enqueued by umgap_classify_reads[_packed]_async: one event per chunk stream behind its last operation, and
// the device slot its classify kernels report an unknown taxon to.
struct umgap_pending {
    const umgap_index* idx = nullptr;
    cudaEvent_t ev[kMaxBufs + 1] = {};
    int nev = 0;
    DevError* err = nullptr;       // the batch's device slot
    DevError* err_host = nullptr;  // its page-locked mirror, valid behind the last event
};
extern "C++" {

// Groups [g_begin, g_end) of the batch (the whole batch by default; umgap_classify_reads_multi hands every GPU its range).
// pend != nullptr: the call returns once everything is enqueued (umgap_pending_wait completes it).
static void classify_host(const umgap_index* idx, const umgap_taxonomy* tax, const umgap_pipeline_opts* opts,
                          const HostReads& hr, const uint64_t* read_off, uint64_t nreads, const uint64_t* group_off,
                          uint64_t ngroups, uint32_t* taxon_out, uint64_t* n_lookups, uint64_t g_begin = 0,
                          uint64_t g_end = ~0ull, umgap_pending* pend = nullptr) {
    check_opts(idx, tax, opts);
    if (!tax) UMGAP_FAIL(UMGAP_ERR_INVALID, "null taxonomy");
    const bool packed = hr.codes != nullptr;
    if ((nreads && ((!hr.nt && !packed) || !read_off)) || (ngroups && (!group_off || !taxon_out)))
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
    // Chunked and software-pipelined over kBufs streams: while the kernels of chunk c run, the
    // nucleotides and offsets of the next chunks upload and the results of the previous one
    // download.  Offsets are uploaded as given and rebased on the device.
    // nucleotides per chunk.  One batch at a time (profiles/r01_e2e_chunk_sweep.log, r02_async_chunk_sweep.log): 16 M for
    // the byte form, 32 M for the packed form, which moves a quarter of the bytes (6.03 -> 5.87 ms per 1 M pairs); the
    // first chunks are shorter so that the kernels start early.  Asynchronous batches overlap each other, so their chunks
    // are as large as the workspaces should get -- equal parts of at most 64 M (5.58 ms; 16 M: 5.94 ms; device-resident:
    // 5.41 ms).  UMGAP_CHUNK_MB, or UMGAP_CHUNK_NT in nucleotides, override it -- read per call, so that tests can move
    // the chunk seams.
    bool chunk_env = false;
    const uint64_t kChunkNt = [&]() -> uint64_t {
        const char* n = getenv("UMGAP_CHUNK_NT");
        chunk_env = true;
        if (n && strtoull(n, nullptr, 10)) return std::max<uint64_t>(512, strtoull(n, nullptr, 10));
        const char* e = getenv("UMGAP_CHUNK_MB");
        const uint64_t mb = e ? strtoull(e, nullptr, 10) : 0;
        if (mb) return mb << 20;
        chunk_env = false;
        if (pend) {
            const uint64_t total = nreads ? read_off[nreads] - read_off[0] : 0, cap = 64ull << 20;
            return std::max<uint64_t>(1ull << 20, ceil_div(total, std::max<uint64_t>(1, ceil_div(total, cap))) + 4096);
        }
        return (hr.codes ? 32ull : 16ull) << 20;
    }();
    const bool ramp = !pend || chunk_env;  // short first chunks
    static const int kBufs = [] {  // chunks in flight (streams); UMGAP_CHUNK_STREAMS overrides
        const char* e = getenv("UMGAP_CHUNK_STREAMS");
        const int v = e ? atoi(e) : 0;
        return v > 0 ? std::min(v, kMaxBufs) : 4;
    }();
    static const bool ordered = getenv("UMGAP_CHUNK_ORDERED") != nullptr;
    if (!idx->chunk_stream[0])
        for (int i = 0; i < kMaxBufs; ++i) {
            UMGAP_CUDA(cudaStreamCreateWithFlags(&idx->chunk_stream[i], cudaStreamNonBlocking));
            UMGAP_CUDA(cudaEventCreateWithFlags(&idx->chunk_done[i], cudaEventDisableTiming));
        }
    cudaStream_t* st = idx->chunk_stream;
    cudaEvent_t* done = idx->chunk_done;
    DevError* err;
    if (pend) {  // a slot of the ring: zero when handed out (cleared once here, and again by the wait that found it set)
        DevError* ring = (DevError*)idx->ws.get(WS_ERRQ, kErrRing * sizeof(DevError));
        if (!idx->errq_ready) {
            UMGAP_CUDA(cudaMemset(ring, 0, kErrRing * sizeof(DevError)));
            UMGAP_CUDA(cudaHostAlloc(&idx->errq_host, kErrRing * sizeof(DevError), cudaHostAllocDefault));
            memset(idx->errq_host, 0, kErrRing * sizeof(DevError));
            UMGAP_CUDA(cudaStreamCreateWithFlags(&idx->errq_stream, cudaStreamNonBlocking));
            idx->errq_ready = true;
        }
        const uint32_t slot = idx->errq_next++ % kErrRing;
        err = ring + slot;
        pend->err = err;
        pend->err_host = (DevError*)idx->errq_host + slot;
    } else {
        err = (DevError*)idx->ws.get(WS_ERR, sizeof(DevError));
        UMGAP_CUDA(cudaMemset(err, 0, sizeof(DevError)));
    }
    // nucleotides before group g (monotone in g): chunk ends are found by bisection
    auto nt_before = [&](uint64_t g) { return read_off[group_off[g]]; };
    g_end = std::min(g_end, ngroups);
    uint64_t g0 = g_begin;
    // asynchronous batches continue the rotation over the chunk streams where the previous batch left it
    int buf = pend ? (int)(idx->chunk_next % (uint32_t)kBufs) : 0, prev = -1;
    try {
        int chunk_no = 0;
        while (g0 < g_end) {
            const uint64_t nt0 = nt_before(g0);
            // the first chunks are short so that the kernels start early: 1/8, 1/4, 1/2 of a chunk, then full ones
            const uint64_t limit = ramp && chunk_no < 3 ? kChunkNt >> (3 - chunk_no) : kChunkNt;
            ++chunk_no;
            uint64_t lo = g0 + 1, hi = g_end;  // largest g1 with nt_before(g1) - nt0 <= limit, at least g0 + 1
            while (lo < hi) {
                const uint64_t mid = lo + (hi - lo + 1) / 2;
                if (nt_before(mid) - nt0 <= limit) lo = mid; else hi = mid - 1;
            }
            const uint64_t g1 = lo;
            const uint64_t r0 = group_off[g0], r1 = group_off[g1];
            // packed chunks start at the 16-nucleotide word that holds their first nucleotide
            const uint64_t shift = packed ? (nt0 & 15u) : 0;
            const uint64_t cnt_nt = read_off[r1] - nt0 + shift, cnt_r = r1 - r0, cnt_g = g1 - g0;
            const uint64_t cap_nt = std::max(cnt_nt, kChunkNt + 16);
            uint64_t* d_roff = (uint64_t*)idx->ws.get(WS_ROFF + buf, (std::max<uint64_t>(cnt_r, kChunkNt / 32) + 1) * 8);
            uint64_t* d_goff = (uint64_t*)idx->ws.get(WS_GOFF + buf, (std::max<uint64_t>(cnt_g, kChunkNt / 32) + 1) * 8);
            uint32_t* d_out = (uint32_t*)idx->ws.get(WS_OUT + buf, std::max<uint64_t>(cnt_g, kChunkNt / 32) * 4 + 16);
            // every stream has its own ids / scratch / hits / codes: the kernels of consecutive chunks are not
            // ordered against each other, so one chunk's classify kernel and the next chunk's lookup kernel fill
            // each other's tails (UMGAP_CHUNK_ORDERED=1 restores the strict order, for measurements)
            uint32_t* ids = (uint32_t*)idx->ws.get(WS_IDS + buf, (2 * cap_nt + 64) * 4);
            uint32_t* scratch = (uint32_t*)idx->ws.get(WS_SCRATCH + buf, (12 * cap_nt + 64) * 4);
            cudaStream_t s = st[buf];
            NtSrc nv;
            uint64_t moved = 0;
            if (!packed) {
                uint8_t* d_nt = (uint8_t*)idx->ws.get(WS_NT + buf, cap_nt + 64);
                UMGAP_CUDA(cudaMemcpyAsync(d_nt, hr.nt + nt0, cnt_nt, cudaMemcpyHostToDevice, s));
                moved = cnt_nt;
                nv.bytes = d_nt;
            } else {
                // device layout of a packed chunk: the code words, then the N-flag words, then the sparse N entries
                const uint64_t w0 = nt0 >> 4, nw = ceil_div(cnt_nt, 16), cap_w = cap_nt / 16 + 2;
                uint8_t* base = (uint8_t*)idx->ws.get(WS_NT + buf, cap_w * 4 + cap_w * 2 + 64 + cap_w * 8);
                uint32_t* d_codes = (uint32_t*)base;
                uint16_t* d_nmask = (uint16_t*)(base + cap_w * 4);
                uint64_t* d_entries = (uint64_t*)(base + ((cap_w * 6 + 63) & ~63ull));
                UMGAP_CUDA(cudaMemcpyAsync(d_codes, hr.codes + w0, nw * 4, cudaMemcpyHostToDevice, s));
                moved = nw * 4;
                UMGAP_CUDA(cudaMemsetAsync(d_nmask, 0, nw * 2, s));
                // the chunk's share of the N entries (ascending word index): found by bisection, uploaded as they are
                const uint64_t* e_lo = std::lower_bound(hr.n_entries, hr.n_entries + hr.n_count, w0 << 16);
                const uint64_t* e_hi = std::lower_bound(e_lo, hr.n_entries + hr.n_count, (w0 + nw) << 16);
                const uint64_t ne = (uint64_t)(e_hi - e_lo);
                if (ne > cap_w) UMGAP_FAIL(UMGAP_ERR_INVALID, "N entries are not strictly ascending by word index");
                if (ne) {
                    UMGAP_CUDA(cudaMemcpyAsync(d_entries, e_lo, ne * 8, cudaMemcpyHostToDevice, s));
                    nmask_scatter_kernel<<<(unsigned)ceil_div(ne, 256), 256, 0, s>>>(d_nmask, d_entries, ne, w0);
                    UMGAP_CUDA(cudaGetLastError());
                    ++g_launch_count;
                    moved += ne * 8;
                }
                nv.codes = d_codes;
                nv.nmask = d_nmask;
            }
            // offsets that form an arithmetic progression (the usual case: reads of one length, pairs) stay on the host
            const uint64_t step_r = uniform_step(read_off + r0, cnt_r), step_g = uniform_step(group_off + g0, cnt_g);
            if (!step_r) UMGAP_CUDA(cudaMemcpyAsync(d_roff, read_off + r0, (cnt_r + 1) * 8, cudaMemcpyHostToDevice, s));
            if (!step_g) UMGAP_CUDA(cudaMemcpyAsync(d_goff, group_off + g0, (cnt_g + 1) * 8, cudaMemcpyHostToDevice, s));
            g_h2d_bytes += moved + (step_r ? 0 : (cnt_r + 1) * 8) + (step_g ? 0 : (cnt_g + 1) * 8);
            g_d2h_bytes += cnt_g * 4;
            rebase_kernel<<<148, 256, 0, s>>>(d_roff, cnt_r + 1, nt0, step_r, shift, d_goff, cnt_g + 1, r0, step_g);
            UMGAP_CUDA(cudaGetLastError());
            ++g_launch_count;
            if (ordered && prev >= 0) UMGAP_CUDA(cudaStreamWaitEvent(s, done[prev], 0));
            uint8_t* hits = (uint8_t*)idx->ws.get(WS_HITS + buf, std::max<uint64_t>(cnt_r, kChunkNt / 16) + 64);
            launch_pipeline(idx, tax, opts, nv, d_roff, cnt_r, cnt_nt, d_goff, cnt_g, ids, scratch, hits, d_out, err, s, buf);
            UMGAP_CUDA(cudaEventRecord(done[buf], s));
            UMGAP_CUDA(cudaMemcpyAsync(taxon_out + g0, d_out, cnt_g * 4, cudaMemcpyDeviceToHost, s));
            prev = buf;
            buf = (buf + 1) % kBufs;
            g0 = g1;
        }
        if (pend) {  // the results are complete when every chunk stream has passed this point
            idx->chunk_next = (uint32_t)buf;
            for (int i = 0; i < kBufs; ++i) {
                UMGAP_CUDA(cudaEventCreateWithFlags(&pend->ev[pend->nev], cudaEventDisableTiming));
                ++pend->nev;
                UMGAP_CUDA(cudaEventRecord(pend->ev[pend->nev - 1], st[i]));
                UMGAP_CUDA(cudaStreamWaitEvent(idx->errq_stream, pend->ev[pend->nev - 1], 0));
            }
            // the error slot goes to its page-locked mirror behind every chunk stream, on a stream of its own (the chunk
            // streams stay independent of each other); the wait then reads host memory (a synchronous copy per batch
            // held the waiting thread for 2.5 ms beside a batch in flight)
            UMGAP_CUDA(cudaMemcpyAsync(pend->err_host, err, sizeof(DevError), cudaMemcpyDeviceToHost, idx->errq_stream));
            UMGAP_CUDA(cudaEventCreateWithFlags(&pend->ev[pend->nev], cudaEventDisableTiming));
            ++pend->nev;
            UMGAP_CUDA(cudaEventRecord(pend->ev[pend->nev - 1], idx->errq_stream));
            return;
        }
        for (int i = 0; i < kBufs; ++i) UMGAP_CUDA(cudaStreamSynchronize(st[i]));
        DevError he;
        UMGAP_CUDA(cudaMemcpy(&he, err, sizeof he, cudaMemcpyDeviceToHost));
        raise_dev_error(he);
    } catch (...) {
        cudaDeviceSynchronize();
        throw;
    }
}
}  // extern "C++"

int umgap_classify_reads(const umgap_index* idx, const umgap_taxonomy* tax,
                         const umgap_pipeline_opts* opts, const uint8_t* nt, const uint64_t* read_off,
                         uint64_t nreads, const uint64_t* group_off, uint64_t ngroups,
                         uint32_t* taxon_out, uint64_t* n_lookups) {
    return guarded([&] {
        HostReads hr;
        hr.nt = nt;
        classify_host(idx, tax, opts, hr, read_off, nreads, group_off, ngroups, taxon_out, n_lookups);
    });
}

int umgap_classify_reads_packed(const umgap_index* idx, const umgap_taxonomy* tax, const umgap_pipeline_opts* opts,
                                const uint32_t* codes, const uint64_t* n_entries, uint64_t n_count, const uint64_t* read_off,
                                uint64_t nreads, const uint64_t* group_off, uint64_t ngroups, uint32_t* taxon_out,
                                uint64_t* n_lookups) {
    return guarded([&] {
        if ((nreads && !codes) || (n_count && !n_entries)) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        HostReads hr;
        hr.codes = codes;
        hr.n_entries = n_entries;
        hr.n_count = n_count;
        classify_host(idx, tax, opts, hr, read_off, nreads, group_off, ngroups, taxon_out, n_lookups);
    });
}

// The asynchronous forms: everything of the batch is enqueued on the handle's chunk streams and the call returns; the
// next batch can be handed over while this one runs, so that its uploads and first kernels fill the tail of this one.
extern "C++" static int classify_async(const umgap_index* idx, const umgap_taxonomy* tax, const umgap_pipeline_opts* opts, const HostReads& hr,
                          const uint64_t* read_off, uint64_t nreads, const uint64_t* group_off, uint64_t ngroups, uint32_t* taxon_out,
                          umgap_pending** out) {
    umgap_pending* p = nullptr;
    const int rc = guarded([&] {
        if (!out) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        *out = nullptr;
        if (!idx) UMGAP_FAIL(UMGAP_ERR_INVALID, "null index");
        if (idx->pending >= kMaxPending) UMGAP_FAIL(UMGAP_ERR_INVALID, "more than %u batches in flight on one index", kMaxPending);
        p = new umgap_pending();
        p->idx = idx;
        classify_host(idx, tax, opts, hr, read_off, nreads, group_off, ngroups, taxon_out, nullptr, 0, ~0ull, p);
        ++idx->pending;
        *out = p;
    });
    if (rc != UMGAP_OK && p) {
        for (int i = 0; i < p->nev; ++i) cudaEventDestroy(p->ev[i]);
        delete p;
    }
    return rc;
}

int umgap_classify_reads_async(const umgap_index* idx, const umgap_taxonomy* tax, const umgap_pipeline_opts* opts, const uint8_t* nt,
                               const uint64_t* read_off, uint64_t nreads, const uint64_t* group_off, uint64_t ngroups,
                               uint32_t* taxon_out, umgap_pending** out) {
    HostReads hr;
    hr.nt = nt;
    return classify_async(idx, tax, opts, hr, read_off, nreads, group_off, ngroups, taxon_out, out);
}

int umgap_classify_reads_packed_async(const umgap_index* idx, const umgap_taxonomy* tax, const umgap_pipeline_opts* opts,
                                      const uint32_t* codes, const uint64_t* n_entries, uint64_t n_count, const uint64_t* read_off,
                                      uint64_t nreads, const uint64_t* group_off, uint64_t ngroups, uint32_t* taxon_out,
                                      umgap_pending** out) {
    if ((nreads && !codes) || (n_count && !n_entries)) return guarded([&] { UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument"); });
    HostReads hr;
    hr.codes = codes;
    hr.n_entries = n_entries;
    hr.n_count = n_count;
    return classify_async(idx, tax, opts, hr, read_off, nreads, group_off, ngroups, taxon_out, out);
}

int umgap_pending_wait(umgap_pending* p) {
    if (!p) return UMGAP_OK;
    const int rc = guarded([&] {
        use_device(p->idx->device);
        for (int i = 0; i < p->nev; ++i) UMGAP_CUDA(cudaEventSynchronize(p->ev[i]));
        DevError he{};
        if (p->err_host) {
            he = *p->err_host;
            if (he.flag) {  // the slot is zero again when the ring comes back to it
                UMGAP_CUDA(cudaMemset(p->err, 0, sizeof(DevError)));
                *p->err_host = DevError{};
            }
        }
        raise_dev_error(he);
    });
    for (int i = 0; i < p->nev; ++i) cudaEventDestroy(p->ev[i]);
    if (p->idx->pending) --p->idx->pending;
    delete p;
    return rc;
}


extern "C++" {
// Replicated index, reads partitioned (SURVEY 8(e) mode 1) inside one process: the groups are cut into one contiguous,
// nucleotide-balanced range per GPU (never inside a uniq group) and one host thread per GPU drives the chunked
// host-buffer path of its replica; the ranges write disjoint parts of taxon_out.  No collective, no device-to-device traffic.
static void classify_host_multi(const umgap_index* const* idx, const umgap_taxonomy* const* tax, int ngpus,
                                const umgap_pipeline_opts* opts, const HostReads& hr, const uint64_t* read_off, uint64_t nreads,
                                const uint64_t* group_off, uint64_t ngroups, uint32_t* taxon_out, uint64_t* n_lookups) {
    if (ngpus < 1 || !idx || !tax) UMGAP_FAIL(UMGAP_ERR_INVALID, "umgap_classify_reads_multi needs at least one index / taxonomy replica");
    for (int i = 0; i < ngpus; ++i) {
        if (!idx[i] || !tax[i]) UMGAP_FAIL(UMGAP_ERR_INVALID, "null replica %d", i);
        if (tax[i]->device != idx[i]->device) UMGAP_FAIL(UMGAP_ERR_INVALID, "replica %d: index and taxonomy live on different devices", i);
        for (int j = 0; j < i; ++j)
            if (idx[j] == idx[i]) UMGAP_FAIL(UMGAP_ERR_INVALID, "replica %d given twice (one in-flight call per handle)", i);
    }
    if (ngpus == 1 || ngroups == 0) {
        classify_host(idx[0], tax[0], opts, hr, read_off, nreads, group_off, ngroups, taxon_out, n_lookups);
        return;
    }
    if ((nreads && !read_off) || !group_off) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
    if (group_off[ngroups] != nreads || group_off[0] != 0) UMGAP_FAIL(UMGAP_ERR_INVALID, "group_off must cover reads 0..nreads");
    std::vector<uint64_t> cut(ngpus + 1, ngroups);
    cut[0] = 0;
    const uint64_t total = read_off[nreads];
    for (int i = 1; i < ngpus; ++i) {  // first group that starts at or after the i-th share of the nucleotides
        const uint64_t want = total / ngpus * i;
        uint64_t lo = cut[i - 1], hi = ngroups;
        while (lo < hi) {
            const uint64_t mid = lo + (hi - lo) / 2;
            if (read_off[group_off[mid]] < want) lo = mid + 1; else hi = mid;
        }
        cut[i] = lo;
    }
    std::vector<int> rc(ngpus, UMGAP_OK);
    std::vector<std::string> msg(ngpus);
    auto run = [&](int i) {
        rc[i] = guarded([&] {
            classify_host(idx[i], tax[i], opts, hr, read_off, nreads, group_off, ngroups, taxon_out, i == 0 ? n_lookups : nullptr,
                          cut[i], cut[i + 1]);
        });
        if (rc[i] != UMGAP_OK) msg[i] = get_error();  // the error text is per thread: hand it to the caller
    };
    std::vector<std::thread> th;
    for (int i = 1; i < ngpus; ++i) th.emplace_back(run, i);
    run(0);
    for (std::thread& t : th) t.join();
    for (int i = 0; i < ngpus; ++i)
        if (rc[i] != UMGAP_OK) {
            set_error("GPU %d: %s", idx[i]->device, msg[i].c_str());
            throw StatusError{rc[i]};
        }
}
}  // extern "C++"

int umgap_classify_reads_multi(const umgap_index* const* idx, const umgap_taxonomy* const* tax, int ngpus,
                               const umgap_pipeline_opts* opts, const uint8_t* nt, const uint64_t* read_off, uint64_t nreads,
                               const uint64_t* group_off, uint64_t ngroups, uint32_t* taxon_out, uint64_t* n_lookups) {
    return guarded([&] {
        HostReads hr;
        hr.nt = nt;
        classify_host_multi(idx, tax, ngpus, opts, hr, read_off, nreads, group_off, ngroups, taxon_out, n_lookups);
    });
}

int umgap_classify_reads_packed_multi(const umgap_index* const* idx, const umgap_taxonomy* const* tax, int ngpus,
                                      const umgap_pipeline_opts* opts, const uint32_t* codes, const uint64_t* n_entries,
                                      uint64_t n_count, const uint64_t* read_off, uint64_t nreads, const uint64_t* group_off,
                                      uint64_t ngroups, uint32_t* taxon_out, uint64_t* n_lookups) {
    return guarded([&] {
        if ((nreads && !codes) || (n_count && !n_entries)) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        HostReads hr;
        hr.codes = codes;
        hr.n_entries = n_entries;
        hr.n_count = n_count;
        classify_host_multi(idx, tax, ngpus, opts, hr, read_off, nreads, group_off, ngroups, taxon_out, n_lookups);
    });
}

}  // extern "C"
