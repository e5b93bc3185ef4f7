// Host side of the packed read form (umgap_pack_reads): nucleotide bytes -> 2-bit codes + N flags, the input of
// umgap_classify_reads_packed.  A byte is A, C, G or T (dna/mod.rs:34-44: upper case only) iff the letter of its
// code ((x >> 1) ^ (x >> 2)) & 3 equals it; anything else is N.  32 bytes per step where the host has AVX2 (found at
// run time), else eight; a thread per slice.
#include <immintrin.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "common.h"

namespace {

// 16 nucleotides -> (codes, flags)
inline void pack16(const uint8_t* p, uint32_t& codes, uint16_t& flags) {
    uint32_t c = 0, f = 0;
    for (int half = 0; half < 2; ++half) {
        uint64_t x;
        memcpy(&x, p + 8 * half, 8);
        const uint64_t t = ((x >> 1) ^ (x >> 2)) & 0x0303030303030303ull;
        const uint64_t lo = t & 0x0101010101010101ull, hi = (t >> 1) & 0x0101010101010101ull;
        // 'A' + 2 [code 1] + 6 [code 2] + 0x13 [code 3]: no byte overflows
        const uint64_t want = 0x4141414141414141ull + 2 * (lo & ~hi) + 6 * (hi & ~lo) + 0x13 * (lo & hi);
        const uint64_t diff = want ^ x;
        const uint64_t nz = (((diff & 0x7F7F7F7F7F7F7F7Full) + 0x7F7F7F7F7F7F7F7Full) | diff) & 0x8080808080808080ull;
        // gather: 2 bits of every byte -> 16 bits; bit 7 of every byte -> 8 bits
        uint64_t a = (t | (t >> 6)) & 0x000F000F000F000Full;
        a = (a | (a >> 12)) & 0x000000FF000000FFull;
        a = (a | (a >> 24)) & 0xFFFFull;
        const uint64_t b = ((nz >> 7) * 0x0102040810204080ull) >> 56;
        c |= (uint32_t)a << (16 * half);
        f |= (uint32_t)b << (8 * half);
    }
    codes = c;
    flags = (uint16_t)f;
}

// 32 nucleotides -> two code words and their flags: the letter of every byte's code comes from a byte shuffle and is
// compared with the byte; the 2-bit codes of four bytes meet in one byte through two multiply-adds (1, 4 | 1, 16).
__attribute__((target("avx2"))) uint64_t pack_words_avx2(const uint8_t* nt, uint64_t w, uint64_t w_end, uint64_t total_nt,
                                                          uint32_t* codes, std::vector<uint64_t>* entries) {
    const __m256i three = _mm256_set1_epi8(3);
    const __m256i letters = _mm256_setr_epi8('A', 'C', 'G', 'T', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                                             'A', 'C', 'G', 'T', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
    const __m256i w14 = _mm256_set1_epi16(0x0401), w116 = _mm256_set1_epi32(0x00100001);
    const __m256i low_bytes = _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                               0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    for (; w + 2 <= w_end && 16 * w + 32 <= total_nt; w += 2) {
        const __m256i x = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(nt + 16 * w));
        const __m256i t = _mm256_and_si256(_mm256_xor_si256(_mm256_srli_epi16(x, 1), _mm256_srli_epi16(x, 2)), three);
        const uint32_t flags = ~(uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_shuffle_epi8(letters, t), x));
        const __m256i c = _mm256_shuffle_epi8(_mm256_madd_epi16(_mm256_maddubs_epi16(t, w14), w116), low_bytes);
        codes[w] = (uint32_t)_mm256_extract_epi32(c, 0);
        codes[w + 1] = (uint32_t)_mm256_extract_epi32(c, 4);
        if (flags) {
            if (flags & 0xFFFFu) entries->push_back((w << 16) | (flags & 0xFFFFu));
            if (flags >> 16) entries->push_back(((w + 1) << 16) | (flags >> 16));
        }
    }
    return w;
}

void pack_range(const uint8_t* nt, uint64_t total_nt, uint64_t w_begin, uint64_t w_end, uint32_t* codes,
                std::vector<uint64_t>* entries) {
    static const bool avx2 = __builtin_cpu_supports("avx2") && !getenv("UMGAP_PACK_SCALAR");
    if (avx2) w_begin = pack_words_avx2(nt, w_begin, w_end, total_nt, codes, entries);
    for (uint64_t w = w_begin; w < w_end; ++w) {
        const uint64_t x0 = 16 * w;
        uint16_t flags;
        if (x0 + 16 <= total_nt) {
            pack16(nt + x0, codes[w], flags);
        } else {  // the last word: the positions past the end are N
            uint8_t tmp[16];
            memset(tmp, 'N', sizeof tmp);
            memcpy(tmp, nt + x0, total_nt - x0);
            pack16(tmp, codes[w], flags);
        }
        if (flags) entries->push_back((w << 16) | flags);
    }
}

}  // namespace

extern "C" {

uint64_t umgap_packed_words(uint64_t total_nt) { return (total_nt + 15) / 16; }

int umgap_pack_reads(const uint8_t* nt, uint64_t total_nt, uint32_t* codes, uint64_t* n_entries, uint64_t n_cap,
                     uint64_t* n_count, int threads) {
    return umgap::guarded([&] {
        if ((total_nt && (!nt || !codes)) || !n_count || (n_cap && !n_entries)) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        const uint64_t nw = umgap_packed_words(total_nt);
        int t = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
        t = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)std::max(1, t), nw / (1u << 16) + 1));
        std::vector<std::vector<uint64_t>> found((size_t)t);
        if (t == 1) {
            pack_range(nt, total_nt, 0, nw, codes, &found[0]);
        } else {
            std::vector<std::thread> pool;
            for (int i = 0; i < t; ++i)
                pool.emplace_back(pack_range, nt, total_nt, nw * i / t, nw * (i + 1) / t, codes, &found[(size_t)i]);
            for (std::thread& th : pool) th.join();
        }
        uint64_t n = 0;
        for (const auto& f : found) n += f.size();
        *n_count = n;
        if (n > n_cap) UMGAP_FAIL(UMGAP_ERR_CAPACITY, "%llu words hold an N, room for %llu", (unsigned long long)n, (unsigned long long)n_cap);
        uint64_t at = 0;
        for (const auto& f : found) {
            if (!f.empty()) memcpy(n_entries + at, f.data(), f.size() * sizeof(uint64_t));
            at += f.size();
        }
    });
}

}  // extern "C"
