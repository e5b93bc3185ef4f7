// Host side of the packed read form (umgap_pack_reads): nucleotide bytes -> 2-bit codes + N flags, the input of
// umgap_classify_reads_packed.  A byte is A, C, G or T (dna/mod.rs:34-44: upper case only) iff the letter of its
// code ((x >> 1) ^ (x >> 2)) & 3 equals it; anything else is N.  Eight bytes per step, a thread per slice.
#include <algorithm>
#include <cstdint>
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

void pack_range(const uint8_t* nt, uint64_t total_nt, uint64_t w_begin, uint64_t w_end, uint32_t* codes,
                std::vector<uint64_t>* entries) {
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
