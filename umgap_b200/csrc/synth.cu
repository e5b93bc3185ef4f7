// Synthetic workload of SURVEY 8(d), generated on the device from a counter-based hash so that
// any part of it can be re-derived elsewhere (tests/synth_ref.py mirrors this file in numpy):
//
//   residue(j, p)  protein j, position p : 20-letter alphabet, UniProt-like frequencies
//   taxon(j)       the protein's home taxon (uniform over the preorder-numbered taxonomy)
//   value(j, p)    value of the 9-mer starting at (j, p): home taxon (home_pct %), a random
//                  ancestor of it (ancestor_pct %), otherwise an unrelated taxon; identical 9-mers
//                  from several places merge to the LCA of their values
//   pair i         hit pair: two 50-residue fragments of one protein, reverse-translated with
//                  random synonymous codons of table 1, a random 0..2 nt frame shift, mate 2
//                  reverse-complemented, 1 % substitutions, 0.1 % N; otherwise uniform nucleotides
//
// Benchmark/test aid only -- no counterpart in the reference.
#include <algorithm>

#include "index.h"
#include "taxdev.cuh"

namespace umgap {

__host__ __device__ __forceinline__ uint64_t sm64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__host__ __device__ __forceinline__ uint64_t rnd3(uint64_t seed, uint64_t a, uint64_t b) {
    return sm64(sm64(seed ^ (a * 0xD6E8FEB86659FD93ull)) ^ b);
}

constexpr uint64_t kStreamTaxon = 0x7461786F6Eull, kStreamValue = 0x76616C7565ull,
                   kStreamMut = 0x6D7574ull, kStreamCodon = 0x636F646F6Eull;

// cumulative 16-bit thresholds for "ACDEFGHIKLMNPQRSTVWY" (A 8.25 %, C 1.37 %, D 5.45 %, ...)
__constant__ uint16_t c_cum[20] = {5407,  6305,  9877,  14300, 16830, 21463, 22951, 26832, 30638, 36969,
                                   38548, 41209, 44309, 46884, 50510, 54855, 58361, 62858, 63572, 65535};
// table-1 codons per residue, as TCAG-order codon indices (16*b0+4*b1+b2; T0 C1 A2 G3)
__constant__ uint8_t c_ncodon[20] = {4, 2, 2, 2, 2, 4, 2, 3, 2, 6, 1, 2, 4, 2, 6, 6, 4, 4, 1, 2};
__constant__ uint8_t c_codon[20][6] = {
    {52, 53, 54, 55, 0, 0},      // A GCx
    {12, 13, 0, 0, 0, 0},        // C TGT TGC
    {56, 57, 0, 0, 0, 0},        // D GAT GAC
    {58, 59, 0, 0, 0, 0},        // E GAA GAG
    {0, 1, 0, 0, 0, 0},          // F TTT TTC
    {60, 61, 62, 63, 0, 0},      // G GGx
    {24, 25, 0, 0, 0, 0},        // H CAT CAC
    {32, 33, 34, 0, 0, 0},       // I ATT ATC ATA
    {42, 43, 0, 0, 0, 0},        // K AAA AAG
    {2, 3, 16, 17, 18, 19},      // L TTA TTG CTx
    {35, 0, 0, 0, 0, 0},         // M ATG
    {40, 41, 0, 0, 0, 0},        // N AAT AAC
    {20, 21, 22, 23, 0, 0},      // P CCx
    {26, 27, 0, 0, 0, 0},        // Q CAA CAG
    {28, 29, 30, 31, 46, 47},    // R CGx AGA AGG
    {4, 5, 6, 7, 44, 45},        // S TCx AGT AGC
    {36, 37, 38, 39, 0, 0},      // T ACx
    {48, 49, 50, 51, 0, 0},      // V GTx
    {15, 0, 0, 0, 0, 0},         // W TGG
    {8, 9, 0, 0, 0, 0},          // Y TAT TAC
};

__device__ __forceinline__ uint32_t residue_index(uint64_t seed, uint64_t j, uint64_t p) {
    const uint32_t u = (uint32_t)(rnd3(seed, j, p) & 0xFFFF);
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 19; ++i) r += u > c_cum[i];
    return r;
}

__device__ __forceinline__ uint32_t protein_taxon(uint64_t seed, uint64_t j, uint32_t ntaxa) {
    return (uint32_t)(rnd3(seed ^ kStreamTaxon, j, 0) % ntaxa);
}

// One thread per k-mer window: packed key (through the index alphabet codes) and value.
__global__ void synth_windows_kernel(umgap_synth_spec spec, TaxView tv, const uint8_t* __restrict__ code_of_aa,
                                     uint64_t first, uint64_t count, uint64_t* __restrict__ keys,
                                     uint32_t* __restrict__ vals) {
    const uint32_t wpp = spec.protein_len - 8;  // windows per protein
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        const uint64_t w = first + i;
        const uint64_t j = w / wpp;
        const uint32_t p = (uint32_t)(w - j * wpp);
        uint64_t key = 0;
        for (int r = 0; r < 9; ++r) key = (key << 5) | code_of_aa[residue_index(spec.seed, j, p + r)];
        const uint32_t home = protein_taxon(spec.seed, j, tv.n);
        const uint64_t rv = rnd3(spec.seed ^ kStreamValue, j, p);
        const uint32_t u = (uint32_t)(rv % 100);
        uint32_t dense;
        if (u < spec.home_pct) {
            dense = home;
        } else if (u < spec.home_pct + spec.ancestor_pct) {
            const uint32_t d = (uint32_t)((rv >> 32) % ((uint32_t)tv.depth[home] + 1));
            dense = tv.anc[(uint64_t)home * tv.stride + d];
        } else {
            dense = (uint32_t)((rv >> 32) % tv.n);
        }
        keys[i] = key;
        vals[i] = tv.id_of[dense];
    }
}

// One thread per codon slot of a read: 50 slots per 150-nt read (slot 49 is partly cut off by
// the frame shift and filled with random nucleotides).
__global__ void synth_reads_kernel(umgap_synth_spec spec, uint64_t read_seed, uint64_t first_pair,
                                   uint64_t npairs, uint32_t read_len, uint32_t hit_pct,
                                   uint8_t* __restrict__ nt) {
    const uint32_t ncod = read_len / 3;
    const uint32_t slots = ncod + 1;  // +1 covers the tail when read_len % 3 != 0 or shift > 0
    const uint64_t total = npairs * 2 * slots;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const char acgt[5] = "ACGT";
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const uint64_t read = t / slots;
        const uint32_t slot = (uint32_t)(t - read * slots);
        const uint64_t pair = first_pair + read / 2;
        const uint32_t mate = (uint32_t)(read & 1);
        const uint64_t rp = rnd3(read_seed, pair, 0);
        const bool hit = (rp % 100) < hit_pct && spec.protein_len >= ncod;
        const uint64_t j = (rp >> 8) % spec.n_proteins;
        const uint64_t rm = rnd3(read_seed, pair, 1 + mate);
        const uint32_t o = (uint32_t)(rm % (spec.protein_len - ncod + 1));  // fragment start
        const uint32_t shift = (uint32_t)((rm >> 32) % 3);                  // leading random nt
        uint8_t* out = nt + read * read_len;
        // the synthetic strand is `shift` random nt, then the codons of residues o, o+1, ... as
        // long as whole codons fit, then random nt to the end
        for (int c = 0; c < 3; ++c) {
            const uint32_t x = 3 * slot + c;  // position in the (forward) synthetic strand
            if (x >= read_len) break;
            uint32_t base;                    // 0..3 = A C G T
            const uint64_t rx = rnd3(read_seed ^ kStreamMut, pair * 2 + mate, x);
            if (hit && x >= shift && (x - shift) / 3 < (read_len - shift) / 3) {
                const uint32_t q = (x - shift) / 3, ph = (x - shift) % 3;
                const uint32_t aa = residue_index(spec.seed, j, o + q);
                const uint64_t rc = rnd3(read_seed ^ kStreamCodon, pair * 2 + mate, q);
                const uint32_t codon = c_codon[aa][rc % c_ncodon[aa]];
                const uint32_t tcag = (codon >> (2 * (2 - ph))) & 3;  // T0 C1 A2 G3
                base = tcag == 0 ? 3u : tcag == 1 ? 1u : tcag == 2 ? 0u : 2u;
                if ((rx & 0xFFFF) % 100 == 0) base = (uint32_t)((rx >> 16) & 3);  // 1 % substitution
            } else {
                base = (uint32_t)((rx >> 16) & 3);
            }
            uint8_t ch = (uint8_t)acgt[base];
            const bool is_n = ((rx >> 32) % 1000) == 0;  // 0.1 % N
            if (mate == 1) {  // mate 2 is the reverse complement
                ch = (uint8_t)acgt[3 - base];
                out[read_len - 1 - x] = is_n ? 'N' : ch;
            } else {
                out[x] = is_n ? 'N' : ch;
            }
        }
    }
}

}  // namespace umgap

using namespace umgap;

extern "C" {

int umgap_index_build_synthetic(const umgap_synth_spec* spec, const umgap_taxonomy* tax, int device,
                                double load_factor, umgap_index** out) {
    return umgap_index_build_synthetic_shard(spec, tax, device, load_factor, 0, 1, out);
}

int umgap_index_build_synthetic_shard(const umgap_synth_spec* spec, const umgap_taxonomy* tax, int device,
                                      double load_factor, int shard, int nshards, umgap_index** out) {
    umgap_index* idx = nullptr;
    int rc = guarded([&] {
        if (!spec || !tax || !out) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        check_shard(shard, nshards, 9);
        if (spec->protein_len < 9 || spec->n_proteins == 0)
            UMGAP_FAIL(UMGAP_ERR_INVALID, "synthetic proteome needs protein_len >= 9 and n_proteins >= 1");
        if (spec->home_pct + spec->ancestor_pct > 100) UMGAP_FAIL(UMGAP_ERR_INVALID, "percentages exceed 100");
        if (tax->device != device) UMGAP_FAIL(UMGAP_ERR_INVALID, "taxonomy lives on another device");
        idx = new umgap_index();
        idx->device = device;
        idx->k = 9;
        idx->shard = shard;
        idx->nshards = nshards;
        memset(idx->code_of_byte, 0xFF, sizeof idx->code_of_byte);
        TableBuilder b;
        try {
            const uint64_t nwin = spec->n_proteins * (uint64_t)(spec->protein_len - 8);
            b.begin(idx, nwin, load_factor);
            b.lca_view = &tax->view;
            uint8_t code_of_aa[20];
            const char* aas = "ACDEFGHIKLMNPQRSTVWY";
            for (int i = 0; i < 20; ++i) code_of_aa[i] = (uint8_t)b.code_for((uint8_t)aas[i]);
            DevBuf<uint8_t> d_codes(20);
            UMGAP_CUDA(cudaMemcpy(d_codes.p, code_of_aa, 20, cudaMemcpyHostToDevice));
            const uint64_t batch = 1ull << 26;
            DevBuf<uint64_t> dk(std::min(nwin, batch));
            DevBuf<uint32_t> dv(std::min(nwin, batch));
            for (uint64_t first = 0; first < nwin; first += batch) {
                const uint64_t m = std::min(batch, nwin - first);
                synth_windows_kernel<<<148 * 16, 256>>>(*spec, tax->view, d_codes.p, first, m, dk.p, dv.p);
                UMGAP_CUDA(cudaGetLastError());
                b.insert_dev(dk.p, dv.p, m);
            }
            b.finish();
        } catch (...) {
            b.abort();
            throw;
        }
        *out = idx;
    });
    if (rc != UMGAP_OK && idx) umgap_index_free(idx);
    return rc;
}

int umgap_synth_reads_dev(const umgap_synth_spec* spec, uint64_t read_seed, uint64_t first_pair,
                          uint64_t npairs, uint32_t read_len, uint32_t hit_pct, uint8_t* nt_dev,
                          void* stream) {
    return guarded([&] {
        if (!spec || !nt_dev) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        if (read_len < 3) UMGAP_FAIL(UMGAP_ERR_INVALID, "read_len must be >= 3");
        if (!npairs) return;
        synth_reads_kernel<<<148 * 16, 256, 0, (cudaStream_t)stream>>>(*spec, read_seed, first_pair, npairs,
                                                                      read_len, hit_pct, nt_dev);
        UMGAP_CUDA(cudaGetLastError());
    });
}

}  // extern "C"
