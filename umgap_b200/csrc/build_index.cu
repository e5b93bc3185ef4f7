// Index construction on the device: `splitkmers | sort | joinkmers | buildindex` from the
// protein / taxon table straight into the GPU-resident k-mer table, without the text streams, the external sort and
// the fst file in between.
//
//   splitkmers  (splitkmers.rs:44-66)    every window of k residues of every protein with the protein's taxon id;
//                                        proteins shorter than k contribute nothing
//   sort                                 by k-mer: here a radix sort of the packed 5-bit-per-residue keys (the order
//                                        differs from the bytewise one, which only `buildindex` needs)
//   joinkmers   (joinkmers.rs:53-105)    per k-mer: every taxon id is replaced by its nearest valid ancestor
//                                        (`validsnapping`; ids the taxonomy does not hold are dropped), the ids are
//                                        counted and aggregated with the hybrid strategy at factor 0.95
//                                        (tree/mix.rs:43-64), the result snapped to a ranked taxon (`ranksnapping`);
//                                        a k-mer none of whose ids survives is not emitted
//   buildindex  (buildindex.rs:32-48)    k-mer -> aggregated taxon: here the insert kernel of table.cu
//
// The sort, the run-length encoding and the prefix sum are CUB library calls (this is the loader, not the hot path);
// the window, aggregation and insert kernels are the library's own.  Ties of the hybrid descent are broken by
// HashSet order in the reference (tree/mix.rs:52-55); here the first maximum in preorder wins.
#include <algorithm>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_run_length_encode.cuh>
#include <cub/device/device_scan.cuh>

#include "index.h"

namespace umgap {

void launch_aggregate(const umgap_taxonomy* tax, int strategy, float factor, float lower_bound, int ranked_only,
                      const uint32_t* taxa_dev, const uint64_t* rec_off_dev, uint64_t nrecs, uint32_t* scratch_dev,
                      uint32_t* out_dev, unsigned int* err_dev, cudaStream_t st);  // stages.cu

constexpr uint64_t kDroppedKey = 1ull << kKeyBits;  // sorts behind every real key (46-bit sort)

// One warp per protein, the lanes stride over its windows.  keys[w] = the packed window, vals[w] = the protein's
// taxon snapped to its nearest valid ancestor; a taxon the taxonomy does not hold drops the window
// (joinkmers.rs:96-98), one beyond the taxonomy's id range is an error (an index panic in the reference).
__global__ void __launch_bounds__(256)
split_kmers_kernel(TaxView tv, const uint8_t* __restrict__ code_of_byte, int k, const uint8_t* __restrict__ aa,
                   const uint64_t* __restrict__ prot_off, const uint64_t* __restrict__ win_off,
                   const uint64_t* __restrict__ prot_taxon, uint64_t nprot, uint64_t* __restrict__ keys,
                   uint32_t* __restrict__ vals, unsigned int* __restrict__ err) {
    __shared__ uint8_t s_code[256];
    s_code[threadIdx.x] = code_of_byte[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t r = warp; r < nprot; r += nwarps) {
        const uint64_t b = prot_off[r], w0 = win_off[r], nw = win_off[r + 1] - w0;
        if (!nw) continue;
        const uint64_t tid = prot_taxon[r];
        uint32_t val = 0;
        if (tid > tv.max_id) {
            if (lane == 0 && atomicCAS(&err[0], 0u, 1u) == 0u) err[1] = (unsigned int)min(tid, (uint64_t)0xFFFFFFFFull);
        } else {
            const uint32_t d = __ldg(tv.dense_of + tid);
            if (d != kNoTaxon) val = __ldg(tv.snap_valid + d);
        }
        for (uint64_t i = lane; i < nw; i += 32) {
            uint64_t key = 0;
            for (int j = 0; j < k; ++j) key = (key << 5) | s_code[aa[b + i + j]];
            keys[w0 + i] = val ? key : kDroppedKey;
            vals[w0 + i] = val;
        }
    }
}

// The same for one pass of a protein table too large to sort at once: only the windows whose hash prefix lies in
// [lo, hi) are kept, appended through a cursor (their order is irrelevant, they are sorted next; the hash spreads the
// k-mers evenly over the passes, and all windows of a k-mer fall into the same pass).
__global__ void __launch_bounds__(256)
split_kmers_range_kernel(TaxView tv, const uint8_t* __restrict__ code_of_byte, int k, const uint8_t* __restrict__ aa,
                         const uint64_t* __restrict__ prot_off, const uint64_t* __restrict__ prot_taxon, uint64_t nprot,
                         uint64_t lo, uint64_t hi, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, uint64_t cap,
                         unsigned long long* __restrict__ cursor, unsigned int* __restrict__ err) {
    __shared__ uint8_t s_code[256];
    s_code[threadIdx.x] = code_of_byte[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t r = warp; r < nprot; r += nwarps) {
        const uint64_t b = prot_off[r], len = prot_off[r + 1] - b;
        if (len < (uint64_t)k) continue;
        const uint64_t nw = len - k + 1;
        const uint64_t tid = prot_taxon[r];
        uint32_t val = 0;
        if (tid > tv.max_id) {
            if (lane == 0 && atomicCAS(&err[0], 0u, 1u) == 0u) err[1] = (unsigned int)min(tid, (uint64_t)0xFFFFFFFFull);
        } else {
            const uint32_t d = __ldg(tv.dense_of + tid);
            if (d != kNoTaxon) val = __ldg(tv.snap_valid + d);
        }
        if (!val) continue;  // taxon the taxonomy does not hold: its windows are dropped (joinkmers.rs:96-98)
        for (uint64_t base = 0; base < nw; base += 32) {
            const uint64_t i = base + lane;
            uint64_t key = 0;
            bool in = false;
            if (i < nw) {
                for (int j = 0; j < k; ++j) key = (key << 5) | s_code[aa[b + i + j]];
                const uint64_t pfx = mix45(key) >> 13;
                in = pfx >= lo && pfx < hi;
            }
            const unsigned m = __ballot_sync(0xffffffffu, in);
            if (!m) continue;
            const int leader = __ffs(m) - 1;
            unsigned long long at = 0;
            if (lane == leader) at = atomicAdd(cursor, (unsigned long long)__popc(m));
            at = __shfl_sync(0xffffffffu, at, leader) + __popc(m & lt_mask);
            if (in) {
                if (at < cap) {
                    keys[at] = key;
                    vals[at] = val;
                } else {
                    err[2] = 1;  // the pass holds more windows than its buffers: the host retries with more passes
                }
            }
        }
    }
}

}  // namespace umgap

using namespace umgap;

#define UMGAP_CUB(call)                                                                                    \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess) UMGAP_FAIL(UMGAP_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_));        \
    } while (0)

extern "C" {

int umgap_index_build_from_proteins(const umgap_taxonomy* tax, const uint8_t* aa, const uint64_t* prot_off,
                                    const uint64_t* prot_taxon, uint64_t nprot, int k, double load_factor,
                                    umgap_index** out) {
    umgap_index* idx = nullptr;
    int rc = guarded([&] {
        if (!tax || !out || (nprot && (!prot_off || !prot_taxon))) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        if (k < 1 || k > 9) UMGAP_FAIL(UMGAP_ERR_INVALID, "k-mer table supports 1 <= k <= 9 (got %d)", k);
        const uint64_t total = nprot ? prot_off[nprot] : 0;
        if (total && !aa) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        use_device(tax->device);
        idx = new umgap_index();
        idx->device = tax->device;
        idx->k = k;
        memset(idx->code_of_byte, 0xFF, sizeof idx->code_of_byte);
        TableBuilder b;
        b.idx = idx;
        // windows per protein; the index alphabet in order of first appearance (at most 32 byte values)
        std::vector<uint64_t> win_off(nprot + 1, 0);
        for (uint64_t r = 0; r < nprot; ++r) {
            const uint64_t len = prot_off[r + 1] - prot_off[r];
            win_off[r + 1] = win_off[r] + (len >= (uint64_t)k ? len - k + 1 : 0);
            if (len >= (uint64_t)k)
                for (uint64_t i = prot_off[r]; i < prot_off[r + 1]; ++i)
                    if (idx->code_of_byte[aa[i]] == 0xFF && b.code_for(aa[i]) < 0)
                        UMGAP_FAIL(UMGAP_ERR_CAPACITY, "index keys use more than 32 distinct byte values");
        }
        const uint64_t W = win_off[nprot];
        try {
            if (!W) {
                b.begin(idx, 0, load_factor);
                b.finish();
                *out = idx;
                return;
            }
            size_t free_b = 0, total_b = 0;
            UMGAP_CUDA(cudaMemGetInfo(&free_b, &total_b));
            // The sort of all windows at once takes about 48 B of device memory per window (and the run-length encoding
            // is a library call with 32-bit counts).  A larger protein table is processed in passes over hash-prefix
            // ranges of the k-mers (UMGAP_BUILD_PASSES forces a number of passes, for tests).
            uint64_t passes = 1;
            if (const char* e = getenv("UMGAP_BUILD_PASSES")) passes = std::max<uint64_t>(1, strtoull(e, nullptr, 10));
            while (passes < 4096 && ((double)W / passes * 1.3 * 64.0 + (double)total > 0.6 * (double)free_b || W / passes * 13 / 10 >= (1ull << 31)))
                ++passes;
            if (passes > 1) {
                DevBuf<uint8_t> d_aa(total), d_code(256);
                DevBuf<uint64_t> d_poff(nprot + 1), d_ptax(nprot);
                DevBuf<unsigned int> d_err(4);
                DevBuf<unsigned long long> d_cursor(1);
                UMGAP_CUDA(cudaMemcpy(d_aa.p, aa, total, cudaMemcpyHostToDevice));
                UMGAP_CUDA(cudaMemcpy(d_code.p, idx->code_of_byte, 256, cudaMemcpyHostToDevice));
                UMGAP_CUDA(cudaMemcpy(d_poff.p, prot_off, (nprot + 1) * 8, cudaMemcpyHostToDevice));
                UMGAP_CUDA(cudaMemcpy(d_ptax.p, prot_taxon, nprot * 8, cudaMemcpyHostToDevice));
                const uint64_t cap = W / passes * 13 / 10 + 65536;
                DevBuf<uint64_t> k_a(cap), k_b(cap), cnt(cap + 1), rec_off(cap + 1), d_nruns(1);
                DevBuf<uint32_t> v_a(cap), v_b(cap), scratch(4 * cap + 8);
                size_t t1 = 0, t2 = 0, t3 = 0;
                UMGAP_CUB(cub::DeviceRadixSort::SortPairs(nullptr, t1, k_a.p, k_b.p, v_a.p, v_b.p, cap, 0, kKeyBits));
                UMGAP_CUB(cub::DeviceRunLengthEncode::Encode(nullptr, t2, k_b.p, k_a.p, cnt.p, d_nruns.p, (int)cap));
                UMGAP_CUB(cub::DeviceScan::ExclusiveSum(nullptr, t3, cnt.p, rec_off.p, cap + 1));
                DevBuf<uint8_t> tmp(std::max(t1, std::max(t2, t3)));
                size_t tmp_bytes = tmp.n;
                std::vector<uint64_t> runs(passes, 0);
                // phase 0 counts the distinct k-mers of every pass (the table is sized before the first insert), phase 1
                // aggregates and inserts
                for (int phase = 0; phase < 2; ++phase) {
                    if (phase == 1) {
                        uint64_t total_keys = 0;
                        for (uint64_t u : runs) total_keys += u;
                        b.begin(idx, total_keys, load_factor);
                    }
                    for (uint64_t ps = 0; ps < passes; ++ps) {
                        const uint64_t lo = (1ull << 32) * ps / passes, hi = (1ull << 32) * (ps + 1) / passes;
                        UMGAP_CUDA(cudaMemset(d_cursor.p, 0, 8));
                        UMGAP_CUDA(cudaMemset(d_err.p, 0, 16));
                        split_kmers_range_kernel<<<(unsigned)std::min<uint64_t>(ceil_div(nprot, 8), 148ull * 16), 256>>>(
                            tax->view, d_code.p, k, d_aa.p, d_poff.p, d_ptax.p, nprot, lo, hi, k_a.p, v_a.p, cap, d_cursor.p, d_err.p);
                        UMGAP_CUDA(cudaGetLastError());
                        unsigned int he[4];
                        unsigned long long n = 0;
                        UMGAP_CUDA(cudaMemcpy(he, d_err.p, 16, cudaMemcpyDeviceToHost));
                        UMGAP_CUDA(cudaMemcpy(&n, d_cursor.p, 8, cudaMemcpyDeviceToHost));
                        if (he[0]) UMGAP_FAIL(UMGAP_ERR_UNKNOWN_TAXON, "taxon id %u lies beyond the taxonomy's id range", he[1]);
                        if (he[2] || n > cap)
                            UMGAP_FAIL(UMGAP_ERR_CAPACITY, "pass %llu of %llu holds %llu windows, more than its buffers (%llu): "
                                       "set UMGAP_BUILD_PASSES higher", (unsigned long long)ps, (unsigned long long)passes, n,
                                       (unsigned long long)cap);
                        if (!n) continue;
                        UMGAP_CUB(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, k_a.p, k_b.p, v_a.p, v_b.p, (uint64_t)n, 0, kKeyBits));
                        UMGAP_CUDA(cudaMemset(cnt.p, 0, (n + 1) * 8));
                        tmp_bytes = tmp.n;
                        UMGAP_CUB(cub::DeviceRunLengthEncode::Encode(tmp.p, tmp_bytes, k_b.p, k_a.p, cnt.p, d_nruns.p, (int)n));
                        uint64_t U = 0;
                        UMGAP_CUDA(cudaMemcpy(&U, d_nruns.p, 8, cudaMemcpyDeviceToHost));
                        if (phase == 0) {
                            runs[ps] = U;
                            tmp_bytes = tmp.n;
                            continue;
                        }
                        tmp_bytes = tmp.n;
                        UMGAP_CUB(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, cnt.p, rec_off.p, U + 1));
                        UMGAP_CUDA(cudaMemset(d_err.p, 0, 16));
                        launch_aggregate(tax, UMGAP_AGG_HYBRID, 0.95f, 0.0f, 1, v_b.p, rec_off.p, U, scratch.p, v_a.p, d_err.p, nullptr);
                        UMGAP_CUDA(cudaDeviceSynchronize());
                        UMGAP_CUDA(cudaMemcpy(he, d_err.p, 8, cudaMemcpyDeviceToHost));
                        if (he[0]) UMGAP_FAIL(UMGAP_ERR_UNKNOWN_TAXON, "Unknown Taxon ID: %u", he[1]);
                        b.insert_dev(k_a.p, v_a.p, U);
                        UMGAP_CUDA(cudaDeviceSynchronize());
                        tmp_bytes = tmp.n;
                    }
                }
                b.finish();
                *out = idx;
                return;
            }
            DevBuf<uint8_t> d_aa(total), d_code(256);
            DevBuf<uint64_t> d_poff(nprot + 1), d_woff(nprot + 1), d_ptax(nprot);
            DevBuf<unsigned int> d_err(2);
            UMGAP_CUDA(cudaMemcpy(d_aa.p, aa, total, cudaMemcpyHostToDevice));
            UMGAP_CUDA(cudaMemcpy(d_code.p, idx->code_of_byte, 256, cudaMemcpyHostToDevice));
            UMGAP_CUDA(cudaMemcpy(d_poff.p, prot_off, (nprot + 1) * 8, cudaMemcpyHostToDevice));
            UMGAP_CUDA(cudaMemcpy(d_woff.p, win_off.data(), (nprot + 1) * 8, cudaMemcpyHostToDevice));
            UMGAP_CUDA(cudaMemcpy(d_ptax.p, prot_taxon, nprot * 8, cudaMemcpyHostToDevice));
            UMGAP_CUDA(cudaMemset(d_err.p, 0, 8));
            DevBuf<uint64_t> k_a(W), k_b(W);
            DevBuf<uint32_t> v_a(W), v_b(W);
            split_kmers_kernel<<<(unsigned)std::min<uint64_t>(ceil_div(nprot, 8), 148ull * 16), 256>>>(
                tax->view, d_code.p, k, d_aa.p, d_poff.p, d_woff.p, d_ptax.p, nprot, k_a.p, v_a.p, d_err.p);
            UMGAP_CUDA(cudaGetLastError());
            unsigned int he[2];
            UMGAP_CUDA(cudaMemcpy(he, d_err.p, 8, cudaMemcpyDeviceToHost));
            if (he[0]) UMGAP_FAIL(UMGAP_ERR_UNKNOWN_TAXON, "taxon id %u lies beyond the taxonomy's id range", he[1]);
            d_aa.free();
            // sort (k-mer, taxon) by k-mer
            size_t tmp_bytes = 0;
            UMGAP_CUB(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k_a.p, k_b.p, v_a.p, v_b.p, W, 0, kKeyBits + 1));
            {
                DevBuf<uint8_t> tmp(tmp_bytes);
                UMGAP_CUB(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, k_a.p, k_b.p, v_a.p, v_b.p, W, 0, kKeyBits + 1));
                UMGAP_CUDA(cudaDeviceSynchronize());
            }
            // runs of equal k-mers: unique keys -> k_a, run lengths -> cnt, then record offsets by a prefix sum
            DevBuf<uint64_t> cnt(W + 1), rec_off(W + 1), d_nruns(1);
            UMGAP_CUDA(cudaMemset(cnt.p, 0, (W + 1) * 8));
            UMGAP_CUB(cub::DeviceRunLengthEncode::Encode(nullptr, tmp_bytes, k_b.p, k_a.p, cnt.p, d_nruns.p, W));
            {
                DevBuf<uint8_t> tmp(tmp_bytes);
                UMGAP_CUB(cub::DeviceRunLengthEncode::Encode(tmp.p, tmp_bytes, k_b.p, k_a.p, cnt.p, d_nruns.p, W));
                UMGAP_CUDA(cudaDeviceSynchronize());
            }
            uint64_t U = 0, last_key = 0;
            UMGAP_CUDA(cudaMemcpy(&U, d_nruns.p, 8, cudaMemcpyDeviceToHost));
            UMGAP_CUDA(cudaMemcpy(&last_key, k_a.p + (U - 1), 8, cudaMemcpyDeviceToHost));
            if (last_key == kDroppedKey) --U;  // the windows of proteins whose taxon the taxonomy does not hold
            UMGAP_CUB(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cnt.p, rec_off.p, U + 1));
            {
                DevBuf<uint8_t> tmp(tmp_bytes);
                UMGAP_CUB(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, cnt.p, rec_off.p, U + 1));
                UMGAP_CUDA(cudaDeviceSynchronize());
            }
            cnt.free();
            k_b.free();
            // joinkmers: hybrid aggregation at 0.95 of every run, ranked snapping (joinkmers.rs:66-75)
            DevBuf<uint32_t> scratch(3 * W + U + 8);
            UMGAP_CUDA(cudaMemset(d_err.p, 0, 8));
            launch_aggregate(tax, UMGAP_AGG_HYBRID, 0.95f, 0.0f, 1, v_b.p, rec_off.p, U, scratch.p, v_a.p, d_err.p, nullptr);
            UMGAP_CUDA(cudaDeviceSynchronize());
            UMGAP_CUDA(cudaMemcpy(he, d_err.p, 8, cudaMemcpyDeviceToHost));
            if (he[0]) UMGAP_FAIL(UMGAP_ERR_UNKNOWN_TAXON, "Unknown Taxon ID: %u", he[1]);
            scratch.free();
            v_b.free();
            rec_off.free();
            // buildindex: the table
            b.begin(idx, U, load_factor);
            b.insert_dev(k_a.p, v_a.p, U);
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

}  // extern "C"
