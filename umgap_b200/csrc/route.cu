// Key-range-sharded lookups with an exchange step (mode 2): every rank packs the
// k-mers of its reads into 45-bit hashes, buckets them by owning shard, the ranks swap the buckets
// (NCCL all-to-all over NVLink, issued by the host through torch.distributed), each rank looks up
// what it received in its own shard at local-HBM speed, the answers travel back and are scattered
// into the ids array the classify kernel reads.
//
//   route_pack_kernel      translate + 9-mer packing (the lookup kernel's front half), appends
//                          hash -> send_h[owner], ids index -> send_pos[owner]; k-mers that cannot
//                          be keys (stop codon, N) are answered MISS on the spot
//   lookup_hashes_kernel   one probe chain per received hash against the local shard
//   route_scatter_kernel   answers back into ids
#include <algorithm>

#include "index.h"

namespace umgap {

struct CodonLut72 {
    uint8_t v[72];
};
void make_code_lut_public(const umgap_index* idx, int table, int methionine, uint8_t* out65);  // pipeline.cu

__device__ __forceinline__ uint32_t nt_code_r(uint8_t c) {
    return c == 'T' ? 0u : c == 'C' ? 1u : c == 'A' ? 2u : c == 'G' ? 3u : 4u;
}

constexpr int kRouteWarps = 8;
constexpr int kRouteTile = 128;

template <int K>
__global__ void __launch_bounds__(kRouteWarps * 32)
route_pack_kernel(CodonLut72 lut, uint32_t nshards, const uint8_t* __restrict__ nt,
                  const uint64_t* __restrict__ read_off, uint64_t nreads, uint64_t cap,
                  const BucketPtrs hp, uint32_t* __restrict__ send_pos,
                  unsigned long long* __restrict__ cursors /* [nshards] + [nshards]: overflow flag */,
                  uint32_t* __restrict__ ids, const uint32_t* __restrict__ list, const uint32_t* __restrict__ list_count,
                  const uint64_t* __restrict__ group_off, uint64_t g_lo) {
    // list set: only the reads list[0 .. *list_count) (those the sampled pack kernel left over: longer than one of its
    // batches), their ids in the frame-major layout the classify kernel reads behind the sampled path
    constexpr int W = kRouteTile + 3 * (K - 1);
    __shared__ uint8_t s_lut[72];
    __shared__ uint8_t s_nt[kRouteWarps][W + 4];
    __shared__ uint8_t s_f[kRouteWarps][W + 4];
    __shared__ uint8_t s_r[kRouteWarps][W + 4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1;
    if (threadIdx.x < 72) s_lut[threadIdx.x] = lut.v[threadIdx.x];
    __syncthreads();
    const uint64_t nwarps = (uint64_t)gridDim.x * kRouteWarps;
    if (list) {
        nreads = *list_count;
        if (group_off) list += group_off[g_lo];  // the list of a group range starts at the range's first read
    }
    for (uint64_t ri = (uint64_t)blockIdx.x * kRouteWarps + warp; ri < nreads; ri += nwarps) {
        const uint64_t r = list ? list[ri] : ri;
        const uint64_t off = read_off[r];
        const uint32_t n = (uint32_t)(read_off[r + 1] - off);
        if (n < 3u * K) continue;
        const uint32_t npos = n - 3u * K + 1;
        uint32_t* out = ids + 2 * off;
        for (uint32_t w0 = 0; w0 < npos; w0 += kRouteTile) {
            for (int i = lane; i < W + 2; i += 32) {
                const uint32_t x = w0 + i;
                s_nt[warp][i] = x < n ? (uint8_t)nt_code_r(nt[off + x]) : (uint8_t)4;
            }
            __syncwarp();
            for (int i = lane; i < W; i += 32) {
                const uint32_t a = s_nt[warp][i], b = s_nt[warp][i + 1], c = s_nt[warp][i + 2];
                const bool has_n = ((a | b | c) & 4u) != 0;
                s_f[warp][i] = s_lut[has_n ? 64 : 16 * a + 4 * b + c];
                s_r[warp][i] = s_lut[has_n ? 64 : 16 * (c ^ 2) + 4 * (b ^ 2) + (a ^ 2)];
            }
            __syncwarp();
#pragma unroll 1
            for (int strand = 0; strand < 2; ++strand) {
                const uint8_t* codes = strand ? s_r[warp] : s_f[warp];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int pl = lane + 32 * u;
                    uint64_t key = 0;
                    uint32_t bad = 0;
#pragma unroll
                    for (int i = 0; i < K; ++i) {
                        const uint32_t c = codes[pl + 3 * (strand ? K - 1 - i : i)];
                        bad |= c;
                        key = (key << 5) | (c & 31u);
                    }
                    const uint32_t p = w0 + pl;
                    const bool live = p < npos;
                    const bool valid = live && !(bad & 0x80u);
                    const uint32_t y = strand ? npos - 1 - p : p;
                    const uint32_t pos = (strand ? n : 0u) + (list ? frame_major_index(n, K, y % 3, y / 3) : y);
                    if (live && !valid) out[pos] = kNoValue;
                    const uint64_t h = mix45(key);
                    uint32_t local32;
                    const uint32_t owner = shard_split(h, nshards, local32);
                    // lanes bound for the same shard claim consecutive slots of its bucket with one atomic
                    const unsigned active = __ballot_sync(0xffffffffu, valid);
                    if (valid) {
                        const unsigned peers = __match_any_sync(active, owner);
                        const int leader = __ffs(peers) - 1;
                        unsigned long long base = 0;
                        if (lane == leader) base = atomicAdd(&cursors[owner], (unsigned long long)__popc(peers));
                        base = __shfl_sync(peers, base, leader);
                        const unsigned long long at = base + __popc(peers & lt_mask);
                        if (at < cap) {
                            hp.p[owner][at] = h;
                            send_pos[(uint64_t)owner * cap + at] = (uint32_t)(2 * off + pos);
                        } else {
                            cursors[nshards + owner] = 1;  // bucket overflow: the host retries with a larger capacity
                            out[pos] = kNoValue;
                        }
                    }
                }
            }
            __syncwarp();
        }
    }
}

// h[src * cap + i], i < counts[src]: hashes received from rank src.  Four probe chains in flight
// per thread.
__global__ void __launch_bounds__(256)
lookup_hashes_kernel(const __grid_constant__ TableView t, const uint64_t* __restrict__ h, const uint64_t* __restrict__ counts,
                     uint32_t nsrc, uint64_t cap, uint32_t* __restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint32_t src = 0; src < nsrc; ++src) {
        const uint64_t n = counts[src] < cap ? counts[src] : cap;
        const uint64_t* hs = h + (uint64_t)src * cap;
        uint32_t* os = out + (uint64_t)src * cap;
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += 4 * stride) {
            uint64_t hv[4];
            ulonglong4 sec[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint64_t j = i + u * stride;
                hv[u] = j < n ? hs[j] : ~0ull;
                if (hv[u] != ~0ull) sec[u] = load_sector(sector_addr(t, hv[u], 0, 0));
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint64_t j = i + u * stride;
                if (hv[u] == ~0ull) continue;
                bool more;
                uint32_t v = probe_sector_data(sec[u], (uint32_t)hv[u] & kTagMask, more);
                if (more) v = probe_continue(t, hv[u], 0, 1);
                os[j] = v;
            }
        }
    }
}

__global__ void route_scatter_kernel(const uint32_t* __restrict__ ans, const uint32_t* __restrict__ send_pos,
                                     const unsigned long long* __restrict__ cursors, uint32_t nshards, uint64_t cap,
                                     uint32_t* __restrict__ ids) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint32_t o = 0; o < nshards; ++o) {
        const uint64_t n = cursors[o] < cap ? cursors[o] : cap;
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
            ids[send_pos[(uint64_t)o * cap + i]] = ans[(uint64_t)o * cap + i];
    }
}

// Answers of the sampled pack phase 1 (send_pos = read * 8 + frame): a non-zero taxon raises the frame's bit in the
// read's frame mask (bytes of frame_hits, OR-ed word-wise).
__global__ void route_scatter_hits_kernel(const uint32_t* __restrict__ ans, const uint32_t* __restrict__ send_pos,
                                          const unsigned long long* __restrict__ cursors, uint32_t nshards, uint64_t cap,
                                          uint32_t* __restrict__ frame_hits_words) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint32_t o = 0; o < nshards; ++o) {
        const uint64_t n = cursors[o] < cap ? cursors[o] : cap;
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
            const uint32_t v = ans[(uint64_t)o * cap + i];
            if (v != 0 && v != kNoValue) {
                const uint32_t pos = send_pos[(uint64_t)o * cap + i], rd = pos >> 3, f = pos & 7u;
                atomicOr(&frame_hits_words[rd >> 2], (1u << f) << (8 * (rd & 3u)));
            }
        }
    }
}

// The plain pack kernel over a device work list (pipeline.cu: umgap_route_pack_sampled_dev, phase 2).
void launch_route_pack_list(const umgap_index* idx, const umgap_pipeline_opts* opts, const uint8_t* nt_dev,
                            const uint64_t* read_off_dev, uint64_t cap, const BucketPtrs& hp, uint32_t* send_pos_dev,
                            uint64_t* cursors_dev, uint32_t* ids_dev, const uint32_t* list, const uint32_t* list_count,
                            const uint64_t* group_off_dev, uint64_t g_lo, cudaStream_t st) {
    CodonLut72 lut{};
    make_code_lut_public(idx, opts->table, opts->methionine, lut.v);
    route_pack_kernel<9><<<148, kRouteWarps * 32, 0, st>>>(lut, (uint32_t)idx->nshards, nt_dev, read_off_dev, 0, cap, hp,
                                                          send_pos_dev, reinterpret_cast<unsigned long long*>(cursors_dev),
                                                          ids_dev, list, list_count, group_off_dev, g_lo);
    UMGAP_CUDA(cudaGetLastError());
}

}  // namespace umgap

using namespace umgap;

extern "C" {

extern "C++" void umgap::route_pack_all(const umgap_index* idx, const umgap_pipeline_opts* opts, const uint8_t* nt_dev,
                         const uint64_t* read_off_dev, uint64_t nreads, uint64_t total_nt, uint64_t cap,
                         const BucketPtrs& hp, uint32_t* send_pos_dev, uint64_t* cursors_dev, uint32_t* ids_dev,
                         cudaStream_t st) {
    {
        if (!idx || !opts || !send_pos_dev || !cursors_dev || !ids_dev) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        if (idx->k != 9) UMGAP_FAIL(UMGAP_ERR_INVALID, "the routed path is built for k = 9");
        if (idx->nshards > kMaxShards) UMGAP_FAIL(UMGAP_ERR_INVALID, "the exchange step supports at most %d shards", kMaxShards);
        if (2 * total_nt >= (1ull << 32)) UMGAP_FAIL(UMGAP_ERR_INVALID, "batch too large for 32-bit ids positions");
        use_device(idx->device);
        UMGAP_CUDA(cudaMemsetAsync(cursors_dev, 0, 2 * (size_t)idx->nshards * sizeof(uint64_t), st));
        if (!nreads) return;
        CodonLut72 lut{};
        make_code_lut_public(idx, opts->table, opts->methionine, lut.v);
        const unsigned blocks = (unsigned)std::min<uint64_t>(ceil_div(nreads, kRouteWarps), 148ull * 32);
        route_pack_kernel<9><<<blocks, kRouteWarps * 32, 0, st>>>(lut, (uint32_t)idx->nshards, nt_dev, read_off_dev, nreads, cap,
                                                                 hp, send_pos_dev,
                                                                 reinterpret_cast<unsigned long long*>(cursors_dev), ids_dev,
                                                                 nullptr, nullptr, nullptr, 0);
        UMGAP_CUDA(cudaGetLastError());
    }
}

int umgap_route_pack_dev(const umgap_index* idx, const umgap_pipeline_opts* opts, const uint8_t* nt_dev,
                         const uint64_t* read_off_dev, uint64_t nreads, uint64_t total_nt, uint64_t cap,
                         uint64_t* send_h_dev, uint32_t* send_pos_dev, uint64_t* cursors_dev, uint32_t* ids_dev,
                         void* stream) {
    return guarded([&] {
        if (!idx || !send_h_dev) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        if (idx->nshards > kMaxShards) UMGAP_FAIL(UMGAP_ERR_INVALID, "the exchange step supports at most %d shards", kMaxShards);
        BucketPtrs hp{};
        for (int o = 0; o < idx->nshards; ++o) hp.p[o] = send_h_dev + (uint64_t)o * cap;
        route_pack_all(idx, opts, nt_dev, read_off_dev, nreads, total_nt, cap, hp, send_pos_dev, cursors_dev, ids_dev, (cudaStream_t)stream);
    });
}

int umgap_lookup_hashes_dev(const umgap_index* idx, const uint64_t* h_dev, const uint64_t* counts_dev, int nsrc,
                            uint64_t cap, uint32_t* out_dev, void* stream) {
    return guarded([&] {
        if (!idx || !h_dev || !counts_dev || !out_dev) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        if (idx->k <= 0) UMGAP_FAIL(UMGAP_ERR_INVALID, "index is not a k-mer table");
        use_device(idx->device);
        // the shard's own view: the hashes were routed here because this shard owns them, and its
        // lines are addressed through the shard-local 32 bits exactly as at build time
        LaunchTimer timer(0, (cudaStream_t)stream);
        lookup_hashes_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(idx->view(), h_dev, counts_dev, (uint32_t)nsrc, cap, out_dev);
        UMGAP_CUDA(cudaGetLastError());
        timer.stop();
    });
}

int umgap_route_scatter_dev(const umgap_index* idx, const uint32_t* ans_dev, const uint32_t* send_pos_dev,
                            const uint64_t* cursors_dev, uint64_t cap, uint32_t* ids_dev, void* stream) {
    return guarded([&] {
        if (!idx || !ans_dev || !send_pos_dev || !cursors_dev || !ids_dev) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        use_device(idx->device);
        route_scatter_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(ans_dev, send_pos_dev,
                                                                       reinterpret_cast<const unsigned long long*>(cursors_dev),
                                                                       (uint32_t)idx->nshards, cap, ids_dev);
        UMGAP_CUDA(cudaGetLastError());
    });
}

int umgap_route_scatter_hits_dev(const umgap_index* idx, const uint32_t* ans_dev, const uint32_t* send_pos_dev,
                                 const uint64_t* cursors_dev, uint64_t cap, uint8_t* frame_hits_dev, void* stream) {
    return guarded([&] {
        if (!idx || !ans_dev || !send_pos_dev || !cursors_dev || !frame_hits_dev) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        if ((uintptr_t)frame_hits_dev & 3u) UMGAP_FAIL(UMGAP_ERR_INVALID, "frame_hits_dev must be 4-byte aligned");
        use_device(idx->device);
        route_scatter_hits_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(ans_dev, send_pos_dev,
                                                                            reinterpret_cast<const unsigned long long*>(cursors_dev),
                                                                            (uint32_t)idx->nshards, cap,
                                                                            reinterpret_cast<uint32_t*>(frame_hits_dev));
        UMGAP_CUDA(cudaGetLastError());
    });
}

}  // extern "C"
