// Construction of the device k-mer table (see table.cuh for the layout) and the random-sector
// gather microbenchmark that measures the roofline denominator on the table's own memory.
#include <algorithm>

#include "index.h"
#include "taxdev.cuh"

namespace umgap {

int build_var_table_from_pairs(umgap_index* idx, const uint8_t* keys, const uint64_t* key_off,
                               const uint64_t* values, uint64_t n, double load_factor);  // tryptic.cu
void free_var_table(void* p);

void check_shard(int shard, int nshards, int k) {
    if (nshards < 1 || nshards > kMaxShards || shard < 0 || shard >= nshards)
        UMGAP_FAIL(UMGAP_ERR_INVALID, "shard %d of %d is out of range (at most %d shards)", shard, nshards, kMaxShards);
    if (nshards > 1 && k <= 0) UMGAP_FAIL(UMGAP_ERR_INVALID, "only k-mer tables can be sharded");
}

enum { C_OVF = 0, C_DUP = 1, C_DISPLACED = 2, C_MAXPROBE = 3, C_INSERTED = 4, C_BADVAL = 5, C_N = 8 };

// Empty table: every meta = kEmptyMeta, every value = kNoValue.
__global__ void fill_empty_kernel(uint4* p, uint64_t n16) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride)
        p[i] = (i & 1) ? make_uint4(kNoValue, kNoValue, kNoValue, kNoValue)
                       : make_uint4(kEmptyMeta, kEmptyMeta, kEmptyMeta, kEmptyMeta);
}

// Folds `val` into a slot's value word.  The word starts as kNoValue; whoever comes first stores
// its value, every later writer of the same key merges (LCA) or is counted as a duplicate.
template <bool LCA>
__device__ __forceinline__ void merge_value(unsigned int* addr, uint32_t val, const TaxView& tv,
                                            unsigned long long* counters) {
    unsigned int old = *reinterpret_cast<volatile unsigned int*>(addr);
    for (;;) {
        if (old == kNoValue) {
            const unsigned int prev = atomicCAS(addr, kNoValue, val);
            if (prev == kNoValue) return;
            old = prev;
        }
        if (!LCA) {
            atomicAdd(&counters[C_DUP], 1ull);
            return;
        }
        const uint32_t merged = lca_ids(tv, old, val);
        if (merged == old) return;
        const unsigned int prev = atomicCAS(addr, old, merged);
        if (prev == old) return;
        old = prev;
    }
}

// One thread per key.  Metas are claimed with 32-bit CAS in slot order, so two threads carrying
// the same key always meet in the same slot (duplicates are detected, never stored twice).
template <bool LCA>
__global__ void insert_kernel(uint32_t* __restrict__ sectors, uint32_t nlines, uint32_t shard, uint32_t nshards,
                              const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                              uint64_t n, int level, uint64_t* __restrict__ ovf_keys,
                              uint32_t* __restrict__ ovf_vals, uint64_t ovf_cap,
                              unsigned long long* __restrict__ counters, TaxView tv) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t key = keys[i];
        const uint32_t val = vals[i];
        if (key == kInvalidKey) continue;
        if (val == kNoValue) {
            atomicAdd(&counters[C_BADVAL], 1ull);
            continue;
        }
        const uint64_t h = mix45(key);
        uint32_t local32;
        if (shard_split(h, nshards, local32) != shard) continue;  // another shard owns this key
        const uint32_t tag = (uint32_t)h & kTagMask;
        bool done = false;
        for (uint32_t d = 0; d < (uint32_t)kMaxDisp && !done; ++d) {
            unsigned int* meta = sectors + probe_sector_local(local32, h, nlines, d) * 8;
            const uint32_t want = (d << 28) | tag;
            for (int j = 0; j < 4 && !done; ++j) {
                unsigned int cur = *reinterpret_cast<volatile unsigned int*>(meta + j);
                if (cur == kEmptyMeta) {
                    const unsigned int prev = atomicCAS(meta + j, kEmptyMeta, want);
                    if (prev == kEmptyMeta) {
                        atomicAdd(&counters[C_INSERTED], 1ull);
                        if (d) atomicAdd(&counters[C_DISPLACED], 1ull);
                        atomicMax(&counters[C_MAXPROBE], (unsigned long long)(level * kMaxDisp + d + 1));
                        merge_value<LCA>(meta + 4 + j, val, tv, counters);
                        done = true;
                        break;
                    }
                    cur = prev;
                }
                if ((cur & ~kFlagBit) == want) {  // the same key is resident
                    merge_value<LCA>(meta + 4 + j, val, tv, counters);
                    done = true;
                }
            }
            if (!done && !(*reinterpret_cast<volatile unsigned int*>(meta) & kFlagBit))
                atomicOr(meta, kFlagBit);
        }
        if (!done) {
            const unsigned long long o = atomicAdd(&counters[C_OVF], 1ull);
            if (o < ovf_cap) {
                ovf_keys[o] = key;
                ovf_vals[o] = val;
            }
        }
    }
}

__global__ void count_flagged_kernel(const uint32_t* __restrict__ sectors, uint64_t nsectors,
                                     unsigned long long* out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long local = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nsectors; i += stride)
        local += sectors[8 * i] >> 31;
    for (int o = 16; o; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(out, local);
}

static void alloc_level(umgap_index* idx, int lv, uint64_t nlines) {
    if (nlines >= (1ull << 32)) UMGAP_FAIL(UMGAP_ERR_CAPACITY, "table level of %llu lines exceeds 2^32", (unsigned long long)nlines);
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, nlines * 128);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        UMGAP_FAIL(UMGAP_ERR_NOMEM, "cannot allocate %.2f GB for table level %d: %s",
                   nlines * 128 / 1e9, lv, cudaGetErrorString(e));
    }
    idx->level_dev[lv] = (uint32_t*)p;
    idx->level_nlines[lv] = (uint32_t)nlines;
    idx->nlevels = lv + 1;
    idx->bytes += nlines * 128;
    fill_empty_kernel<<<148 * 8, 256>>>((uint4*)p, nlines * 8);
    UMGAP_CUDA(cudaGetLastError());
    UMGAP_CUDA(cudaDeviceSynchronize());
}

static uint64_t lines_for(uint64_t keys, double load) {
    const uint64_t nl = (uint64_t)((double)keys / (16.0 * load)) + 1;
    return std::max<uint64_t>(nl, kMinLines);
}

void TableBuilder::begin(umgap_index* i, uint64_t expected_keys, double load_factor) {
    idx = i;
    if (idx->nshards > 1)  // a shard holds its hash range's share of the keys (+3 % for imbalance)
        expected_keys = expected_keys / idx->nshards + expected_keys / (32 * idx->nshards) + 1024;
    expected = expected_keys;
    if (load_factor > 1.0) UMGAP_FAIL(UMGAP_ERR_INVALID, "load factor must be in (0,1]");
    use_device(idx->device);
    if (load_factor <= 0) {
        // Default policy: spend spare HBM on a sparser table.  Measured on the 1e9-key bench index
        // (profiles/README.md): 23 % of the sectors are flagged at load 0.7 and 8 % at 0.5, and the
        // lookup kernel runs 11.5 ms vs 8.6 ms per 1 M read pairs.  Take the sparsest of 0.5 / 0.6 /
        // 0.7 / 0.8 whose table stays below half of the free device memory, else 0.85.
        size_t free_b = 0, total_b = 0;
        UMGAP_CUDA(cudaMemGetInfo(&free_b, &total_b));
        load_factor = 0.85;
        for (double lf : {0.5, 0.6, 0.7, 0.8})
            if ((double)lines_for(expected_keys, lf) * 128.0 <= 0.5 * (double)free_b) {
                load_factor = lf;
                break;
            }
    }
    idx->load_factor = load_factor;
    alloc_level(idx, 0, lines_for(expected_keys, load_factor));
    ovf_cap = std::max<uint64_t>(1u << 20, expected_keys / 8);
    UMGAP_CUDA(cudaMalloc((void**)&ovf_keys, ovf_cap * sizeof(uint64_t)));
    UMGAP_CUDA(cudaMalloc((void**)&ovf_vals, ovf_cap * sizeof(uint32_t)));
    UMGAP_CUDA(cudaMalloc((void**)&counters, C_N * sizeof(unsigned long long)));
    UMGAP_CUDA(cudaMemset(counters, 0, C_N * sizeof(unsigned long long)));
}

static void launch_insert(umgap_index* idx, int lv, const uint64_t* keys, const uint32_t* vals,
                          uint64_t n, uint64_t* ok, uint32_t* ov, uint64_t cap,
                          unsigned long long* counters, const TaxView* tv, cudaStream_t st) {
    if (!n) return;
    const int threads = 256;
    const unsigned blocks = (unsigned)std::min<uint64_t>(ceil_div(n, threads), 148ull * 64);
    if (tv)
        insert_kernel<true><<<blocks, threads, 0, st>>>(idx->level_dev[lv], idx->level_nlines[lv], (uint32_t)idx->shard,
                                                        (uint32_t)idx->nshards, keys, vals, n, lv, ok, ov, cap, counters, *tv);
    else
        insert_kernel<false><<<blocks, threads, 0, st>>>(idx->level_dev[lv], idx->level_nlines[lv], (uint32_t)idx->shard,
                                                         (uint32_t)idx->nshards, keys, vals, n, lv, ok, ov, cap, counters,
                                                         TaxView{});
    UMGAP_CUDA(cudaGetLastError());
}

void TableBuilder::insert_dev(const uint64_t* keys_dev, const uint32_t* vals_dev, uint64_t n,
                              cudaStream_t stream) {
    launch_insert(idx, 0, keys_dev, vals_dev, n, ovf_keys, ovf_vals, ovf_cap, counters, lca_view,
                  stream);
}

void TableBuilder::finish() {
    unsigned long long h[C_N];
    UMGAP_CUDA(cudaDeviceSynchronize());
    UMGAP_CUDA(cudaMemcpy(h, counters, sizeof h, cudaMemcpyDeviceToHost));
    int lv = 0;
    while (h[C_OVF] > 0) {
        if (h[C_OVF] > ovf_cap)
            UMGAP_FAIL(UMGAP_ERR_CAPACITY,
                       "%llu keys overflowed table level %d (buffer holds %llu): lower the load factor",
                       h[C_OVF], lv, (unsigned long long)ovf_cap);
        if (lv + 1 >= kMaxLevels)
            UMGAP_FAIL(UMGAP_ERR_CAPACITY, "table needs more than %d overflow levels", kMaxLevels);
        // move the overflow list aside, reset the counter, build the next level from it
        const uint64_t m = h[C_OVF];
        DevBuf<uint64_t> k2(m);
        DevBuf<uint32_t> v2(m);
        UMGAP_CUDA(cudaMemcpy(k2.p, ovf_keys, m * sizeof(uint64_t), cudaMemcpyDeviceToDevice));
        UMGAP_CUDA(cudaMemcpy(v2.p, ovf_vals, m * sizeof(uint32_t), cudaMemcpyDeviceToDevice));
        UMGAP_CUDA(cudaMemset(counters + C_OVF, 0, sizeof(unsigned long long)));
        ++lv;
        alloc_level(idx, lv, lines_for(m, 0.25));
        launch_insert(idx, lv, k2.p, v2.p, m, ovf_keys, ovf_vals, ovf_cap, counters, lca_view, nullptr);
        UMGAP_CUDA(cudaDeviceSynchronize());
        UMGAP_CUDA(cudaMemcpy(h, counters, sizeof h, cudaMemcpyDeviceToHost));
    }
    if (h[C_BADVAL])
        UMGAP_FAIL(UMGAP_ERR_CAPACITY, "%llu values do not fit the table (must be < 2^32-1)", h[C_BADVAL]);
    if (h[C_DUP] && !lca_view)
        UMGAP_FAIL(UMGAP_ERR_INVALID, "%llu duplicate keys in index input", h[C_DUP]);
    idx->n_keys = h[C_INSERTED];
    idx->n_displaced = h[C_DISPLACED];
    idx->max_probe = h[C_MAXPROBE];
    // flagged buckets of level 0 (what a lookup's second probe depends on)
    UMGAP_CUDA(cudaMemset(counters, 0, sizeof(unsigned long long)));
    count_flagged_kernel<<<148 * 8, 256>>>(idx->level_dev[0], (uint64_t)idx->level_nlines[0] * 4, counters);
    UMGAP_CUDA(cudaGetLastError());
    UMGAP_CUDA(cudaMemcpy(h, counters, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    idx->n_flagged = h[0];
    abort();
}

void TableBuilder::abort() {
    if (ovf_keys) cudaFree(ovf_keys);
    if (ovf_vals) cudaFree(ovf_vals);
    if (counters) cudaFree(counters);
    ovf_keys = nullptr;
    ovf_vals = nullptr;
    counters = nullptr;
}

int TableBuilder::code_for(uint8_t byte) {
    uint8_t& c = idx->code_of_byte[byte];
    if (c == 0xFF) {
        if (idx->alphabet_size >= 32) return -1;
        c = (uint8_t)idx->alphabet_size++;
    }
    return c;
}

// ---- random-sector gather: the measured denominator of the random-sector roofline ------------
__global__ void randsector_kernel(const ulonglong4* __restrict__ table, uint32_t nlines, uint64_t n,
                                  uint64_t seed, unsigned long long* sink) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t acc = 0;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) {  // 4 independent sectors in flight per thread
        ulonglong4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint64_t h = mix45((i + u * stride + seed) & kKeyMask);
            v[u] = load_sector(table + probe_sector(h, nlines, 0));
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    for (; i < n; i += stride) {
        const uint64_t h = mix45((i + seed) & kKeyMask);
        const ulonglong4 v = load_sector(table + probe_sector(h, nlines, 0));
        acc += v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x123456789abcdefull) atomicAdd(sink, 1ull);  // keeps the loads alive
}

}  // namespace umgap

using namespace umgap;

extern "C" {

int umgap_index_from_pairs(const uint8_t* keys, const uint64_t* key_off, const uint64_t* values,
                           uint64_t n, int k, int device, double load_factor, umgap_index** out) {
    return umgap_index_from_pairs_shard(keys, key_off, values, n, k, device, load_factor, 0, 1, out);
}

int umgap_index_from_pairs_shard(const uint8_t* keys, const uint64_t* key_off, const uint64_t* values,
                                 uint64_t n, int k, int device, double load_factor, int shard, int nshards,
                                 umgap_index** out) {
    umgap_index* idx = nullptr;
    int rc = guarded([&] {
        if (!out || (n && (!keys || !values))) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        check_shard(shard, nshards, k);
        if (k < 0 || k > 9)
            UMGAP_FAIL(UMGAP_ERR_INVALID, "k-mer table supports 1 <= k <= 9 (got %d)", k);
        idx = new umgap_index();
        idx->device = device;
        idx->k = k;
        idx->shard = shard;
        idx->nshards = nshards;
        memset(idx->code_of_byte, 0xFF, sizeof idx->code_of_byte);
        if (k == 0) {  // variable-length peptide table (prot2tryp2lca)
            if (n && !key_off) UMGAP_FAIL(UMGAP_ERR_INVALID, "key_off is required for a variable-length table");
            const int r = build_var_table_from_pairs(idx, keys, key_off, values, n, load_factor);
            if (r != UMGAP_OK) throw StatusError{r};
            *out = idx;
            return;
        }
        TableBuilder b;
        try {
            b.begin(idx, n, load_factor);
            const uint64_t batch = 1ull << 24;
            std::vector<uint64_t> hk;
            std::vector<uint32_t> hv;
            DevBuf<uint64_t> dk(std::min(n, batch));
            DevBuf<uint32_t> dv(std::min(n, batch));
            for (uint64_t base = 0; base < n; base += batch) {
                const uint64_t m = std::min(batch, n - base);
                hk.assign(m, kInvalidKey);
                hv.assign(m, 0);
                for (uint64_t i = 0; i < m; ++i) {
                    const uint64_t g = base + i;
                    const uint64_t o = key_off ? key_off[g] : g * (uint64_t)k;
                    const uint64_t len = key_off ? key_off[g + 1] - o : (uint64_t)k;
                    if (len != (uint64_t)k) {
                        ++idx->n_skipped;
                        continue;
                    }
                    uint64_t code = 0;
                    for (int r = 0; r < k; ++r) {
                        const int c = b.code_for(keys[o + r]);
                        if (c < 0)
                            UMGAP_FAIL(UMGAP_ERR_CAPACITY,
                                       "index keys use more than 32 distinct byte values");
                        code = (code << 5) | (uint64_t)c;
                    }
                    if (values[g] >= 0xFFFFFFFFull)
                        UMGAP_FAIL(UMGAP_ERR_CAPACITY, "value %llu of key %llu does not fit 32 bits",
                                   (unsigned long long)values[g], (unsigned long long)g);
                    hk[i] = code;
                    hv[i] = (uint32_t)values[g];
                }
                UMGAP_CUDA(cudaMemcpy(dk.p, hk.data(), m * sizeof(uint64_t), cudaMemcpyHostToDevice));
                UMGAP_CUDA(cudaMemcpy(dv.p, hv.data(), m * sizeof(uint32_t), cudaMemcpyHostToDevice));
                b.insert_dev(dk.p, dv.p, m);
                UMGAP_CUDA(cudaDeviceSynchronize());
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

int umgap_index_shard_desc(const umgap_index* idx, umgap_shard_desc* desc) {
    return guarded([&] {
        if (!idx || !desc) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        if (idx->k <= 0) UMGAP_FAIL(UMGAP_ERR_INVALID, "only k-mer tables can be sharded");
        use_device(idx->device);
        memset(desc, 0, sizeof *desc);
        desc->shard = idx->shard;
        desc->nshards = idx->nshards;
        desc->device = idx->device;
        desc->nlevels = idx->nlevels;
        desc->alphabet_size = idx->alphabet_size;
        memcpy(desc->code_of_byte, idx->code_of_byte, 256);
        for (int lv = 0; lv < idx->nlevels; ++lv) {
            desc->nlines[lv] = idx->level_nlines[lv];
            cudaIpcMemHandle_t h;
            UMGAP_CUDA(cudaIpcGetMemHandle(&h, idx->level_dev[lv]));
            static_assert(sizeof h == 64, "IPC handle size");
            memcpy(desc->ipc[lv], &h, 64);
        }
    });
}

int umgap_index_attach_shards(umgap_index* idx, const umgap_shard_desc* descs, int nshards) {
    return guarded([&] {
        if (!idx || !descs) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        if (nshards != idx->nshards || nshards < 1 || nshards > kMaxShards)
            UMGAP_FAIL(UMGAP_ERR_INVALID, "expected %d shard descriptors", idx->nshards);
        if (idx->attached) UMGAP_FAIL(UMGAP_ERR_INVALID, "shards are already attached");
        use_device(idx->device);
        ShardedView& v = idx->sharded;
        memset(&v, 0, sizeof v);
        v.nshards = nshards;
        v.k = idx->k;
        for (int o = 0; o < nshards; ++o) {
            const umgap_shard_desc& d = descs[o];
            if (d.shard != o || d.nshards != nshards) UMGAP_FAIL(UMGAP_ERR_INVALID, "descriptor %d is not shard %d of %d", o, o, nshards);
            if (memcmp(d.code_of_byte, idx->code_of_byte, 256) != 0)
                UMGAP_FAIL(UMGAP_ERR_INVALID, "shard %d was built with a different residue alphabet order", o);
            v.nlevels[o] = d.nlevels;
            for (int lv = 0; lv < d.nlevels; ++lv) {
                v.nlines[o][lv] = d.nlines[lv];
                if (o == idx->shard) {
                    v.level[o][lv] = reinterpret_cast<const ulonglong4*>(idx->level_dev[lv]);
                } else {  // the peer's HBM, mapped into this process over NVLink
                    cudaIpcMemHandle_t h;
                    memcpy(&h, d.ipc[lv], 64);
                    void* p = nullptr;
                    UMGAP_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
                    idx->ipc_opened.push_back(p);
                    v.level[o][lv] = reinterpret_cast<const ulonglong4*>(p);
                }
            }
        }
        idx->attached = true;
    });
}

int umgap_index_attach_shards_local(umgap_index* idx, umgap_index* const* shards, int nshards) {
    return guarded([&] {
        if (!idx || !shards) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        if (nshards != idx->nshards || nshards < 1 || nshards > kMaxShards)
            UMGAP_FAIL(UMGAP_ERR_INVALID, "expected %d shards", idx->nshards);
        if (idx->attached) UMGAP_FAIL(UMGAP_ERR_INVALID, "shards are already attached");
        use_device(idx->device);
        ShardedView& v = idx->sharded;
        memset(&v, 0, sizeof v);
        v.nshards = nshards;
        v.k = idx->k;
        for (int o = 0; o < nshards; ++o) {
            const umgap_index* s = shards[o];
            if (!s || s->shard != o || s->nshards != nshards || s->k != idx->k)
                UMGAP_FAIL(UMGAP_ERR_INVALID, "handle %d is not shard %d of %d", o, o, nshards);
            if (memcmp(s->code_of_byte, idx->code_of_byte, 256) != 0)
                UMGAP_FAIL(UMGAP_ERR_INVALID, "shard %d was built with a different residue alphabet order", o);
            if (s->device != idx->device) {  // same process, another GPU: plain peer access over NVLink
                const cudaError_t e = cudaDeviceEnablePeerAccess(s->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) UMGAP_CUDA(e);
                (void)cudaGetLastError();
            }
            v.nlevels[o] = s->nlevels;
            for (int lv = 0; lv < s->nlevels; ++lv) {
                v.nlines[o][lv] = s->level_nlines[lv];
                v.level[o][lv] = reinterpret_cast<const ulonglong4*>(s->level_dev[lv]);
            }
        }
        idx->attached = true;
    });
}

void umgap_index_free(umgap_index* idx) {
    if (!idx) return;
    cudaSetDevice(idx->device);
    for (void* p : idx->ipc_opened) cudaIpcCloseMemHandle(p);
    for (int i = 0; i < idx->nlevels; ++i)
        if (idx->level_dev[i]) cudaFree(idx->level_dev[i]);
    if (idx->var_table) free_var_table(idx->var_table);
    idx->ws.release();
    for (int i = 0; i < 2; ++i) {
        if (idx->aux_stream[i]) cudaStreamDestroy(idx->aux_stream[i]);
        if (idx->aux_join[i]) cudaEventDestroy(idx->aux_join[i]);
    }
    if (idx->aux_fork) cudaEventDestroy(idx->aux_fork);
    for (int i = 0; i < 6; ++i) {
        if (idx->chunk_stream[i]) cudaStreamDestroy(idx->chunk_stream[i]);
        if (idx->chunk_done[i]) cudaEventDestroy(idx->chunk_done[i]);
    }
    if (idx->errq_stream) cudaStreamDestroy(idx->errq_stream);
    if (idx->errq_host) cudaFreeHost(idx->errq_host);
    delete idx;
}

// A replica of a fixed-length table on another device (multi-GPU mode 1: index replicated, reads partitioned):
// the levels are copied device to device (NVLink when the GPUs are peers) instead of streaming the file again.
int umgap_index_replicate(const umgap_index* src, int device, umgap_index** out) {
    umgap_index* idx = nullptr;
    int rc = guarded([&] {
        if (!src || !out) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        if (src->k <= 0 || src->var_table) UMGAP_FAIL(UMGAP_ERR_INVALID, "only fixed-length k-mer tables can be replicated");
        if (src->nshards > 1) UMGAP_FAIL(UMGAP_ERR_INVALID, "a shard of a key-range-sharded index cannot be replicated");
        use_device(device);
        idx = new umgap_index();
        idx->device = device;
        idx->k = src->k;
        idx->nlevels = src->nlevels;
        memcpy(idx->code_of_byte, src->code_of_byte, sizeof idx->code_of_byte);
        idx->alphabet_size = src->alphabet_size;
        idx->n_keys = src->n_keys;
        idx->n_skipped = src->n_skipped;
        idx->n_flagged = src->n_flagged;
        idx->n_displaced = src->n_displaced;
        idx->max_probe = src->max_probe;
        idx->bytes = src->bytes;
        idx->load_factor = src->load_factor;
        idx->region_bytes = src->region_bytes;
        for (int i = 0; i < src->nlevels; ++i) {
            const size_t bytes = (size_t)src->level_nlines[i] * 128;
            idx->level_nlines[i] = src->level_nlines[i];
            UMGAP_CUDA(cudaMalloc((void**)&idx->level_dev[i], bytes));
            UMGAP_CUDA(cudaMemcpyPeer(idx->level_dev[i], device, src->level_dev[i], src->device, bytes));
        }
        UMGAP_CUDA(cudaDeviceSynchronize());
        *out = idx;
    });
    if (rc != UMGAP_OK && idx) umgap_index_free(idx);
    return rc;
}

int umgap_index_get_info(const umgap_index* idx, umgap_index_info* info) {
    return guarded([&] {
        if (!idx || !info) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        info->n_keys = idx->n_keys;
        info->n_buckets = 0;
        for (int i = 0; i < idx->nlevels; ++i) info->n_buckets += (uint64_t)idx->level_nlines[i] * 4;
        info->bytes = idx->bytes;
        info->n_skipped = idx->n_skipped;
        info->n_flagged = idx->n_flagged;
        info->n_displaced = idx->n_displaced;
        info->max_probe = idx->max_probe;
        info->k = idx->k;
        info->device = idx->device;
        info->alphabet_size = idx->alphabet_size;
        info->load_factor = idx->load_factor;
    });
}

int umgap_randsector_bench(const umgap_index* idx, uint64_t n_gathers, int iters, double* rate) {
    return guarded([&] {
        if (!idx || !rate || idx->nlevels == 0) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        use_device(idx->device);
        DevBuf<unsigned long long> sink(1);
        UMGAP_CUDA(cudaMemset(sink.p, 0, sizeof(unsigned long long)));
        cudaEvent_t e0, e1;
        UMGAP_CUDA(cudaEventCreate(&e0));
        UMGAP_CUDA(cudaEventCreate(&e1));
        const ulonglong4* t = (const ulonglong4*)idx->level_dev[0];
        const int threads = 256, blocks = 148 * 8;
        for (int w = 0; w < 2; ++w)
            randsector_kernel<<<blocks, threads>>>(t, idx->level_nlines[0], n_gathers, 17 + w, sink.p);
        double best = 0;
        for (int it = 0; it < iters; ++it) {
            UMGAP_CUDA(cudaEventRecord(e0));
            randsector_kernel<<<blocks, threads>>>(t, idx->level_nlines[0], n_gathers,
                                                   1000003ull * (it + 3), sink.p);
            UMGAP_CUDA(cudaEventRecord(e1));
            UMGAP_CUDA(cudaEventSynchronize(e1));
            float ms = 0;
            UMGAP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            best = std::max(best, (double)n_gathers / (ms * 1e-3));
        }
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        *rate = best;
    });
}

}  // extern "C"
