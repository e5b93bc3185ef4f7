// The exchange step of the key-range-sharded index (mode 2) without the host in the loop and without a
// collective library: the kernels themselves move the data over NVLink peer mappings.
//
// Every rank (one GPU) owns a shard of the table and an *exchange region* in its HBM that all ranks map:
//     inbox_h [G][cap] u64   hashes rank s wants looked up here    (stored by rank s's pack kernel)
//     ansbox  [G][cap] u32   answers of owner o to this rank       (stored by owner o's lookup kernel)
//     inbox_cnt[G] u64, flag_h[G] u32, flag_a[G] u32               (fills and epoch flags, stored by the peers)
// One round of a batch on rank `me`:
//     pack kernel      k-mer hashes -> the OWNER's inbox_h[me] (8-byte peer stores, slots claimed from local cursors)
//     signal / wait    fill counts + epoch flag to every owner; spin until every sender's flag shows this epoch
//     lookup kernel    the local shard answers every sender's inbox -> the SENDER's ansbox[me] (4-byte peer stores)
//     signal / wait    epoch flag to every sender; spin until every owner has answered
//     scatter kernel   local ansbox + local send_pos -> frame masks (phase 1) or ids (phase 2)
// Behind `-o | seedextend -s S` a batch takes two rounds (the sampled positions, then the live frames: pipeline.cu,
// lookup_sampled_kernel MODE 3 / 4), otherwise one with every position; the classify kernel follows on the same stream.
// The host only enqueues: no counts exchange, no read-back, no NCCL.  A rank never runs ahead far enough to overwrite a
// box that is still being read: it starts round r + 1 only after every owner has signalled its answers of round r,
// i.e. after every owner's lookup kernel of round r is complete.
//
// The spin kernels assume that every rank has its own GPU (kernels of two ranks waiting on each other cannot share
// one); they give up after ~10 s and raise a status flag instead of hanging.
#include <algorithm>
#include <memory>
#include <thread>

#include "index.h"

namespace umgap {

void pipeline_reserve(const umgap_index* idx, uint64_t nreads, uint64_t total_nt, int buf);  // pipeline.cu
void pipeline_take_error(const umgap_index* idx);
void classify_ids_buf(const umgap_index* idx, const umgap_taxonomy* tax, const umgap_pipeline_opts* opts, const uint32_t* ids_dev,
                      const uint64_t* read_off_dev, uint64_t total_nt, const uint64_t* group_off_dev, uint64_t ngroups,
                      const uint8_t* frame_hits_dev, bool frame_major, uint32_t* taxon_out_dev, int buf, cudaStream_t st);

struct RegionLayout {
    uint64_t inbox_h, ansbox, inbox_cnt, flag_h, flag_a, bytes;
    RegionLayout(int n, uint64_t cap) {
        inbox_h = 0;
        ansbox = inbox_h + (uint64_t)n * cap * 8;
        inbox_cnt = (ansbox + (uint64_t)n * cap * 4 + 127) & ~127ull;
        flag_h = inbox_cnt + 128;  // kMaxShards * 8 <= 128
        flag_a = flag_h + 128;
        bytes = flag_a + 128;
    }
};

struct AnsPtrs {
    uint32_t* p[kMaxShards];
};
struct SignalPtrs {
    unsigned long long* cnt[kMaxShards];  // peer o: &inbox_cnt[me]
    uint32_t* flag[kMaxShards];           // peer o: &flag[me]
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Thread o: (the fills of the buckets this rank packed for owner o, then) this round's epoch into owner o's flag.
// Runs after the kernel whose peer stores it publishes, on the same stream.
__global__ void exchange_signal_kernel(SignalPtrs sp, const unsigned long long* __restrict__ cursors, uint64_t cap, uint32_t epoch,
                                       int n, int with_counts) {
    const int o = threadIdx.x;
    if (o >= n) return;
    if (with_counts) {
        const unsigned long long c = cursors[o];
        *(volatile unsigned long long*)sp.cnt[o] = c < cap ? c : cap;
    }
    __threadfence_system();
    st_release_sys(sp.flag[o], epoch);
}

// Thread s: spin until rank s has published this round (flags only grow; wrap-safe comparison).
__global__ void exchange_wait_kernel(const uint32_t* __restrict__ flags, uint32_t epoch, int n, uint32_t* status, long long timeout_cycles) {
    const int s = threadIdx.x;
    if (s >= n) return;
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(flags + s) - epoch) < 0) {
        if (clock64() - t0 > timeout_cycles) {
            atomicOr(status, 1u);
            break;
        }
        __nanosleep(200);
    }
    __threadfence_system();
}

// inbox_h[src][i], i < inbox_cnt[src]: the hashes rank src wants looked up in this shard; the answers go to
// out.p[src][i] = rank src's ansbox[me][i].  Four probe chains in flight per thread (route.cu: lookup_hashes_kernel).
__global__ void __launch_bounds__(256)
exchange_lookup_kernel(const __grid_constant__ TableView t, const uint64_t* __restrict__ inbox_h,
                       const unsigned long long* __restrict__ inbox_cnt, uint32_t nsrc, uint64_t cap, const AnsPtrs out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint32_t src = 0; src < nsrc; ++src) {
        const uint64_t n = inbox_cnt[src] < cap ? inbox_cnt[src] : cap;
        const uint64_t* hs = inbox_h + (uint64_t)src * cap;
        uint32_t* os = out.p[src];
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

}  // namespace umgap

using namespace umgap;

struct umgap_exchange {
    const umgap_index* shard = nullptr;
    const umgap_taxonomy* tax = nullptr;
    int rank = 0, n = 1, device = 0;
    int lane = 0;  // workspace set of the shard's handle this context uses (contexts that run side by side on one shard differ in it)
    uint64_t cap = 0, max_total_nt = 0, max_reads = 0;
    uint8_t* region[kMaxShards] = {};  // this rank's mapping of every rank's exchange region
    uint32_t* send_pos = nullptr;
    uint64_t* cursors = nullptr;   // [n] fills + [n] overflow flags of the last round
    uint64_t* totals = nullptr;    // [0] overflow seen, [1] lookups routed (accumulated on the device)
    uint32_t* ids = nullptr;
    uint8_t* frame_hits = nullptr;
    uint32_t* status = nullptr;    // bit 0: a wait timed out
    uint32_t epoch = 0;
    long long timeout_cycles = 20000000000ll;
};

namespace {

__global__ void exchange_tally_kernel(const unsigned long long* __restrict__ cursors, int n, unsigned long long* totals) {
    unsigned long long sum = 0, ovf = 0;
    for (int o = 0; o < n; ++o) {
        sum += cursors[o];
        ovf |= cursors[n + o];
    }
    if (ovf) totals[0] = 1;
    totals[1] += sum;
}

void check_rc(int rc) {
    if (rc != UMGAP_OK) throw StatusError{rc};
}

// One exchange round after the pack kernel of `ex` (see the head of this file).
void exchange_round(umgap_exchange* ex, cudaStream_t st) {
    const RegionLayout L(ex->n, ex->cap);
    const uint32_t epoch = ++ex->epoch;
    SignalPtrs sh{}, sa{};
    AnsPtrs ans{};
    for (int o = 0; o < ex->n; ++o) {
        sh.cnt[o] = reinterpret_cast<unsigned long long*>(ex->region[o] + L.inbox_cnt) + ex->rank;
        sh.flag[o] = reinterpret_cast<uint32_t*>(ex->region[o] + L.flag_h) + ex->rank;
        sa.cnt[o] = nullptr;
        sa.flag[o] = reinterpret_cast<uint32_t*>(ex->region[o] + L.flag_a) + ex->rank;
        ans.p[o] = reinterpret_cast<uint32_t*>(ex->region[o] + L.ansbox) + (uint64_t)ex->rank * ex->cap;
    }
    uint8_t* mine = ex->region[ex->rank];
    unsigned long long* cursors = reinterpret_cast<unsigned long long*>(ex->cursors);
    exchange_tally_kernel<<<1, 1, 0, st>>>(cursors, ex->n, reinterpret_cast<unsigned long long*>(ex->totals));
    exchange_signal_kernel<<<1, 32, 0, st>>>(sh, cursors, ex->cap, epoch, ex->n, 1);
    {
        LaunchTimer timer(3, st);
        exchange_wait_kernel<<<1, 32, 0, st>>>(reinterpret_cast<const uint32_t*>(mine + L.flag_h), epoch, ex->n, ex->status, ex->timeout_cycles);
        timer.stop();
    }
    {
        LaunchTimer timer(0, st);
        // CTAs per SM of the lookup kernel (UMGAP_XLOOKUP_CTAS): four probe chains per thread reach the random-request
        // ceiling with a fraction of the SM's threads, and what it leaves free runs the other lane's pack / scatter /
        // classify kernels beside it
        static const unsigned per_sm = [] {
            const char* e = getenv("UMGAP_XLOOKUP_CTAS");
            const int v = e ? atoi(e) : 0;
            return (unsigned)(v > 0 ? std::min(v, 16) : 8);
        }();
        exchange_lookup_kernel<<<148 * per_sm, 256, 0, st>>>(ex->shard->view(), reinterpret_cast<const uint64_t*>(mine + L.inbox_h),
                                                       reinterpret_cast<const unsigned long long*>(mine + L.inbox_cnt), (uint32_t)ex->n,
                                                       ex->cap, ans);
        timer.stop();
    }
    exchange_signal_kernel<<<1, 32, 0, st>>>(sa, cursors, ex->cap, epoch, ex->n, 0);
    {
        LaunchTimer timer(3, st);
        exchange_wait_kernel<<<1, 32, 0, st>>>(reinterpret_cast<const uint32_t*>(mine + L.flag_a), epoch, ex->n, ex->status, ex->timeout_cycles);
        timer.stop();
    }
    UMGAP_CUDA(cudaGetLastError());
}

void exchange_classify(umgap_exchange* ex, const umgap_pipeline_opts* opts, const uint8_t* nt_dev, const uint64_t* read_off_dev,
                       uint64_t nreads, uint64_t total_nt, const uint64_t* group_off_dev, uint64_t ngroups, uint32_t* out_dev,
                       cudaStream_t st) {
    if (!ex || !opts) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
    if (total_nt > ex->max_total_nt || nreads > ex->max_reads)
        UMGAP_FAIL(UMGAP_ERR_INVALID, "batch larger than the buffers of this exchange context (%llu nt, %llu reads)",
                   (unsigned long long)ex->max_total_nt, (unsigned long long)ex->max_reads);
    if (nreads && (!nt_dev || !read_off_dev)) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
    if (ngroups && (!group_off_dev || !out_dev)) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
    use_device(ex->device);
    const RegionLayout L(ex->n, ex->cap);
    BucketPtrs hp{};
    for (int o = 0; o < ex->n; ++o) hp.p[o] = reinterpret_cast<uint64_t*>(ex->region[o] + L.inbox_h) + (uint64_t)ex->rank * ex->cap;
    const uint32_t* ansbox = reinterpret_cast<const uint32_t*>(ex->region[ex->rank] + L.ansbox);
    // every rank must take the same number of rounds: the choice depends on the options only
    const bool sampled = umgap_route_sampled_applies(ex->shard, opts) != 0;
    if (sampled && ((uintptr_t)nt_dev & 15u)) UMGAP_FAIL(UMGAP_ERR_INVALID, "nt_dev must be 16-byte aligned");
    if (sampled) {
        for (int phase = 1; phase <= 2; ++phase) {
            {
                LaunchTimer timer(2, st);
                route_pack_sampled(ex->shard, opts, phase, nt_dev, read_off_dev, nreads, total_nt, ex->cap, hp, ex->send_pos, ex->cursors,
                                   ex->frame_hits, ex->ids, nullptr, 0, 0, ex->lane, st);
                timer.stop();
            }
            exchange_round(ex, st);
            LaunchTimer timer(4, st);
            if (phase == 1)
                check_rc(umgap_route_scatter_hits_dev(ex->shard, ansbox, ex->send_pos, ex->cursors, ex->cap, ex->frame_hits, st));
            else
                check_rc(umgap_route_scatter_dev(ex->shard, ansbox, ex->send_pos, ex->cursors, ex->cap, ex->ids, st));
            timer.stop();
        }
        if (ngroups)
            classify_ids_buf(ex->shard, ex->tax, opts, ex->ids, read_off_dev, total_nt, group_off_dev, ngroups, ex->frame_hits, true, out_dev,
                             ex->lane, st);
    } else {
        route_pack_all(ex->shard, opts, nt_dev, read_off_dev, nreads, total_nt, ex->cap, hp, ex->send_pos, ex->cursors, ex->ids, st);
        exchange_round(ex, st);
        check_rc(umgap_route_scatter_dev(ex->shard, ansbox, ex->send_pos, ex->cursors, ex->cap, ex->ids, st));
        if (ngroups)
            classify_ids_buf(ex->shard, ex->tax, opts, ex->ids, read_off_dev, total_nt, group_off_dev, ngroups, nullptr, false, out_dev,
                             ex->lane, st);
    }
}

}  // namespace

extern "C" {

uint64_t umgap_exchange_bucket_cap(int nranks, uint64_t max_total_nt) {
    if (nranks < 1) return 0;
    // every one of the 2 * nt lookups of a batch valid and evenly spread over the owners, plus 8 % for skew
    return (uint64_t)(2.0 * (double)max_total_nt / nranks * 1.08) + 4096;
}

uint64_t umgap_exchange_region_bytes(int nranks, uint64_t max_total_nt) {
    if (nranks < 1 || nranks > kMaxShards) return 0;
    return RegionLayout(nranks, umgap_exchange_bucket_cap(nranks, max_total_nt)).bytes;
}

int umgap_exchange_create(const umgap_index* shard, const umgap_taxonomy* tax, int rank, int nranks, uint64_t max_total_nt,
                          void* const* regions, umgap_exchange** out) {
    return umgap_exchange_create_lane(shard, tax, rank, nranks, max_total_nt, regions, 0, out);
}

int umgap_exchange_create_lane(const umgap_index* shard, const umgap_taxonomy* tax, int rank, int nranks, uint64_t max_total_nt,
                               void* const* regions, int lane, umgap_exchange** out) {
    umgap_exchange* ex = nullptr;
    int rc = guarded([&] {
        if (lane < 0 || lane > 5) UMGAP_FAIL(UMGAP_ERR_INVALID, "lane must be in 0..5");
        if (!shard || !tax || !regions || !out) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        if (nranks < 1 || nranks > kMaxShards || rank < 0 || rank >= nranks) UMGAP_FAIL(UMGAP_ERR_INVALID, "rank %d of %d", rank, nranks);
        if (shard->nshards != nranks || shard->shard != rank)
            UMGAP_FAIL(UMGAP_ERR_INVALID, "index handle is shard %d of %d, expected shard %d of %d", shard->shard, shard->nshards, rank, nranks);
        if (shard->k != 9) UMGAP_FAIL(UMGAP_ERR_INVALID, "the exchange step is built for k = 9");
        if (tax->device != shard->device) UMGAP_FAIL(UMGAP_ERR_INVALID, "index and taxonomy live on different devices");
        if (2 * max_total_nt >= (1ull << 32)) UMGAP_FAIL(UMGAP_ERR_INVALID, "batch too large for 32-bit ids positions");
        use_device(shard->device);
        ex = new umgap_exchange();
        ex->shard = shard;
        ex->tax = tax;
        ex->rank = rank;
        ex->n = nranks;
        ex->device = shard->device;
        ex->lane = lane;
        ex->max_total_nt = max_total_nt;
        ex->max_reads = max_total_nt / 27 + 1;
        ex->cap = umgap_exchange_bucket_cap(nranks, max_total_nt);
        for (int o = 0; o < nranks; ++o) {
            if (!regions[o]) UMGAP_FAIL(UMGAP_ERR_INVALID, "null exchange region %d", o);
            ex->region[o] = (uint8_t*)regions[o];
        }
        UMGAP_CUDA(cudaMalloc((void**)&ex->send_pos, (size_t)nranks * ex->cap * 4));
        UMGAP_CUDA(cudaMalloc((void**)&ex->cursors, 2 * kMaxShards * 8));
        UMGAP_CUDA(cudaMalloc((void**)&ex->totals, 16));
        UMGAP_CUDA(cudaMalloc((void**)&ex->ids, (2 * max_total_nt + 64) * 4));
        UMGAP_CUDA(cudaMalloc((void**)&ex->frame_hits, ex->max_reads + 64));
        UMGAP_CUDA(cudaMalloc((void**)&ex->status, 4));
        UMGAP_CUDA(cudaMemset(ex->totals, 0, 16));
        UMGAP_CUDA(cudaMemset(ex->status, 0, 4));
        UMGAP_CUDA(cudaMemset(ex->cursors, 0, 2 * kMaxShards * 8));
        // the workspaces of the pack and classify kernels, so that no launch of a batch allocates (an allocation
        // synchronises the device, and a device may be spinning in a wait kernel for a peer that is enqueued later)
        pipeline_reserve(shard, ex->max_reads, max_total_nt, lane);
        UMGAP_CUDA(cudaDeviceSynchronize());
        *out = ex;
    });
    if (rc != UMGAP_OK && ex) {
        umgap_exchange_free(ex);
    }
    return rc;
}

void umgap_exchange_free(umgap_exchange* ex) {
    if (!ex) return;
    cudaSetDevice(ex->device);
    cudaFree(ex->send_pos);
    cudaFree(ex->cursors);
    cudaFree(ex->totals);
    cudaFree(ex->ids);
    cudaFree(ex->frame_hits);
    cudaFree(ex->status);
    delete ex;
}

int umgap_exchange_classify_dev(umgap_exchange* ex, const umgap_pipeline_opts* opts, const uint8_t* nt_dev,
                                const uint64_t* read_off_dev, uint64_t nreads, uint64_t total_nt, const uint64_t* group_off_dev,
                                uint64_t ngroups, uint32_t* taxon_out_dev, void* stream) {
    return guarded([&] { exchange_classify(ex, opts, nt_dev, read_off_dev, nreads, total_nt, group_off_dev, ngroups, taxon_out_dev, (cudaStream_t)stream); });
}

int umgap_exchange_status(umgap_exchange* ex, uint64_t* lookups_routed) {
    return guarded([&] {
        if (!ex) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        use_device(ex->device);
        uint64_t totals[2];
        uint32_t status = 0;
        UMGAP_CUDA(cudaMemcpy(totals, ex->totals, 16, cudaMemcpyDeviceToHost));
        UMGAP_CUDA(cudaMemcpy(&status, ex->status, 4, cudaMemcpyDeviceToHost));
        UMGAP_CUDA(cudaMemset(ex->totals, 0, 16));
        UMGAP_CUDA(cudaMemset(ex->status, 0, 4));
        if (lookups_routed) *lookups_routed = totals[1];
        if (status & 1u) UMGAP_FAIL(UMGAP_ERR_CUDA, "exchange step: rank %d waited more than the time limit for a peer (results of the batch are invalid)", ex->rank);
        if (totals[0]) UMGAP_FAIL(UMGAP_ERR_CAPACITY, "exchange step: a bucket of rank %d overflowed (key skew beyond the 8 %% slack); results of the batch are invalid", ex->rank);
        pipeline_take_error(ex->shard);
    });
}

}  // extern "C"

// ---- one process driving all the shards (the CLI, a Rust host): peer access between the GPUs, regions from cudaMalloc ----
struct umgap_sharded {
    int n = 0;
    // lanes: complete sets of exchange contexts (regions, buckets, streams) that consecutive batches alternate between, so
    // that one batch's pack kernels run beside the previous batch's lookup kernels (UMGAP_SHARDED_LANES, default 2)
    int nlanes = 1;
    uint64_t calls = 0;
    uint64_t max_total_nt = 0;
    std::vector<umgap_exchange*> ex;     // [lane * n + i]
    std::vector<void*> region;           // [lane * n + i]
    std::vector<int> device;
    std::vector<cudaStream_t> stream;    // [lane * n + i]
    // staging of the host-buffer entry point, per GPU
    std::vector<uint8_t*> d_nt;
    std::vector<uint64_t*> d_roff, d_goff;
    std::vector<uint32_t*> d_out;
    std::vector<uint64_t> cap_reads, cap_groups;
};

extern "C" {

void umgap_sharded_free(umgap_sharded* s) {
    if (!s) return;
    for (umgap_exchange* e : s->ex) umgap_exchange_free(e);
    for (size_t x = 0; x < s->region.size(); ++x) {
        cudaSetDevice(s->device[x % s->n]);
        cudaFree(s->region[x]);
    }
    for (size_t x = 0; x < s->stream.size(); ++x) {
        cudaSetDevice(s->device[x % s->n]);
        if (s->stream[x]) cudaStreamDestroy(s->stream[x]);
    }
    for (int i = 0; i < s->n && i < (int)s->device.size(); ++i) {
        cudaSetDevice(s->device[i]);
        if (i < (int)s->d_nt.size()) {
            cudaFree(s->d_nt[i]);
            cudaFree(s->d_roff[i]);
            cudaFree(s->d_goff[i]);
            cudaFree(s->d_out[i]);
        }
    }
    delete s;
}

int umgap_sharded_create(const umgap_index* const* shards, const umgap_taxonomy* const* tax, int n, uint64_t max_total_nt,
                         umgap_sharded** out) {
    umgap_sharded* s = nullptr;
    int rc = guarded([&] {
        if (!shards || !tax || !out) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        if (n < 1 || n > kMaxShards) UMGAP_FAIL(UMGAP_ERR_INVALID, "1..%d shards", kMaxShards);
        s = new umgap_sharded();
        s->n = n;
        s->max_total_nt = max_total_nt;
        for (int i = 0; i < n; ++i) {
            if (!shards[i] || !tax[i]) UMGAP_FAIL(UMGAP_ERR_INVALID, "null shard %d", i);
            s->device.push_back(shards[i]->device);
            for (int j = 0; j < i; ++j)
                if (s->device[j] == s->device[i]) UMGAP_FAIL(UMGAP_ERR_INVALID, "shards %d and %d share device %d: the exchange step needs a GPU per shard", j, i, s->device[i]);
        }
        const uint64_t bytes = umgap_exchange_region_bytes(n, max_total_nt);
        {
            const char* e = getenv("UMGAP_SHARDED_LANES");
            const int v = e ? atoi(e) : 2;
            s->nlanes = std::max(1, std::min(v, 4));
        }
        for (int i = 0; i < n; ++i) {
            use_device(s->device[i]);
            for (int j = 0; j < n; ++j) {
                if (j == i) continue;
                int can = 0;
                UMGAP_CUDA(cudaDeviceCanAccessPeer(&can, s->device[i], s->device[j]));
                if (!can) UMGAP_FAIL(UMGAP_ERR_CUDA, "GPU %d cannot map GPU %d's memory (no peer access)", s->device[i], s->device[j]);
                cudaError_t e = cudaDeviceEnablePeerAccess(s->device[j], 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) (void)cudaGetLastError();
                else UMGAP_CUDA(e);
            }
        }
        for (int lane = 0; lane < s->nlanes; ++lane)
            for (int i = 0; i < n; ++i) {
                use_device(s->device[i]);
                void* r = nullptr;
                UMGAP_CUDA(cudaMalloc(&r, bytes));
                UMGAP_CUDA(cudaMemset(r, 0, bytes));
                s->region.push_back(r);
                cudaStream_t st;
                UMGAP_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
                s->stream.push_back(st);
            }
        for (int lane = 0; lane < s->nlanes; ++lane)
            for (int i = 0; i < n; ++i) {
                umgap_exchange* ex = nullptr;
                check_rc(umgap_exchange_create_lane(shards[i], tax[i], i, n, max_total_nt, s->region.data() + (size_t)lane * n, lane, &ex));
                s->ex.push_back(ex);
            }
        s->d_nt.assign(n, nullptr);
        s->d_roff.assign(n, nullptr);
        s->d_goff.assign(n, nullptr);
        s->d_out.assign(n, nullptr);
        s->cap_reads.assign(n, 0);
        s->cap_groups.assign(n, 0);
        *out = s;
    });
    if (rc != UMGAP_OK && s) umgap_sharded_free(s);
    return rc;
}

// One device-resident batch per shard, enqueued on the context's own streams; returns when everything is enqueued.
int umgap_classify_reads_sharded_dev(umgap_sharded* s, const umgap_pipeline_opts* opts, const uint8_t* const* nt_dev,
                                     const uint64_t* const* read_off_dev, const uint64_t* nreads, const uint64_t* total_nt,
                                     const uint64_t* const* group_off_dev, const uint64_t* ngroups, uint32_t* const* taxon_out_dev) {
    return guarded([&] {
        if (!s || !opts || !nt_dev || !read_off_dev || !nreads || !total_nt || !group_off_dev || !ngroups || !taxon_out_dev)
            UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        // One host thread per GPU enqueues that GPU's batch.  A launch can block the host while its device is busy (the
        // first launch of a kernel loads it -- CUDA's lazy module loading -- and that waits for the device to drain); a
        // device may at that moment be spinning in a wait kernel for a peer whose work is enqueued by this same call, so
        // the GPUs must not share an enqueuing thread.
        std::vector<int> rc(s->n, UMGAP_OK);
        std::vector<std::string> msg(s->n);
        const size_t lane0 = (size_t)(s->calls++ % s->nlanes) * s->n;
        auto run = [&](int i) {
            rc[i] = guarded([&] {
                exchange_classify(s->ex[lane0 + i], opts, nt_dev[i], read_off_dev[i], nreads[i], total_nt[i], group_off_dev[i], ngroups[i],
                                  taxon_out_dev[i], s->stream[lane0 + i]);
            });
            if (rc[i] != UMGAP_OK) msg[i] = get_error();
        };
        std::vector<std::thread> th;
        for (int i = 1; i < s->n; ++i) th.emplace_back(run, i);
        run(0);
        for (std::thread& t : th) t.join();
        for (int i = 0; i < s->n; ++i)
            if (rc[i] != UMGAP_OK) {
                set_error("shard %d: %s", i, msg[i].c_str());
                throw StatusError{rc[i]};
            }
    });
}

int umgap_sharded_sync(umgap_sharded* s, uint64_t* lookups_routed) {
    return guarded([&] {
        if (!s) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        for (size_t x = 0; x < s->stream.size(); ++x) {
            use_device(s->device[x % s->n]);
            UMGAP_CUDA(cudaStreamSynchronize(s->stream[x]));
        }
        uint64_t sum = 0;
        int first_rc = UMGAP_OK;
        std::string msg;
        for (size_t i = 0; i < s->ex.size(); ++i) {
            uint64_t r = 0;
            const int rc = umgap_exchange_status(s->ex[i], &r);
            sum += r;
            if (rc != UMGAP_OK && first_rc == UMGAP_OK) {
                first_rc = rc;
                msg = get_error();
            }
        }
        if (lookups_routed) *lookups_routed = sum;
        if (first_rc != UMGAP_OK) {
            set_error("%s", msg.c_str());
            throw StatusError{first_rc};
        }
    });
}

// Host buffers: the groups are cut into one nucleotide-balanced range per GPU; batches larger than the context's buffers
// go through in several passes.  Results are those of umgap_classify_reads on an unsharded table.
int umgap_classify_reads_sharded(umgap_sharded* s, const umgap_pipeline_opts* opts, const uint8_t* nt, const uint64_t* read_off,
                                 uint64_t nreads, const uint64_t* group_off, uint64_t ngroups, uint32_t* taxon_out,
                                 uint64_t* n_lookups) {
    return guarded([&] {
        if (!s || !opts) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        if ((nreads && (!nt || !read_off)) || (ngroups && (!group_off || !taxon_out))) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        if (n_lookups) {
            uint64_t c = 0;
            for (uint64_t r = 0; r < nreads; ++r) {
                const uint64_t len = read_off[r + 1] - read_off[r];
                if (len >= 27) c += 2 * (len - 26);
            }
            *n_lookups = c;
        }
        if (!ngroups) return;
        if (group_off[ngroups] != nreads || group_off[0] != 0) UMGAP_FAIL(UMGAP_ERR_INVALID, "group_off must cover reads 0..nreads");
        const int n = s->n;
        auto nt_before = [&](uint64_t g) { return read_off[group_off[g]]; };
        std::vector<uint64_t> roff, goff;
        uint64_t g0 = 0;
        while (g0 < ngroups) {
            // this pass: up to max_total_nt nucleotides per GPU
            std::vector<uint64_t> cut(n + 1, g0);
            for (int i = 0; i < n; ++i) {
                const uint64_t base = nt_before(cut[i]);
                uint64_t lo = cut[i], hi = ngroups;  // largest g with nt_before(g) - base <= max_total_nt
                while (lo < hi) {
                    const uint64_t mid = lo + (hi - lo + 1) / 2;
                    if (nt_before(mid) - base <= s->max_total_nt && group_off[mid] - group_off[cut[i]] <= s->max_total_nt / 27)
                        lo = mid;
                    else
                        hi = mid - 1;
                }
                cut[i + 1] = lo;
            }
            if (cut[n] == g0) UMGAP_FAIL(UMGAP_ERR_INVALID, "a group of reads is larger than the exchange context's buffers");
            // balance the pass over the GPUs when it is the last one (fewer nucleotides than n full shares)
            if (cut[n] == ngroups && (nt_before(ngroups) - nt_before(g0)) * 10 <= s->max_total_nt * 9 * (uint64_t)n &&
                (group_off[ngroups] - group_off[g0]) * 10 <= s->max_total_nt / 27 * 9 * (uint64_t)n) {
                const uint64_t a = nt_before(g0), total = nt_before(ngroups) - a;
                for (int i = 1; i < n; ++i) {
                    const uint64_t want = a + total / n * i;
                    uint64_t lo = cut[i - 1], hi = ngroups;
                    while (lo < hi) {
                        const uint64_t mid = lo + (hi - lo) / 2;
                        if (nt_before(mid) < want) lo = mid + 1; else hi = mid;
                    }
                    cut[i] = lo;
                }
            }
            std::vector<const uint8_t*> p_nt(n);
            std::vector<const uint64_t*> p_roff(n), p_goff(n);
            std::vector<uint32_t*> p_out(n);
            std::vector<uint64_t> c_reads(n), c_nt(n), c_groups(n);
            for (int i = 0; i < n; ++i) {
                use_device(s->device[i]);
                const uint64_t ga = cut[i], gb = cut[i + 1], ra = group_off[ga], rb = group_off[gb], na = read_off[ra], nb = read_off[rb];
                c_reads[i] = rb - ra;
                c_nt[i] = nb - na;
                c_groups[i] = gb - ga;
                if (!s->d_nt[i]) UMGAP_CUDA(cudaMalloc((void**)&s->d_nt[i], s->max_total_nt + 64));
                if (c_reads[i] + 1 > s->cap_reads[i]) {
                    cudaFree(s->d_roff[i]);
                    s->cap_reads[i] = c_reads[i] + 1 + c_reads[i] / 4;
                    UMGAP_CUDA(cudaMalloc((void**)&s->d_roff[i], s->cap_reads[i] * 8));
                }
                if (c_groups[i] + 1 > s->cap_groups[i]) {
                    cudaFree(s->d_goff[i]);
                    cudaFree(s->d_out[i]);
                    s->cap_groups[i] = c_groups[i] + 1 + c_groups[i] / 4;
                    UMGAP_CUDA(cudaMalloc((void**)&s->d_goff[i], s->cap_groups[i] * 8));
                    UMGAP_CUDA(cudaMalloc((void**)&s->d_out[i], s->cap_groups[i] * 4));
                }
                roff.resize(c_reads[i] + 1);
                goff.resize(c_groups[i] + 1);
                for (uint64_t r = 0; r <= c_reads[i]; ++r) roff[r] = read_off[ra + r] - na;
                for (uint64_t g = 0; g <= c_groups[i]; ++g) goff[g] = group_off[ga + g] - ra;
                // blocking copies (pageable memory): done before anything of this pass is enqueued
                UMGAP_CUDA(cudaMemcpy(s->d_nt[i], nt + na, c_nt[i], cudaMemcpyHostToDevice));
                UMGAP_CUDA(cudaMemcpy(s->d_roff[i], roff.data(), (c_reads[i] + 1) * 8, cudaMemcpyHostToDevice));
                UMGAP_CUDA(cudaMemcpy(s->d_goff[i], goff.data(), (c_groups[i] + 1) * 8, cudaMemcpyHostToDevice));
                p_nt[i] = s->d_nt[i];
                p_roff[i] = s->d_roff[i];
                p_goff[i] = s->d_goff[i];
                p_out[i] = s->d_out[i];
            }
            check_rc(umgap_classify_reads_sharded_dev(s, opts, p_nt.data(), p_roff.data(), c_reads.data(), c_nt.data(), p_goff.data(),
                                                      c_groups.data(), p_out.data()));
            check_rc(umgap_sharded_sync(s, nullptr));
            for (int i = 0; i < n; ++i) {
                use_device(s->device[i]);
                if (c_groups[i]) UMGAP_CUDA(cudaMemcpy(taxon_out + cut[i], s->d_out[i], c_groups[i] * 4, cudaMemcpyDeviceToHost));
            }
            g0 = cut[n];
        }
    });
}

}  // extern "C"
