// umgap_index_load_fst: streams an fst Map file once into the device table (replaces
// fst::Map::from_path / from_bytes at prot2kmer2lca.rs:109-114 and prot2tryp2lca.rs:89-94).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>

#include "index.h"

namespace umgap {

int build_var_table_from_fst(const char* path, umgap_index* idx, double load_factor);  // tryptic.cu

namespace {

// Packs k-byte keys into 45-bit codes and hands them to the builder in device batches.
struct KmerSink : FstSink {
    TableBuilder& b;
    umgap_index* idx;
    const size_t k;
    static const size_t kBatch = 1u << 22;
    std::vector<uint64_t> hk;
    std::vector<uint32_t> hv;
    DevBuf<uint64_t> dk;
    DevBuf<uint32_t> dv;

    KmerSink(TableBuilder& builder, umgap_index* i) : b(builder), idx(i), k((size_t)i->k), dk(kBatch), dv(kBatch) {
        hk.reserve(kBatch);
        hv.reserve(kBatch);
    }
    void flush() {
        if (hk.empty()) return;
        UMGAP_CUDA(cudaDeviceSynchronize());  // previous batch's insert has read dk/dv
        UMGAP_CUDA(cudaMemcpy(dk.p, hk.data(), hk.size() * 8, cudaMemcpyHostToDevice));
        UMGAP_CUDA(cudaMemcpy(dv.p, hv.data(), hv.size() * 4, cudaMemcpyHostToDevice));
        b.insert_dev(dk.p, dv.p, hk.size());
        hk.clear();
        hv.clear();
    }
    void on_key(const uint8_t* key, size_t len, uint64_t value) override {
        if (len != k) {
            ++idx->n_skipped;
            return;
        }
        uint64_t code = 0;
        for (size_t i = 0; i < k; ++i) {
            const int c = b.code_for(key[i]);
            if (c < 0) UMGAP_FAIL(UMGAP_ERR_CAPACITY, "index keys use more than 32 distinct byte values");
            code = (code << 5) | (uint64_t)c;
        }
        if (value >= 0xFFFFFFFFull)
            UMGAP_FAIL(UMGAP_ERR_CAPACITY, "index value %llu does not fit 32 bits", (unsigned long long)value);
        hk.push_back(code);
        hv.push_back((uint32_t)value);
        if (hk.size() == kBatch) flush();
    }
};


// ---- the same load on several host threads -----------------------------------------------------------------------
// The file is cut below its first three key bytes into independent subtrees (FstFile::split); worker threads walk
// them and pack the k-byte keys into pinned batches; this thread uploads full batches (cudaMemcpyAsync on a ring of
// streams, each with its own device buffers) and launches the inserts, so the walk, the copies and the insert
// kernels overlap.  The residue alphabet is fixed before the workers start: the bytes met on the first three levels,
// in walk order (a deterministic function of the file, hence the same on every rank of a sharded load).  A key with a
// byte first met deeper than that makes the load start over on the serial path above.
struct alignas(64) PinnedBatch {   // a cache line of its own: the workers fill neighbouring batches
    uint64_t* k = nullptr;
    uint32_t* v = nullptr;
    size_t n = 0;
};
constexpr size_t kPBatch = 1u << 20;

struct BatchQueues {
    std::mutex mu;
    std::condition_variable cv;
    std::vector<PinnedBatch*> free_, full_;
    bool stop = false;
    PinnedBatch* take_free() {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return !free_.empty() || stop; });
        if (free_.empty()) return nullptr;
        PinnedBatch* b = free_.back();
        free_.pop_back();
        return b;
    }
    void give_free(PinnedBatch* b) {
        {
            std::lock_guard<std::mutex> lk(mu);
            b->n = 0;
            free_.push_back(b);
        }
        cv.notify_all();
    }
    void give_full(PinnedBatch* b) {
        {
            std::lock_guard<std::mutex> lk(mu);
            full_.push_back(b);
        }
        cv.notify_all();
    }
};

struct alignas(64) WorkerSink : FstSink {
    uint8_t code_of_byte[256];  // a private copy of the alphabet
    const size_t k;
    BatchQueues& q;
    PinnedBatch* cur = nullptr;
    uint64_t* ck = nullptr;     // the current batch's arrays and fill, kept here while it is being filled
    uint32_t* cv = nullptr;
    size_t cn = 0;
    uint64_t skipped = 0;
    bool unknown_byte = false, bad_value = false;
    WorkerSink(const uint8_t* codes, size_t kk, BatchQueues& queues) : k(kk), q(queues) { memcpy(code_of_byte, codes, 256); }
    void on_key(const uint8_t* key, size_t len, uint64_t value) override {
        if (len != k) {
            ++skipped;
            return;
        }
        uint64_t code = 0;
        for (size_t i = 0; i < k; ++i) {
            const uint8_t c = code_of_byte[key[i]];
            if (c == 0xFF) {
                unknown_byte = true;
                return;
            }
            code = (code << 5) | c;
        }
        if (value >= 0xFFFFFFFFull) {
            bad_value = true;
            return;
        }
        if (!cur) {
            cur = q.take_free();
            if (!cur) return;  // the load was stopped
            ck = cur->k;
            cv = cur->v;
            cn = 0;
        }
        ck[cn] = code;
        cv[cn] = (uint32_t)value;
        if (++cn == kPBatch) {
            cur->n = cn;
            q.give_full(cur);
            cur = nullptr;
        }
    }
    void finish() {
        if (cur) {
            cur->n = cn;
            if (cn) q.give_full(cur);
            else q.give_free(cur);
        }
        cur = nullptr;
    }
};

// Returns false when the file needs the serial path (a residue byte outside the alphabet of the first levels).
bool load_parallel(const char* path, TableBuilder& b, umgap_index* idx, int threads) {
    FstFile file(path);
    const size_t k = (size_t)idx->k;
    std::vector<FstTask> tasks;
    std::vector<uint8_t> seen;
    // keys of up to min(3, k - 1) bytes end inside the split: none of them has k bytes, they only count as skipped
    struct Shallow : FstSink {
        uint64_t n = 0;
        void on_key(const uint8_t*, size_t, uint64_t) override { ++n; }
    } shallow;
    const size_t depth = std::min<size_t>(3, k - 1);
    if (depth == 0) return false;
    file.split(depth, tasks, shallow, seen);
    for (uint8_t byte : seen)
        if (b.code_for(byte) < 0) UMGAP_FAIL(UMGAP_ERR_CAPACITY, "index keys use more than 32 distinct byte values");
    idx->n_skipped += shallow.n;
    if (tasks.empty()) return true;

    const int kRing = 4;
    const int nbatches = threads + kRing + 2;
    std::vector<PinnedBatch> store(nbatches);
    BatchQueues q;
    cudaStream_t st[kRing] = {};
    cudaEvent_t ev[kRing] = {};
    PinnedBatch* inflight[kRing] = {};
    DevBuf<uint64_t> dk[kRing];
    DevBuf<uint32_t> dv[kRing];
    std::vector<std::thread> workers;
    std::vector<WorkerSink*> sinks;
    std::atomic<size_t> next_task{0};
    std::atomic<int> running{threads};
    std::atomic<bool> failed{false};
    std::string worker_error;
    int worker_rc = UMGAP_OK;
    auto cleanup = [&] {
        {
            std::lock_guard<std::mutex> lk(q.mu);
            q.stop = true;
        }
        q.cv.notify_all();
        for (std::thread& t : workers) t.join();
        for (WorkerSink* s : sinks) delete s;
        cudaDeviceSynchronize();
        for (int i = 0; i < kRing; ++i) {
            if (st[i]) cudaStreamDestroy(st[i]);
            if (ev[i]) cudaEventDestroy(ev[i]);
        }
        for (PinnedBatch& pb : store) {
            if (pb.k) cudaFreeHost(pb.k);
            if (pb.v) cudaFreeHost(pb.v);
        }
    };
    try {
        for (PinnedBatch& pb : store) {
            UMGAP_CUDA(cudaHostAlloc((void**)&pb.k, kPBatch * 8, cudaHostAllocDefault));
            UMGAP_CUDA(cudaHostAlloc((void**)&pb.v, kPBatch * 4, cudaHostAllocDefault));
            q.free_.push_back(&pb);
        }
        for (int i = 0; i < kRing; ++i) {
            UMGAP_CUDA(cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking));
            UMGAP_CUDA(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
            dk[i].alloc(kPBatch);
            dv[i].alloc(kPBatch);
        }
        UMGAP_CUDA(cudaDeviceSynchronize());  // the table's memsets (legacy stream) before inserts on the ring's streams
        std::mutex err_mu;
        for (int t = 0; t < threads; ++t) {
            WorkerSink* sink = new WorkerSink(idx->code_of_byte, k, q);
            sinks.push_back(sink);
            workers.emplace_back([&, sink] {
                const int rc = guarded([&] {
                    for (;;) {
                        const size_t i = next_task.fetch_add(1);
                        if (i >= tasks.size() || failed.load() || sink->unknown_byte || sink->bad_value) break;
                        file.stream(tasks[i], *sink);
                    }
                    sink->finish();
                });
                if (rc != UMGAP_OK) {
                    std::lock_guard<std::mutex> lk(err_mu);
                    if (worker_rc == UMGAP_OK) {
                        worker_rc = rc;
                        worker_error = get_error();
                    }
                    failed = true;
                }
                if (sink->unknown_byte || sink->bad_value) failed = true;
                {
                    std::lock_guard<std::mutex> lk(q.mu);
                    --running;
                }
                q.cv.notify_all();
            });
        }
        // uploader
        uint64_t seq = 0;
        for (;;) {
            PinnedBatch* pb = nullptr;
            {
                std::unique_lock<std::mutex> lk(q.mu);
                q.cv.wait(lk, [&] { return !q.full_.empty() || running.load() == 0; });
                if (q.full_.empty()) break;
                pb = q.full_.back();
                q.full_.pop_back();
            }
            const int slot = (int)(seq++ % kRing);
            if (inflight[slot]) {
                UMGAP_CUDA(cudaEventSynchronize(ev[slot]));
                q.give_free(inflight[slot]);
                inflight[slot] = nullptr;
            }
            UMGAP_CUDA(cudaMemcpyAsync(dk[slot].p, pb->k, pb->n * 8, cudaMemcpyHostToDevice, st[slot]));
            UMGAP_CUDA(cudaMemcpyAsync(dv[slot].p, pb->v, pb->n * 4, cudaMemcpyHostToDevice, st[slot]));
            b.insert_dev(dk[slot].p, dv[slot].p, pb->n, st[slot]);
            UMGAP_CUDA(cudaEventRecord(ev[slot], st[slot]));
            inflight[slot] = pb;
        }
        UMGAP_CUDA(cudaDeviceSynchronize());
    } catch (...) {
        failed = true;
        cleanup();
        throw;
    }
    bool unknown = false, bad = false;
    for (WorkerSink* s : sinks) {
        unknown |= s->unknown_byte;
        bad |= s->bad_value;
        idx->n_skipped += s->skipped;
    }
    cleanup();
    if (worker_rc != UMGAP_OK) {
        set_error("%s", worker_error.c_str());
        throw StatusError{worker_rc};
    }
    if (bad) UMGAP_FAIL(UMGAP_ERR_CAPACITY, "an index value does not fit 32 bits");
    return !unknown;
}

}  // namespace
}  // namespace umgap

using namespace umgap;

extern "C" int umgap_index_load_fst(const char* path, int k, int device, double load_factor,
                                    umgap_index** out) {
    return umgap_index_load_fst_shard(path, k, device, load_factor, 0, 1, out);
}

// Every rank streams the whole file (the alphabet codes are assigned in file order, hence identical
// on all ranks) and keeps the keys of its own hash range.
extern "C" int umgap_index_load_fst_shard(const char* path, int k, int device, double load_factor, int shard,
                                          int nshards, umgap_index** out) {
    umgap_index* idx = nullptr;
    int rc = guarded([&] {
        if (!path || !out) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        if (k < 0 || k > 9) UMGAP_FAIL(UMGAP_ERR_INVALID, "k-mer table supports 1 <= k <= 9 (got %d)", k);
        check_shard(shard, nshards, k);
        idx = new umgap_index();
        idx->device = device;
        idx->k = k;
        idx->shard = shard;
        idx->nshards = nshards;
        memset(idx->code_of_byte, 0xFF, sizeof idx->code_of_byte);
        if (k == 0) {
            const int r = build_var_table_from_fst(path, idx, load_factor);
            if (r != UMGAP_OK) throw StatusError{r};
            *out = idx;
            return;
        }
        const uint64_t n = fst_file_len(path);
        // host threads of the walk (UMGAP_LOAD_THREADS; 1 = the serial path)
        const char* te = getenv("UMGAP_LOAD_THREADS");
        int threads = te ? atoi(te) : (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
        if (n < (1u << 16)) threads = std::min(threads, 2);
        bool done = false;
        const bool verbose = getenv("UMGAP_LOAD_VERBOSE") != nullptr;
        auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
        if (threads > 1) {
            TableBuilder b;
            try {
                const double t0 = now();
                b.begin(idx, n, load_factor);
                const double t1 = now();
                done = load_parallel(path, b, idx, threads);
                const double t2 = now();
                if (done) b.finish();
                else b.abort();
                if (verbose)
                    fprintf(stderr, "umgap_index_load_fst: %llu keys, %d threads: table allocation %.3f s, walk + upload + insert %.3f s, "
                            "overflow levels + statistics %.3f s\n", (unsigned long long)n, threads, t1 - t0, t2 - t1, now() - t2);
            } catch (...) {
                b.abort();
                throw;
            }
            if (!done) {  // start over on one thread: free the half-built table, forget the alphabet
                for (int i = 0; i < kMaxLevels; ++i) {
                    if (idx->level_dev[i]) cudaFree(idx->level_dev[i]);
                    idx->level_dev[i] = nullptr;
                    idx->level_nlines[i] = 0;
                }
                idx->nlevels = 0;
                idx->bytes = 0;
                idx->n_skipped = 0;
                idx->alphabet_size = 0;
                memset(idx->code_of_byte, 0xFF, sizeof idx->code_of_byte);
            }
        }
        if (!done) {
            TableBuilder b;
            try {
                b.begin(idx, n, load_factor);
                KmerSink sink(b, idx);
                fst_stream_file(path, sink, nullptr);
                sink.flush();
                b.finish();
            } catch (...) {
                b.abort();
                throw;
            }
        }
        *out = idx;
    });
    if (rc != UMGAP_OK && idx) umgap_index_free(idx);
    return rc;
}
