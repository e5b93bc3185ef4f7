// Host-side handles behind the opaque C types of include/umgap_gpu.h.
#pragma once
#include "common.h"
#include "table.cuh"

namespace umgap {

// Device view of the taxonomy.  Taxa are renumbered in PREORDER ("dense" index): the subtree
// of x is the contiguous index range [x, last[x]], so "a is an ancestor-or-self of b" is
// a <= b <= last[a], the induced tree of a sorted taxon list is laid out in DFS order, and the
// LCA of a set is the LCA of its smallest and largest member.  anc[x*stride + d] is the
// ancestor of x at depth d (root = depth 0, x itself at depth[x]); replaces the parent walks of
// tree/mod.rs:29-48 and rmq/rtl.rs:41-50.
struct TaxView {
    const uint32_t* dense_of;    // taxon id -> dense index, kNoTaxon when unknown   [max_id+1]
    const uint32_t* id_of;       // dense -> taxon id                                [n]
    const uint32_t* last;        // dense -> last dense index of its subtree         [n]
    const uint32_t* parent;      // dense -> dense parent (root: itself)             [n]
    const uint8_t* depth;        // dense -> depth                                   [n]
    const uint32_t* anc;         // ancestor matrix                                  [n*stride]
    const uint32_t* snap_valid;  // dense -> taxon id after snapping (taxon.rs:294-301)
    const uint32_t* snap_ranked; // same with ranked_only
    const uint8_t* seed_score;   // dense -> TaxonList::score (taxon.rs:181-191, rank.rs:86-99), 0 = None
    uint32_t n;
    uint32_t max_id;
    uint32_t stride;             // max_depth+1 rounded up to a multiple of 8
};
constexpr uint32_t kNoTaxon = 0xFFFFFFFFu;

}  // namespace umgap

struct umgap_taxonomy {
    int device = 0;
    umgap::TaxView view{};
    std::vector<void*> dev_allocs;
    // host copies (error reporting, synthetic generators, tests)
    std::vector<uint64_t> ids, parents;  // as given
    std::vector<uint8_t> ranks, valids;  // as given (umgap_taxonomy_replicate)
    std::vector<uint32_t> h_dense_of, h_id_of, h_parent;
    std::vector<uint8_t> h_depth;
    uint64_t root = 0, max_id = 0;
    uint32_t max_depth = 0;
};

struct umgap_index {
    int device = 0;
    int k = 9;  // 0: variable-length (tryptic) table
    int nlevels = 0;
    uint32_t* level_dev[umgap::kMaxLevels] = {};  // sector arrays (8 words per sector)
    uint32_t level_nlines[umgap::kMaxLevels] = {};
    uint8_t code_of_byte[256];  // 0xFF = byte not in the index alphabet
    int alphabet_size = 0;
    uint64_t n_keys = 0, n_skipped = 0, n_flagged = 0, n_displaced = 0, max_probe = 0;
    uint64_t bytes = 0;
    double load_factor = 0;  // of level 0, as chosen at build time
    uint64_t region_bytes = 0;  // probe-region size of the lookup kernel (0 = default 60 GiB)
    // key-range sharding (multi-GPU, table larger than one GPU): this handle holds shard `shard` of
    // `nshards`; after umgap_index_attach_shards() `sharded` maps every shard (peers through CUDA IPC)
    int shard = 0, nshards = 1;
    bool attached = false;
    umgap::ShardedView sharded{};
    std::vector<void*> ipc_opened;
    // variable-length table (k == 0), see tryptic.cu
    void* var_table = nullptr;
    mutable umgap::Workspace ws;
    // two internal streams + events of the sliced device path (pipeline.cu), created on first use
    mutable cudaStream_t aux_stream[2] = {};
    mutable cudaEvent_t aux_fork = nullptr, aux_join[2] = {};
    // streams + events of the chunked host-buffer path
    mutable cudaStream_t chunk_stream[6] = {};
    mutable cudaEvent_t chunk_done[6] = {};
    // batches enqueued by the asynchronous host-buffer calls and not yet waited for; their device error slots
    mutable uint32_t pending = 0, errq_next = 0, chunk_next = 0;
    mutable bool errq_ready = false;
    // page-locked mirror of the error slots: a batch's slot is copied into it behind its last kernel, on a stream of
    // its own, so that the wait reads host memory instead of issuing a copy
    mutable void* errq_host = nullptr;
    mutable cudaStream_t errq_stream = nullptr;

    umgap::TableView view() const {
        umgap::TableView v{};
        for (int i = 0; i < nlevels; ++i) {
            v.level[i] = reinterpret_cast<const ulonglong4*>(level_dev[i]);
            v.nlines[i] = level_nlines[i];
        }
        v.nlevels = nlevels;
        v.k = k;
        v.nshards = (uint32_t)nshards;
        return v;
    }
};

namespace umgap {

// Incremental builder used by the FST loader, from_pairs and the synthetic generator.
struct TableBuilder {
    umgap_index* idx = nullptr;
    uint64_t expected = 0;
    uint64_t* ovf_keys = nullptr;  // device
    uint32_t* ovf_vals = nullptr;
    unsigned long long* counters = nullptr;  // device: [0] overflow count, [1] dups, [2] displaced,
                                             // [3] max probe, [4] inserted
    uint64_t ovf_cap = 0;
    const TaxView* lca_view = nullptr;  // non-null: duplicate keys combine by LCA of the values

    void begin(umgap_index* idx, uint64_t expected_keys, double load_factor);
    // keys: packed 45-bit keys on the device; vals on the device
    void insert_dev(const uint64_t* keys_dev, const uint32_t* vals_dev, uint64_t n,
                    cudaStream_t stream = nullptr);
    void finish();  // builds the overflow levels, fills stats; throws on duplicates unless LCA mode
    void abort();
    // byte -> code for the index alphabet, assigning a new code on first sight
    int code_for(uint8_t byte);
};

void check_shard(int shard, int nshards, int k);  // table.cu

// The exchange step of the key-range-sharded mode (route.cu, pipeline.cu, exchange.cu): where the pack kernels put the
// hashes bound for shard o -- a bucket in this rank's memory that the host sends (NCCL), or the owner's inbox, stored
// to directly over NVLink peer mappings.

struct BucketPtrs {
    uint64_t* p[kMaxShards];
};
void route_pack_sampled(const umgap_index* idx, const umgap_pipeline_opts* opts, int phase, const uint8_t* nt_dev,
                        const uint64_t* read_off_dev, uint64_t nreads, uint64_t total_nt, uint64_t cap, const BucketPtrs& hp,
                        uint32_t* send_pos_dev, uint64_t* cursors_dev, uint8_t* frame_hits_dev, uint32_t* ids_dev,
                        const uint64_t* group_off_dev, uint64_t g_lo, uint64_t g_hi, int slot, cudaStream_t st);
void route_pack_all(const umgap_index* idx, const umgap_pipeline_opts* opts, const uint8_t* nt_dev, const uint64_t* read_off_dev,
                    uint64_t nreads, uint64_t total_nt, uint64_t cap, const BucketPtrs& hp, uint32_t* send_pos_dev,
                    uint64_t* cursors_dev, uint32_t* ids_dev, cudaStream_t st);

// Per-launch timing of the hot kernels (pipeline.cu): kind 0 = lookup, 1 = classify.
struct TimedLaunch {
    cudaEvent_t a, b;
    int kind;
    int dev;
};
struct LaunchTimer {
    cudaStream_t st;
    TimedLaunch t{};
    bool on;
    LaunchTimer(int kind, cudaStream_t s, bool enabled = true);
    void stop();
    void cancel();  // drop the bracket (events go back to the pool)
};

// FST v2 stream reader (fst_stream.cpp): calls `sink(key, len, value)` for every key in order.
struct FstSink {
    virtual void on_key(const uint8_t* key, size_t len, uint64_t value) = 0;
    virtual ~FstSink() {}
};
void fst_stream_file(const char* path, FstSink& sink, uint64_t* n_keys_footer);
// The same walk cut into independent subtrees (fst_load.cu walks them on several threads).
struct FstTask {
    uint64_t addr, out;   // node below the prefix, outputs summed along the prefix
    uint8_t prefix[7];
    uint8_t plen;
};
class FstFile {
  public:
    explicit FstFile(const char* path);
    ~FstFile();
    uint64_t len() const;
    void split(size_t depth, std::vector<FstTask>& tasks, FstSink& shallow, std::vector<uint8_t>& bytes_seen) const;
    void stream(const FstTask& t, FstSink& sink) const;
    FstFile(const FstFile&) = delete;
    FstFile& operator=(const FstFile&) = delete;

  private:
    struct Impl;
    Impl* impl_;
};
uint64_t fst_file_len(const char* path);  // number of keys recorded in the footer

}  // namespace umgap
