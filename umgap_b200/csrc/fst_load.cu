// umgap_index_load_fst: streams an fst Map file once into the device table (replaces
// fst::Map::from_path / from_bytes at prot2kmer2lca.rs:109-114 and prot2tryp2lca.rs:89-94).
#include <algorithm>

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
        *out = idx;
    });
    if (rc != UMGAP_OK && idx) umgap_index_free(idx);
    return rc;
}
