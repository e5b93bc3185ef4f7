// Taxonomy loader: TSV -> preorder-numbered tree, ancestor matrix and snapping vectors in HBM.
// Mirrors taxon::read_taxa_file / Taxon::from_str (taxon.rs:89-128), TaxonTree::new (:224-247),
// TaxonList::new/ancestry (:135-163) and TaxonTree::snapping (:251-301); ranks rank.rs:9-78.
#include <algorithm>
#include <fstream>
#include <unordered_map>

#include "index.h"

namespace umgap {

static const char* kRanks[] = {
    "no rank", "superkingdom", "domain", "realm", "kingdom", "subkingdom", "superphylum", "phylum",
    "subphylum", "superclass", "class", "subclass", "infraclass", "superorder", "order", "suborder",
    "infraorder", "parvorder", "superfamily", "family", "subfamily", "tribe", "subtribe", "genus",
    "subgenus", "species group", "species subgroup", "species", "subspecies", "varietas", "forma",
    "strain"};

static bool parse_usize(const std::string& s, uint64_t& out) {
    // Rust usize::from_str: optional '+', then ASCII digits only.
    size_t i = 0;
    if (!s.empty() && s[0] == '+') i = 1;
    if (i >= s.size()) return false;
    unsigned __int128 v = 0;
    for (; i < s.size(); ++i) {
        if (s[i] < '0' || s[i] > '9') return false;
        v = v * 10 + (unsigned)(s[i] - '0');
        if (v > (unsigned __int128)UINT64_MAX) return false;
    }
    out = (uint64_t)v;
    return true;
}

static bool rust_is_whitespace(unsigned char c) {
    // ASCII subset of char::is_whitespace (trim_end on the line, taxon.rs:90)
    return c == ' ' || (c >= 0x09 && c <= 0x0D);
}

template <class T>
static const T* upload(umgap_taxonomy* tax, const std::vector<T>& v) {
    void* p = nullptr;
    const size_t bytes = std::max<size_t>(v.size() * sizeof(T), 16);
    UMGAP_CUDA(cudaMalloc(&p, bytes));
    tax->dev_allocs.push_back(p);
    if (!v.empty()) UMGAP_CUDA(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return (const T*)p;
}

static void build(umgap_taxonomy* tax, const std::vector<uint64_t>& ids,
                  const std::vector<uint64_t>& parents, const std::vector<uint8_t>& rank,
                  const std::vector<uint8_t>& valid) {
    const size_t n = ids.size();
    if (n == 0) UMGAP_FAIL(UMGAP_ERR_INVALID, "empty taxonomy");
    uint64_t max_id = 0;
    for (uint64_t id : ids) max_id = std::max(max_id, id);
    if (max_id >= 0xFFFFFFF0ull) UMGAP_FAIL(UMGAP_ERR_CAPACITY, "taxon ids must be below 2^32 - 16");
    // TaxonList::new: dense by id, a later line with the same id replaces an earlier one
    std::vector<int64_t> row_of(max_id + 1, -1);
    for (size_t i = 0; i < n; ++i) row_of[ids[i]] = (int64_t)i;
    // TaxonTree::new: children map over ALL lines, root = the id never listed with parent != id
    std::unordered_map<uint64_t, std::vector<uint64_t>> children;
    children.reserve(n * 2);
    std::vector<uint8_t> is_child(max_id + 1, 0);
    for (size_t i = 0; i < n; ++i) {
        if (ids[i] == parents[i]) continue;
        children[parents[i]].push_back(ids[i]);
        is_child[ids[i]] = 1;
    }
    uint64_t root = 0;
    size_t nroots = 0;
    for (uint64_t id = 0; id <= max_id; ++id)
        if (row_of[id] >= 0 && !is_child[id]) {
            if (nroots++ == 0) root = id;
        }
    if (nroots > 1) UMGAP_FAIL(UMGAP_ERR_INVALID, "More than one root!");
    if (nroots == 0) UMGAP_FAIL(UMGAP_ERR_INVALID, "There's no root!");

    // Preorder numbering from the root (iterative DFS; children in ascending id order).
    std::vector<uint32_t> dense_of(max_id + 1, kNoTaxon), id_of, last, parent;
    std::vector<uint8_t> depth;
    id_of.reserve(n);
    struct Frame {
        uint64_t id;
        uint32_t dense;
        size_t next;
        std::vector<uint64_t>* ch;
    };
    std::vector<Frame> stack;
    auto open = [&](uint64_t id, uint32_t par_dense, uint32_t d) {
        if (id > max_id || row_of[id] < 0) return;       // child line whose id has no entry: cannot happen
        if (dense_of[id] != kNoTaxon) return;            // duplicate line / cycle guard
        if (d > 254) UMGAP_FAIL(UMGAP_ERR_CAPACITY, "taxonomy deeper than 254 levels");
        const uint32_t me = (uint32_t)id_of.size();
        dense_of[id] = me;
        id_of.push_back((uint32_t)id);
        last.push_back(me);
        parent.push_back(par_dense == kNoTaxon ? me : par_dense);
        depth.push_back((uint8_t)d);
        auto it = children.find(id);
        std::vector<uint64_t>* ch = nullptr;
        if (it != children.end()) {
            std::sort(it->second.begin(), it->second.end());
            ch = &it->second;
        }
        stack.push_back(Frame{id, me, 0, ch});
    };
    open(root, kNoTaxon, 0);
    while (!stack.empty()) {
        Frame& f = stack.back();
        if (f.ch && f.next < f.ch->size()) {
            const uint64_t c = (*f.ch)[f.next++];
            open(c, f.dense, (uint32_t)depth[f.dense] + 1);
        } else {
            const uint32_t me = f.dense;
            const uint32_t end = (uint32_t)id_of.size() - 1;
            last[me] = end;
            stack.pop_back();
        }
    }
    const uint32_t nd = (uint32_t)id_of.size();
    uint32_t max_depth = 0;
    for (uint8_t d : depth) max_depth = std::max<uint32_t>(max_depth, d);
    const uint32_t stride = ((max_depth + 1 + 7) / 8) * 8;
    std::vector<uint32_t> anc((size_t)nd * stride, kNoTaxon);
    for (uint32_t x = 0; x < nd; ++x) {
        uint32_t* row = &anc[(size_t)x * stride];
        const uint32_t d = depth[x];
        if (d > 0) memcpy(row, &anc[(size_t)parent[x] * stride], d * sizeof(uint32_t));
        row[d] = x;
    }
    // snapping: nearest ancestor-or-self passing the filter; the root maps to itself (taxon.rs:279)
    std::vector<uint32_t> snap_valid(nd), snap_ranked(nd);
    for (uint32_t x = 0; x < nd; ++x) {  // preorder: parents come first
        const int64_t r = row_of[id_of[x]];
        const bool v = valid[r] != 0, rk = v && rank[r] != 0;
        snap_valid[x] = v ? id_of[x] : (x == 0 ? id_of[0] : snap_valid[parent[x]]);
        snap_ranked[x] = rk ? id_of[x] : (x == 0 ? id_of[0] : snap_ranked[parent[x]]);
    }

    // seedextend -r: the rank score of a taxon is that of its nearest ancestor-or-self with a rank (the root ends the
    // walk whatever its rank), taxon.rs:181-191.  Rank::score (rank.rs:86-99) is a ladder of `self < X` tests whose first
    // rung, `self < Species`, already holds for every rank above species: those score 12, species and below fall
    // through every rung to None, and so does "no rank" (its comparisons are all false).  Restated as written.
    std::vector<uint8_t> seed_score(nd), eff_rank(nd);
    for (uint32_t x = 0; x < nd; ++x) {
        const uint8_t rk = rank[row_of[id_of[x]]];
        eff_rank[x] = (x == 0 || rk != 0) ? rk : eff_rank[parent[x]];
        seed_score[x] = (eff_rank[x] != 0 && eff_rank[x] < 27 /* Rank::Species */) ? 12 : 0;
    }

    tax->ids = ids;
    tax->parents = parents;
    tax->ranks = rank;
    tax->valids = valid;
    tax->root = root;
    tax->max_id = max_id;
    tax->max_depth = max_depth;
    tax->h_dense_of = dense_of;
    tax->h_id_of = id_of;
    tax->h_parent = parent;
    tax->h_depth = depth;

    use_device(tax->device);
    TaxView& v = tax->view;
    v.dense_of = upload(tax, dense_of);
    v.id_of = upload(tax, id_of);
    v.last = upload(tax, last);
    v.parent = upload(tax, parent);
    v.depth = upload(tax, depth);
    v.anc = upload(tax, anc);
    v.snap_valid = upload(tax, snap_valid);
    v.snap_ranked = upload(tax, snap_ranked);
    v.seed_score = upload(tax, seed_score);
    v.n = nd;
    v.max_id = (uint32_t)max_id;
    v.stride = stride;
}

}  // namespace umgap

using namespace umgap;

extern "C" {

int umgap_taxonomy_from_arrays(const uint64_t* ids, const uint64_t* parents, const uint8_t* rank,
                               const uint8_t* valid, uint64_t n, int device, umgap_taxonomy** out) {
    umgap_taxonomy* tax = nullptr;
    int rc = guarded([&] {
        if (!ids || !parents || !rank || !valid || !out) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        tax = new umgap_taxonomy();
        tax->device = device;
        build(tax, std::vector<uint64_t>(ids, ids + n), std::vector<uint64_t>(parents, parents + n),
              std::vector<uint8_t>(rank, rank + n), std::vector<uint8_t>(valid, valid + n));
        *out = tax;
    });
    if (rc != UMGAP_OK && tax) umgap_taxonomy_free(tax);
    return rc;
}

int umgap_taxonomy_replicate(const umgap_taxonomy* src, int device, umgap_taxonomy** out) {
    if (!src || !out) {
        set_error("null argument");
        return UMGAP_ERR_INVALID;
    }
    return umgap_taxonomy_from_arrays(src->ids.data(), src->parents.data(), src->ranks.data(), src->valids.data(),
                                      src->ids.size(), device, out);
}

int umgap_taxonomy_load(const char* tsv_path, int device, umgap_taxonomy** out) {
    umgap_taxonomy* tax = nullptr;
    int rc = guarded([&] {
        if (!tsv_path || !out) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        std::ifstream f(tsv_path, std::ios::binary);
        if (!f) UMGAP_FAIL(UMGAP_ERR_IO, "Failed opening taxon file.");
        std::vector<uint64_t> ids, parents;
        std::vector<uint8_t> rank, valid;
        std::string line;
        while (std::getline(f, line)) {
            if (!line.empty() && line.back() == '\r') line.pop_back();  // BufRead::lines
            while (!line.empty() && rust_is_whitespace((unsigned char)line.back())) line.pop_back();
            std::vector<std::string> split;
            size_t start = 0;
            for (;;) {
                const size_t p = line.find('\t', start);
                if (p == std::string::npos) {
                    split.push_back(line.substr(start));
                    break;
                }
                split.push_back(line.substr(start, p - start));
                start = p + 1;
            }
            if (split.size() != 5) UMGAP_FAIL(UMGAP_ERR_IO, "Taxon requires five fields");
            uint64_t id, parent;
            if (!parse_usize(split[0], id)) UMGAP_FAIL(UMGAP_ERR_IO, "invalid digit found in string");
            int r = -1;
            for (int i = 0; i < 32; ++i)
                if (split[2] == kRanks[i]) r = i;
            if (r < 0) UMGAP_FAIL(UMGAP_ERR_IO, "Matching variant not found");
            if (!parse_usize(split[3], parent)) UMGAP_FAIL(UMGAP_ERR_IO, "invalid digit found in string");
            bool v;
            if (split[4] == std::string("\x01", 1)) v = true;
            else if (split[4] == std::string("\x00", 1)) v = false;
            else UMGAP_FAIL(UMGAP_ERR_IO, "Couldn't parse the valid byte");
            ids.push_back(id);
            parents.push_back(parent);
            rank.push_back((uint8_t)r);
            valid.push_back(v);
        }
        tax = new umgap_taxonomy();
        tax->device = device;
        build(tax, ids, parents, rank, valid);
        *out = tax;
    });
    if (rc != UMGAP_OK && tax) umgap_taxonomy_free(tax);
    return rc;
}

void umgap_taxonomy_free(umgap_taxonomy* tax) {
    if (!tax) return;
    cudaSetDevice(tax->device);
    for (void* p : tax->dev_allocs) cudaFree(p);
    delete tax;
}

int umgap_taxonomy_get_info(const umgap_taxonomy* tax, umgap_taxonomy_info* info) {
    return guarded([&] {
        if (!tax || !info) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        info->n_taxa = tax->view.n;
        info->max_id = tax->max_id;
        info->root = tax->root;
        info->max_depth = tax->max_depth;
        info->device = tax->device;
    });
}

}  // extern "C"
