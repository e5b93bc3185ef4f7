// Streaming reader for `fst` 0.3.x Map files (on-disk format version 2), the index format
// `umgap buildindex` writes (buildindex.rs:32-48) and the lookup commands open
// (prot2kmer2lca.rs:109-114).  The crate is a third-party dependency that is not vendored in the
// reference (Cargo.toml:23); this file implements its published node encoding:
//
//   file  = u64le version, u64le type, nodes..., u64le len, u64le root_addr
//   node  = fields laid out BEFORE its state byte (the node's address); decoded backwards
//   state = 11cccccc OneTransNext | 10cccccc OneTrans | 0fnnnnnn AnyTrans (f = final)
//
// The file is walked depth first exactly once; every key reaches the sink with the sum of the
// outputs along its path (fst::Map::stream semantics, printindex.rs:44-47).
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cerrno>

#include "index.h"

namespace umgap {

namespace {

const char kCommonInputsInv[] =
    "te/oasripcnw.hlm-du012g=:bf3y5&_4v9678k%?xCDASFIBEjPTzRNM+LOqHGWUV,YKJZXQ;)(~[]$!'*@";

inline uint64_t unpack(const uint8_t* p, unsigned size) {
    uint64_t v = 0;
    for (unsigned i = 0; i < size; ++i) v |= (uint64_t)p[i] << (8 * i);
    return v;
}

struct Node {
    const uint8_t* data;
    uint64_t addr;
    bool final_ = false;
    uint64_t final_output = 0;
    uint32_t ntrans = 0;
    // one-transition forms
    bool single = false;
    uint8_t s_input = 0;
    uint64_t s_output = 0, s_target = 0;
    // AnyTrans geometry
    uint64_t base = 0, start = 0;
    unsigned tsz = 0, osz = 0, isz = 0;

    Node(const uint8_t* d, uint64_t a, uint64_t size) : data(d), addr(a) {
        if (a == 0) {  // EMPTY_ADDRESS: final, no transitions, zero output
            final_ = true;
            return;
        }
        if (a >= size) UMGAP_FAIL(UMGAP_ERR_IO, "fst: node address out of range");
        const uint8_t s = d[a];
        const unsigned kind = s >> 6;
        if (kind >= 2) {
            single = true;
            ntrans = 1;
            const unsigned c = s & 0x3F;
            const unsigned il = c == 0 ? 1 : 0;
            s_input = il ? d[a - 1] : (uint8_t)kCommonInputsInv[c - 1];
            if (kind == 3) {  // OneTransNext: target is the node just before this one
                s_target = a - il - 1;
                return;
            }
            const uint8_t z = d[a - il - 1];
            const unsigned t = z >> 4, o = z & 15;
            if (t > 8 || o > 8 || a < il + 1 + t + o) UMGAP_FAIL(UMGAP_ERR_IO, "fst: corrupt OneTrans node");
            const uint64_t dpos = a - il - 1 - t;
            const uint64_t delta = unpack(d + dpos, t);
            const uint64_t st = dpos - o;
            s_output = o ? unpack(d + st, o) : 0;
            s_target = delta ? st - delta : 0;
            return;
        }
        final_ = (s & 0x40) != 0;
        uint32_t n = s & 0x3F;
        unsigned nl = 0;
        if (n == 0) {
            nl = 1;
            n = d[a - 1];
            if (n == 1) n = 256;
        }
        ntrans = n;
        base = a - nl - 1;
        const uint8_t z = d[base];
        tsz = z >> 4;
        osz = z & 15;
        if (tsz > 8 || osz > 8) UMGAP_FAIL(UMGAP_ERR_IO, "fst: corrupt AnyTrans node");
        isz = n > 32 ? 256 : 0;
        const uint64_t need = (uint64_t)isz + n + (uint64_t)n * tsz + (uint64_t)n * osz + (final_ ? osz : 0);
        if (base < need) UMGAP_FAIL(UMGAP_ERR_IO, "fst: corrupt AnyTrans node");
        start = base - need;
        final_output = (final_ && osz) ? unpack(d + start, osz) : 0;
    }

    void trans(uint32_t i, uint8_t& input, uint64_t& output, uint64_t& target) const {
        if (single) {
            input = s_input;
            output = s_output;
            target = s_target;
            return;
        }
        input = data[base - isz - 1 - i];
        const uint64_t delta = unpack(data + base - isz - ntrans - (uint64_t)(i + 1) * tsz, tsz);
        output = osz ? unpack(data + base - isz - ntrans - (uint64_t)ntrans * tsz - (uint64_t)(i + 1) * osz, osz) : 0;
        target = delta ? start - delta : 0;
    }
};

struct Mapped {
    const uint8_t* data = nullptr;
    size_t size = 0;
    int fd = -1;
    explicit Mapped(const char* path) {
        fd = ::open(path, O_RDONLY);
        if (fd < 0) UMGAP_FAIL(UMGAP_ERR_IO, "cannot open %s: %s", path, strerror(errno));
        struct stat st;
        if (fstat(fd, &st) != 0) UMGAP_FAIL(UMGAP_ERR_IO, "cannot stat %s", path);
        size = (size_t)st.st_size;
        if (size < 32) UMGAP_FAIL(UMGAP_ERR_IO, "%s is too short to be an fst", path);
        void* p = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (p == MAP_FAILED) UMGAP_FAIL(UMGAP_ERR_IO, "cannot mmap %s: %s", path, strerror(errno));
        data = (const uint8_t*)p;
        madvise(p, size, MADV_SEQUENTIAL);
    }
    ~Mapped() {
        if (data) munmap((void*)data, size);
        if (fd >= 0) ::close(fd);
    }
};

void check_header(const Mapped& m, uint64_t& len, uint64_t& root) {
    const uint64_t version = unpack(m.data, 8);
    // Version 1 lacks the 256-byte transition index of nodes with more than 32 transitions, which the
    // node decoder assumes; `fst` 0.3.5 (Cargo.toml:23) writes version 2 only.
    if (version != 2) UMGAP_FAIL(UMGAP_ERR_IO, "unsupported fst version %llu (expected 2)", (unsigned long long)version);
    len = unpack(m.data + m.size - 16, 8);
    root = unpack(m.data + m.size - 8, 8);
    if (!((root == 0 && m.size == 32) || root + 17 == m.size))
        UMGAP_FAIL(UMGAP_ERR_IO, "fst root address inconsistent with file length");
}

}  // namespace

uint64_t fst_file_len(const char* path) {
    Mapped m(path);
    uint64_t len, root;
    check_header(m, len, root);
    return len;
}

namespace {
struct Frame {
    Node node;
    uint32_t next;
    uint64_t out;
};

// Depth-first walk below `start` (reached by `key` with outputs summing to `out`), at most max_depth more bytes:
// keys go to the sink; with `cut`, nodes at max_depth are not entered but recorded as tasks.
void walk(const uint8_t* data, uint64_t node_limit, uint64_t start, uint64_t out0, std::vector<uint8_t>& key, FstSink& sink,
          size_t max_depth, std::vector<FstTask>* cut, std::vector<uint8_t>* bytes_seen) {
    const size_t base = key.size();
    std::vector<Frame> stack;
    stack.push_back(Frame{Node(data, start, node_limit), 0, out0});
    while (!stack.empty()) {
        Frame& f = stack.back();
        if (f.next >= f.node.ntrans) {
            stack.pop_back();
            if (key.size() > base) key.pop_back();
            continue;
        }
        uint8_t input;
        uint64_t output, target;
        f.node.trans(f.next++, input, output, target);
        const uint64_t out = f.out + output;
        key.push_back(input);
        if (bytes_seen) bytes_seen->push_back(input);
        if (key.size() > 4096) UMGAP_FAIL(UMGAP_ERR_IO, "fst: key longer than 4096 bytes (cycle?)");
        Node child(data, target, node_limit);
        if (child.final_) sink.on_key(key.data(), key.size(), out + child.final_output);
        if (child.ntrans == 0) {
            key.pop_back();
            continue;
        }
        if (cut && key.size() - base >= max_depth) {
            FstTask t;
            t.addr = target;
            t.out = out;
            t.plen = (uint8_t)key.size();
            memcpy(t.prefix, key.data(), key.size());
            cut->push_back(t);
            key.pop_back();
            continue;
        }
        stack.push_back(Frame{child, 0, out});
    }
}
}  // namespace

void fst_stream_file(const char* path, FstSink& sink, uint64_t* n_keys_footer) {
    Mapped m(path);
    uint64_t len, root;
    check_header(m, len, root);
    if (n_keys_footer) *n_keys_footer = len;
    if (m.size == 32) return;  // empty fst
    const uint64_t node_limit = m.size - 16;
    std::vector<uint8_t> key;
    Node r(m.data, root, node_limit);
    if (r.final_) sink.on_key(key.data(), 0, r.final_output);
    walk(m.data, node_limit, root, 0, key, sink, 0, nullptr, nullptr);
}

// ---- the same file walked by several threads: the top `depth` levels once (their keys go to `shallow`, the input
// bytes seen there to `bytes_seen` in walk order), the subtrees below as independent tasks ------------------------
struct FstFile::Impl {
    Mapped m;
    uint64_t len = 0, root = 0;
    explicit Impl(const char* path) : m(path) { check_header(m, len, root); }
};
FstFile::FstFile(const char* path) : impl_(new Impl(path)) {}
FstFile::~FstFile() { delete impl_; }
uint64_t FstFile::len() const { return impl_->len; }
void FstFile::split(size_t depth, std::vector<FstTask>& tasks, FstSink& shallow, std::vector<uint8_t>& bytes_seen) const {
    const Mapped& m = impl_->m;
    if (m.size == 32) return;
    if (depth > sizeof(FstTask().prefix)) depth = sizeof(FstTask().prefix);
    const uint64_t node_limit = m.size - 16;
    std::vector<uint8_t> key;
    Node r(m.data, impl_->root, node_limit);
    if (r.final_) shallow.on_key(key.data(), 0, r.final_output);
    walk(m.data, node_limit, impl_->root, 0, key, shallow, depth, &tasks, &bytes_seen);
}
void FstFile::stream(const FstTask& t, FstSink& sink) const {
    const Mapped& m = impl_->m;
    std::vector<uint8_t> key(t.prefix, t.prefix + t.plen);
    walk(m.data, m.size - 16, t.addr, t.out, key, sink, 0, nullptr, nullptr);
}

}  // namespace umgap
