// Writer for `fst` 0.3.x Map files (format version 2): what `umgap buildindex` produces from its sorted TSV
// (buildindex.rs:32-48, fst::MapBuilder::{new,insert,finish}), and the key stream of `umgap printindex`
// (printindex.rs:38-51).  Host code only.  The byte format is the one fst_stream.cu reads:
//
//   file  = u64le 2 (version), u64le 0 (type Map), nodes..., u64le number of keys, u64le root address
//   node  = fields written before its state byte; the node's address is that of the state byte
//
// Construction is the classic incremental one for sorted input: the path of the last key stays unfrozen, a new
// key freezes everything below its common prefix, outputs are pushed down so that the outputs along a path sum
// to the key's value.  Frozen nodes are written at once; no node registry is kept (the crate shares equal suffixes
// through a bounded one), so files come out larger than the crate's, never different in content: every reader
// that follows the format returns the same key/value pairs.  Parity of the BYTES with files written by the real
// crate is unpinned (no such file exists offline).
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "index.h"

namespace umgap {
namespace {

// fst's COMMON_INPUTS, inverse direction (index + 1 in the six low bits of a one-transition state byte)
const char kCommonInv[] = "te/oasripcnw.hlm-du012g=:bf3y5&_4v9678k%?xCDASFIBEjPTzRNM+LOqHGWUV,YKJZXQ;)(~[]$!'*@";

unsigned common_code(uint8_t byte) {  // 0: not a common input
    const char* p = byte ? (const char*)memchr(kCommonInv, byte, sizeof kCommonInv - 1) : nullptr;
    const unsigned code = p ? (unsigned)(p - kCommonInv) + 1 : 0;
    return code <= 63 ? code : 0;
}

unsigned bytes_for(uint64_t v) {
    unsigned n = 1;
    while (n < 8 && (v >> (8 * n)) != 0) ++n;
    return n;
}

struct Arc {
    uint8_t input;
    uint64_t output, target;
};
struct Pending {             // a node on the unfrozen path
    bool final_ = false;
    uint64_t final_output = 0;
    std::vector<Arc> arcs;   // frozen transitions
    bool open = false;       // the transition the path continues on (target not known yet)
    uint8_t open_input = 0;
    uint64_t open_output = 0;
    void reset() {           // a node of the path is reused with the room its transitions already have
        final_ = false;
        final_output = 0;
        arcs.clear();
        open = false;
        open_input = 0;
        open_output = 0;
    }
};

}  // namespace
}  // namespace umgap

using namespace umgap;

struct umgap_fst_writer {
    FILE* f = nullptr;
    bool owns = false;
    uint64_t pos = 0;          // bytes written so far
    uint64_t last_addr = 0;    // address of the node written last
    uint64_t nkeys = 0;
    std::vector<Pending> path; // path[i], i < plen, is reached by the first i bytes of the last key; the rest is spare
    size_t plen = 0;
    std::string last_key;
    bool any = false;
    std::vector<uint8_t> buf;  // bytes not yet handed to the file (nodes are a few bytes each)

    void push_node() {
        if (plen == path.size()) path.emplace_back();
        else path[plen].reset();
        ++plen;
    }
    void flush_buf() {
        if (!buf.empty() && fwrite(buf.data(), 1, buf.size(), f) != buf.size()) UMGAP_FAIL(UMGAP_ERR_IO, "failed writing the fst");
        buf.clear();
    }
    void put(const void* p, size_t n) {
        if (buf.size() + n > (1u << 20)) flush_buf();
        buf.insert(buf.end(), (const uint8_t*)p, (const uint8_t*)p + n);
        pos += n;
    }
    void put_le(uint64_t v, unsigned n) {
        uint8_t b[8];
        for (unsigned i = 0; i < n; ++i) b[i] = (uint8_t)(v >> (8 * i));
        put(b, n);
    }
    // Writes a frozen node, returns its address (0 = the shared empty final node, which is never written).
    uint64_t emit(const Pending& nd) {
        if (nd.final_ && nd.arcs.empty() && nd.final_output == 0) return 0;
        const uint64_t start = pos;
        if (!nd.final_ && nd.arcs.size() == 1) {
            const Arc& a = nd.arcs[0];
            const unsigned code = common_code(a.input);
            if (a.output == 0 && a.target != 0 && a.target == last_addr && start == a.target + 1) {  // OneTransNext
                if (!code) put(&a.input, 1);
                const uint8_t st = (uint8_t)(0xC0 | code);
                put(&st, 1);
                return pos - 1;
            }
            const uint64_t delta = a.target ? start - a.target : 0;  // OneTrans
            const unsigned osz = a.output ? bytes_for(a.output) : 0, tsz = bytes_for(delta);
            if (osz) put_le(a.output, osz);
            put_le(delta, tsz);
            const uint8_t sizes = (uint8_t)(tsz << 4 | osz);
            put(&sizes, 1);
            if (!code) put(&a.input, 1);
            const uint8_t st = (uint8_t)(0x80 | code);
            put(&st, 1);
            return pos - 1;
        }
        // AnyTrans: [final output][outputs, last transition first][deltas][inputs][256-byte index][sizes][count][state]
        const size_t n = nd.arcs.size();
        unsigned tsz = 1, osz = 0;
        bool any_out = nd.final_ && nd.final_output;
        for (const Arc& a : nd.arcs) {
            any_out |= a.output != 0;
            tsz = std::max(tsz, bytes_for(a.target ? start - a.target : 0));
        }
        if (any_out) {
            osz = nd.final_ ? bytes_for(nd.final_output) : 1;
            for (const Arc& a : nd.arcs) osz = std::max(osz, bytes_for(a.output));
        }
        if (nd.final_ && osz) put_le(nd.final_output, osz);
        if (osz)
            for (size_t i = n; i-- > 0;) put_le(nd.arcs[i].output, osz);
        for (size_t i = n; i-- > 0;) put_le(nd.arcs[i].target ? start - nd.arcs[i].target : 0, tsz);
        for (size_t i = n; i-- > 0;) put(&nd.arcs[i].input, 1);
        if (n > 32) {
            uint8_t index[256];
            memset(index, 0xFF, sizeof index);
            for (size_t i = 0; i < n; ++i) index[nd.arcs[i].input] = (uint8_t)i;
            put(index, sizeof index);
        }
        const uint8_t sizes = (uint8_t)(tsz << 4 | osz);
        put(&sizes, 1);
        uint8_t st = nd.final_ ? 0x40 : 0;
        if (n >= 1 && n <= 63) {
            st |= (uint8_t)n;
        } else {  // the count has a byte of its own; 256 is written as 1 (a one-transition node never takes this form)
            const uint8_t cnt = n == 256 ? 1 : (uint8_t)n;
            put(&cnt, 1);
        }
        put(&st, 1);
        return pos - 1;
    }
    // Freezes path[keep + 1 ..]: deepest first, each linked into its parent's open transition.
    void freeze_below(size_t keep) {
        while (plen > keep + 1) {
            Pending& nd = path[plen - 1];
            const uint64_t addr = emit(nd);
            if (addr) last_addr = addr;
            --plen;
            Pending& parent = path[plen - 1];
            parent.arcs.push_back(Arc{parent.open_input, parent.open_output, addr});
            parent.open = false;
        }
    }
};

extern "C" {

int umgap_fst_writer_open(const char* path, umgap_fst_writer** out) {
    umgap_fst_writer* w = nullptr;
    int rc = guarded([&] {
        if (!out) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        w = new umgap_fst_writer();
        if (path && strcmp(path, "-") != 0) {
            w->f = fopen(path, "wb");
            if (!w->f) UMGAP_FAIL(UMGAP_ERR_IO, "cannot create %s", path);
            w->owns = true;
        } else {
            w->f = stdout;
        }
        w->buf.reserve((1u << 20) + 512);
        w->put_le(2, 8);
        w->put_le(0, 8);
        w->push_node();
        *out = w;
    });
    if (rc != UMGAP_OK && w) {
        if (w->owns && w->f) fclose(w->f);
        delete w;
    }
    return rc;
}

int umgap_fst_writer_insert(umgap_fst_writer* w, const uint8_t* key, size_t len, uint64_t value) {
    return guarded([&] {
        if (!w || (len && !key)) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        // fst::MapBuilder::insert: keys must arrive in strictly increasing lexicographic order
        if (w->any) {
            const size_t m = std::min(len, w->last_key.size());
            const int c = memcmp(w->last_key.data(), key, m);
            if (c > 0 || (c == 0 && w->last_key.size() >= len)) {
                if (c == 0 && w->last_key.size() == len) UMGAP_FAIL(UMGAP_ERR_INVALID, "fst: duplicate key '%.*s'", (int)len, (const char*)key);
                UMGAP_FAIL(UMGAP_ERR_INVALID, "fst: keys out of order: '%.*s' after '%s'", (int)len, (const char*)key, w->last_key.c_str());
            }
        }
        // walk the common prefix, leaving on each shared transition the part of its output the new value also has
        size_t p = 0;
        uint64_t rest = value;
        while (p < len && p + 1 < w->plen && w->path[p].open && w->path[p].open_input == key[p]) {
            Pending& nd = w->path[p];
            const uint64_t common = std::min(nd.open_output, rest);
            const uint64_t push = nd.open_output - common;
            if (push) {  // the remainder moves one node down: onto every way out of the child
                nd.open_output = common;
                Pending& child = w->path[p + 1];
                if (child.final_) child.final_output += push;
                for (Arc& a : child.arcs) a.output += push;
                if (child.open) child.open_output += push;
            }
            rest -= common;
            ++p;
        }
        w->freeze_below(p);
        if (len == 0) {  // the empty key can only come first
            w->path[0].final_ = true;
            w->path[0].final_output = value;
        } else {
            for (size_t i = p; i < len; ++i) {
                Pending& nd = w->path[i];
                nd.open = true;
                nd.open_input = key[i];
                nd.open_output = i == p ? rest : 0;
                w->push_node();
            }
            w->path[w->plen - 1].final_ = true;
        }
        w->last_key.assign((const char*)key, len);
        w->any = true;
        ++w->nkeys;
    });
}

int umgap_fst_writer_finish(umgap_fst_writer* w) {
    if (!w) return UMGAP_OK;
    int rc = guarded([&] {
        w->freeze_below(0);
        const uint64_t root = w->emit(w->path[0]);
        w->put_le(w->nkeys, 8);
        w->put_le(root, 8);
        w->flush_buf();
        if (fflush(w->f) != 0) UMGAP_FAIL(UMGAP_ERR_IO, "failed writing the fst");
    });
    if (w->owns && w->f) fclose(w->f);
    delete w;
    return rc;
}

void umgap_fst_writer_abort(umgap_fst_writer* w) {
    if (!w) return;
    if (w->owns && w->f) fclose(w->f);
    delete w;
}

// fst::Map::stream (printindex.rs:44-47): every key in lexicographic order with its value.
int umgap_fst_stream(const char* path, umgap_fst_key_fn fn, void* user, uint64_t* n_keys) {
    return guarded([&] {
        if (!path || !fn) UMGAP_FAIL(UMGAP_ERR_INVALID, "null argument");
        struct Sink : FstSink {
            umgap_fst_key_fn fn;
            void* user;
            void on_key(const uint8_t* key, size_t len, uint64_t value) override {
                if (fn(key, len, value, user) != 0) UMGAP_FAIL(UMGAP_ERR_IO, "key callback failed");
            }
        } sink;
        sink.fn = fn;
        sink.user = user;
        fst_stream_file(path, sink, n_keys);
    });
}

}  // extern "C"
