// `umgap` command line on top of libumgap_gpu.so: the reference's argv surface and FASTA stream
// formats for the hot-path subcommands (src/main.rs:8-63; flags per SURVEY Appendix C).
//
//   translate      src/commands/translate.rs:46-133      -> umgap_translate
//   prot2kmer2lca  src/commands/prot2kmer2lca.rs:70-193  -> umgap_index_load_fst + umgap_kmer_lookup
//   prot2tryp2lca  src/commands/prot2tryp2lca.rs:41-140  -> umgap_index_load_fst(k=0) + umgap_tryp_lookup
//   seedextend     src/commands/seedextend.rs:63-179     -> umgap_seedextend
//   taxa2agg       src/commands/taxa2agg.rs:61-183       -> umgap_taxonomy_load + umgap_aggregate
//   uniq           src/commands/uniq.rs:41-84            (host only)
//   fastq2fasta    src/commands/fastq2fasta.rs:55-84     (host only)
//   buildindex     src/commands/buildindex.rs:32-48      -> umgap_fst_writer_* (host only)
//   printindex     src/commands/printindex.rs:38-51      -> umgap_fst_stream (host only)
//   snaptaxon      src/commands/snaptaxon.rs:66-108      (host only)
//   taxa2freq      src/commands/taxa2freq.rs:86-169      (host only)
//   bestof         src/commands/bestof.rs:50-79          (host only)
//   classify       the fused preset pipeline (extension) -> umgap_classify_reads, one replica per GPU
//
// Stream rules follow src/io/fasta.rs:30-67,164-180.  Errors go to stderr as `Error: ...`, exit 1
// (quick_main!, src/main.rs:8).  The host does parsing and formatting only; every stage's
// arithmetic runs on the GPU through the C ABI.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/un.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "umgap_gpu.h"

namespace {

struct Fail : std::runtime_error {
    using std::runtime_error::runtime_error;
};
[[noreturn]] void fail(const std::string& m) { throw Fail(m); }
void check(int rc) {
    if (rc != UMGAP_OK) fail(umgap_last_error());
}

// ---- FASTA stream (src/io/fasta.rs) -------------------------------------------------------------
struct Record {
    std::string header;
    std::vector<std::string> seq;
};

class LineSource {
  public:
    explicit LineSource(FILE* f) : f_(f), buf_(1 << 20) {}
    bool next(std::string& line) {  // BufRead::lines(): strips "\n" or "\r\n"
        line.clear();
        bool any = false;
        for (;;) {
            if (pos_ == len_) {
                len_ = fread(buf_.data(), 1, buf_.size(), f_);
                pos_ = 0;
                if (len_ == 0) break;
            }
            const char* p = (const char*)memchr(buf_.data() + pos_, '\n', len_ - pos_);
            if (p) {
                line.append(buf_.data() + pos_, p - (buf_.data() + pos_));
                pos_ = (p - buf_.data()) + 1;
                if (!line.empty() && line.back() == '\r') line.pop_back();
                return true;
            }
            line.append(buf_.data() + pos_, len_ - pos_);
            pos_ = len_;
            any = true;
        }
        return any || !line.empty();
    }

  private:
    FILE* f_;
    std::vector<char> buf_;
    size_t pos_ = 0, len_ = 0;
};

class FastaReader {
  public:
    FastaReader(FILE* f, bool unwrap) : src_(f), unwrap_(unwrap) { have_ = src_.next(pending_); }
    bool next(Record& r) {  // fasta.rs:38-67
        if (!have_) return false;
        if (pending_.empty() || pending_[0] != '>') fail("Expected > at beginning of fasta header.");
        r.header.assign(pending_, 1, std::string::npos);
        r.seq.clear();
        while ((have_ = src_.next(pending_)) && !(pending_.size() && pending_[0] == '>')) r.seq.push_back(pending_);
        if (unwrap_) {
            std::string all;
            for (auto& s : r.seq) all += s;
            r.seq.assign(1, all);
        }
        return true;
    }

  private:
    LineSource src_;
    bool unwrap_, have_;
    std::string pending_;
};

void write_record(std::string& out, const std::string& header, const std::vector<std::string>& items,
                  const std::string& sep, bool wrap) {  // fasta.rs:164-180
    out += '>';
    out += header;
    std::string s;
    for (size_t i = 0; i < items.size(); ++i) {
        if (i) s += sep;
        s += items[i];
    }
    if (!wrap) {
        out += '\n';
        out += s;
    } else {
        for (size_t k = 0; k < s.size(); k += 70) {
            out += '\n';
            out.append(s, k, 70);
        }
    }
    if (!s.empty()) out += '\n';
}

void put(FILE* f, const std::string& s) {
    if (!s.empty() && fwrite(s.data(), 1, s.size(), f) != s.size()) fail("failed writing output");
}

// ---- block-wise reading: the hot stages parse the stream in place, without a string per line ------
// Hands out regions of the stream that hold whole records (a record starts at a line that begins with '>').
class BlockReader {
  public:
    explicit BlockReader(FILE* f, size_t block = 32u << 20) : f_(f), buf_(block) {}
    bool next(const char*& p, const char*& end) {
        memmove(buf_.data(), buf_.data() + used_, have_ - used_);
        have_ -= used_;
        used_ = 0;
        for (;;) {
            if (!eof_) {
                if (have_ == buf_.size()) buf_.resize(buf_.size() * 2);  // one record larger than the block
                const size_t n = fread(buf_.data() + have_, 1, buf_.size() - have_, f_);
                if (n == 0) eof_ = true;
                have_ += n;
            }
            if (first_ && have_) {
                if (buf_[0] != '>') fail("Expected > at beginning of fasta header.");  // fasta.rs:44-49
                first_ = false;
            }
            size_t limit = have_;
            if (!eof_) {  // everything before the last line that starts with '>'
                limit = 0;
                for (size_t q = have_; q > 1;) {
                    const void* m = memrchr(buf_.data(), '>', q);
                    if (!m) break;
                    const size_t at = (const char*)m - buf_.data();
                    if (at > 0 && buf_[at - 1] == '\n') {
                        limit = at;
                        break;
                    }
                    q = at;
                }
                if (limit == 0) continue;  // no complete record yet: read on
            }
            if (limit == 0) return false;
            p = buf_.data();
            end = buf_.data() + limit;
            used_ = limit;
            return true;
        }
    }

  private:
    FILE* f_;
    std::vector<char> buf_;
    size_t have_ = 0, used_ = 0;
    bool eof_ = false, first_ = true;
};

// The next line of [p, end) without its "\n" / "\r\n" (BufRead::lines()); advances p.
inline void take_line(const char*& p, const char* end, const char*& ls, size_t& ll) {
    const char* e = (const char*)memchr(p, '\n', end - p);
    const char* le = e ? e : end;
    ls = p;
    ll = le - p;
    if (e && ll && ls[ll - 1] == '\r') --ll;  // "\r\n" is a line end, a "\r" before the end of the stream is not
    p = e ? e + 1 : end;
}

// Decimal taxon id of one line (Rust usize::from_str, then the 32-bit range of the tables).
inline uint32_t taxon_of_line(const char* s, size_t n) {
    size_t i = (n && s[0] == '+') ? 1 : 0;
    if (i >= n) fail(n == 0 ? "cannot parse integer from empty string" : "invalid digit found in string");
    uint64_t v = 0;
    for (; i < n; ++i) {
        const unsigned d = (unsigned)(s[i] - '0');
        if (d > 9) fail("invalid digit found in string");
        if (v > (UINT64_MAX - d) / 10) fail("number too large to fit in target type");
        v = v * 10 + d;
    }
    if (v >= 0xFFFFFFFFull) fail("taxon id " + std::string(s, n) + " does not fit 32 bits");
    return (uint32_t)v;
}

inline void append_u32(std::string& out, uint32_t v) {
    char tmp[10];
    int n = 0;
    do {
        tmp[n++] = (char)('0' + v % 10);
        v /= 10;
    } while (v);
    while (n) out += tmp[--n];
}

// Decimal digits of v at w (room for 10), two at a time; returns the end.
inline char* put_u32(char* w, uint32_t v) {
    static const char kPairs[201] =
        "00010203040506070809101112131415161718192021222324252627282930313233343536373839404142434445464748495051525354555657585960616263646566676869707172737475767778798081828384858687888990919293949596979899";
    if (v < 10) {
        *w = (char)('0' + v);
        return w + 1;
    }
    char tmp[10];
    int n = 10;
    while (v >= 100) {
        const uint32_t q = v / 100, r = v - q * 100;
        tmp[--n] = kPairs[2 * r + 1];
        tmp[--n] = kPairs[2 * r];
        v = q;
    }
    if (v >= 10) {
        tmp[--n] = kPairs[2 * v + 1];
        tmp[--n] = kPairs[2 * v];
    } else {
        tmp[--n] = (char)('0' + v);
    }
    memcpy(w, tmp + n, 10 - n);
    return w + (10 - n);
}

// The records of a block with their sequence lines joined (fasta::Reader with unwrap, fasta.rs:38-67): the sequences
// one after the other in `seq`, off[i] .. off[i + 1] the i-th, the headers as views into the block.
struct UnwrappedBlock {
    std::string seq;
    std::vector<uint64_t> off;
    std::vector<std::pair<const char*, size_t>> heads;
    void parse(const char* p, const char* end) {
        seq.clear();
        off.assign(1, 0);
        heads.clear();
        while (p < end) {
            const char* ls;
            size_t ll;
            take_line(p, end, ls, ll);
            heads.emplace_back(ls + 1, ll - 1);
            while (p < end && *p != '>') {
                take_line(p, end, ls, ll);
                seq.append(ls, ll);
            }
            off.push_back(seq.size());
        }
    }
    size_t size() const { return heads.size(); }
};

// A batch of records whose items are taxon ids, one per line (the streams between prot2kmer2lca, seedextend, uniq
// and taxa2agg): headers in one arena, ids flattened, CSR offsets.
struct IdBatch {
    std::string harena;
    std::vector<uint64_t> hoff{0}, off{0};
    std::vector<uint32_t> ids;
    size_t size() const { return off.size() - 1; }
    void clear() {
        harena.clear();
        hoff.assign(1, 0);
        off.assign(1, 0);
        ids.clear();
    }
};
// Parses the records of [p, end) into b (appending); stops after `max_records`.  Returns where it stopped.
inline const char* parse_id_records(const char* p, const char* end, IdBatch& b, size_t max_records) {
    while (p < end && b.size() < max_records) {
        const char* ls;
        size_t ll;
        take_line(p, end, ls, ll);  // header line: '>' + header
        b.harena.append(ls + 1, ll - 1);
        b.hoff.push_back(b.harena.size());
        while (p < end && *p != '>') {
            take_line(p, end, ls, ll);
            b.ids.push_back(taxon_of_line(ls, ll));
        }
        b.off.push_back(b.ids.size());
    }
    return p;
}

// ---- argv ---------------------------------------------------------------------------------------
struct Spec {
    char shortf;
    const char* longf;
    bool takes_value;
};
struct Args {
    std::map<std::string, std::vector<std::string>> opt;  // keyed by long name
    std::vector<std::string> pos;
    bool has(const char* k) const { return opt.count(k) != 0; }
    std::string get(const char* k, const std::string& dflt) const {
        auto it = opt.find(k);
        return it == opt.end() ? dflt : it->second.back();
    }
};

Args parse(int argc, char** argv, int from, const std::vector<Spec>& specs) {
    Args a;
    for (int i = from; i < argc; ++i) {
        std::string t = argv[i];
        if (t == "--") {
            for (++i; i < argc; ++i) a.pos.push_back(argv[i]);
            break;
        }
        const Spec* sp = nullptr;
        std::string inline_val;
        bool has_inline = false;
        if (t.size() > 2 && t[0] == '-' && t[1] == '-') {
            std::string name = t.substr(2);
            const size_t eq = name.find('=');
            if (eq != std::string::npos) {
                inline_val = name.substr(eq + 1);
                name = name.substr(0, eq);
                has_inline = true;
            }
            for (auto& s : specs)
                if (name == s.longf) sp = &s;
            if (!sp) fail("Found argument '" + t + "' which wasn't expected, or isn't valid in this context");
        } else if (t.size() >= 2 && t[0] == '-' && !(t[1] >= '0' && t[1] <= '9')) {
            for (auto& s : specs)
                if (t[1] == s.shortf) sp = &s;
            if (!sp) fail("Found argument '" + t + "' which wasn't expected, or isn't valid in this context");
            if (t.size() > 2) {
                if (sp->takes_value) {
                    inline_val = t.substr(t[2] == '=' ? 3 : 2);
                    has_inline = true;
                } else {  // bundled flags, e.g. -mo
                    a.opt[sp->longf].push_back("");
                    for (size_t k = 2; k < t.size(); ++k) {
                        const Spec* q = nullptr;
                        for (auto& s : specs)
                            if (t[k] == s.shortf) q = &s;
                        if (!q || q->takes_value) fail("Found argument '" + t + "' which wasn't expected, or isn't valid in this context");
                        a.opt[q->longf].push_back("");
                    }
                    continue;
                }
            }
        } else {
            a.pos.push_back(t);
            continue;
        }
        if (sp->takes_value) {
            if (!has_inline) {
                if (i + 1 >= argc) fail(std::string("The argument '--") + sp->longf + " <" + sp->longf + ">' requires a value but none was supplied");
                inline_val = argv[++i];
            }
            a.opt[sp->longf].push_back(inline_val);
        } else {
            a.opt[sp->longf].push_back("");
        }
    }
    return a;
}

uint64_t parse_usize(const std::string& s) {  // Rust usize::from_str
    size_t i = 0;
    if (!s.empty() && s[0] == '+') i = 1;
    if (i >= s.size()) fail(s.empty() ? "cannot parse integer from empty string" : "invalid digit found in string");
    unsigned __int128 v = 0;
    for (; i < s.size(); ++i) {
        if (s[i] < '0' || s[i] > '9') fail("invalid digit found in string");
        v = v * 10 + (unsigned)(s[i] - '0');
        if (v > (unsigned __int128)UINT64_MAX) fail("number too large to fit in target type");
    }
    return (uint64_t)v;
}
float parse_f32(const std::string& s) {
    char* end = nullptr;
    errno = 0;
    const float v = strtof(s.c_str(), &end);
    if (s.empty() || *end) fail("invalid float literal");
    return v;
}

struct IndexHandle {
    umgap_index* p = nullptr;
    ~IndexHandle() { umgap_index_free(p); }
};
struct TaxHandle {
    umgap_taxonomy* p = nullptr;
    ~TaxHandle() { umgap_taxonomy_free(p); }
};

const size_t kBatchRecords = 1 << 16;

// ---- translate ----------------------------------------------------------------------------------
int cmd_translate(int argc, char** argv) {
    Args a = parse(argc, argv, 2, {{'m', "methionine", false}, {'a', "all-frames", false}, {'f', "frame", true},
                                   {'n', "append-name", false}, {'t', "table", true}, {'s', "show-table", false}});
    if (a.has("all-frames") && a.has("frame")) fail("The argument '--all-frames' cannot be used with '--frame <frames>...'");
    static const char* names[6] = {"1", "2", "3", "1R", "2R", "3R"};
    const int table = (int)parse_usize(a.get("table", "1"));
    const int meth = a.has("methionine");
    // frames in the order given (translate.rs:82-93); the C ABI emits in reference order, so map
    std::vector<int> order;
    if (a.has("all-frames")) {
        for (int i = 0; i < 6; ++i) order.push_back(i);
    } else if (a.has("frame")) {
        for (auto& f : a.opt["frame"]) {
            int k = -1;
            for (int i = 0; i < 6; ++i)
                if (f == names[i]) k = i;
            if (k < 0) fail("Invalid frame: " + f);
            order.push_back(k);
        }
    }
    if (a.has("show-table")) {  // TranslationTable::print (dna/translation.rs:146-173)
        static const std::map<int, const char*> table_names = {
            {1, "universal"}, {2, "vertebrate_mitochondrial"}, {3, "yeast_mitochondrial"}, {4, "mold_mitochondrial"},
            {5, "invertebrate_mitochondrial"}, {6, "ciliate_nuclear"}, {9, "echinoderm_mitochondrial"}, {10, "euplotid_nuclear"},
            {11, "bacterial"}, {12, "alternative_yeast_nuclear"}, {13, "ascidian_mitochondrial"}, {14, "flatworm_mitochondrial"},
            {15, "blepharisma_macronuclear"}, {16, "chlorophycean_mitochondrial"}, {21, "trematode_mitochondrial"},
            {22, "scenedesmus_mitochondrial"}, {23, "thraustochytrium_mitochondrial"}};
        std::string nt, b1, b2, b3;
        const char* tcag = "TCAG";
        for (int i = 0; i < 64; ++i) {
            b1 += tcag[i / 16];
            b2 += tcag[(i / 4) % 4];
            b3 += tcag[i % 4];
            nt += b1.back();
            nt += b2.back();
            nt += b3.back();
        }
        std::vector<uint64_t> off = {0, 192};
        std::vector<uint8_t> aa(umgap_translate_bound(192, 1, 1)), am(aa.size());
        std::vector<uint64_t> aoff(2);
        check(umgap_translate(0, (const uint8_t*)nt.data(), off.data(), 1, table, 0, 1, aa.data(), aoff.data()));
        check(umgap_translate(0, (const uint8_t*)nt.data(), off.data(), 1, table, 1, 1, am.data(), aoff.data()));
        std::string starts;  // a start codon is one that -m turns into M
        for (int i = 0; i < 64; ++i) starts += (am[i] == 'M' && (aa[i] != 'M' || i == 35)) ? 'M' : '-';
        auto it = table_names.find(table);
        std::string out = std::string(it == table_names.end() ? "table" : it->second) + "=" + std::to_string(table) + "\n";
        out += "AAs    = " + std::string((const char*)aa.data(), 64) + "\n";
        out += "Starts = " + starts + "\n";
        out += "Base1  = " + b1 + "\nBase2  = " + b2 + "\nBase3  = " + b3 + "\n";
        put(stdout, out);
        return 0;
    }
    uint8_t mask = 0;
    for (int k : order) mask |= (uint8_t)(1u << k);
    const int nsel = __builtin_popcount(mask);
    int slot_of[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0, s = 0; i < 6; ++i)
        if (mask >> i & 1) slot_of[i] = s++;
    // an unknown table is an error even on empty input (translate.rs:79)
    {
        const uint64_t z[1] = {0};
        uint64_t ao[1];
        uint8_t dummy[1];
        check(umgap_translate(0, dummy, z, 0, table, meth, 1, dummy, ao));
    }
    // blocks of whole records parsed in place, one call per block, the text of the block formatted into one buffer
    BlockReader br(stdin, 8u << 20);
    UnwrappedBlock B;
    std::vector<uint8_t> aa;
    std::vector<uint64_t> aoff;
    std::vector<char> out;
    const bool named = a.has("append-name");
    const char *p, *end;
    while (br.next(p, end)) {
        B.parse(p, end);
        const size_t n = B.size();
        if (!nsel || !n) continue;
        aa.resize(umgap_translate_bound(B.seq.size(), n, mask));
        aoff.resize(n * nsel + 1);
        check(umgap_translate(0, (const uint8_t*)B.seq.data(), B.off.data(), n, table, meth, mask, aa.data(), aoff.data()));
        size_t need = 0;
        for (size_t i = 0; i < n; ++i) need += order.size() * (B.heads[i].second + 6);
        for (int k : order)
            for (size_t i = 0; i < n; ++i) need += aoff[i * nsel + slot_of[k] + 1] - aoff[i * nsel + slot_of[k]];
        out.resize(need);
        char* w = out.data();
        for (size_t i = 0; i < n; ++i)
            for (int k : order) {  // fasta::Writer, empty separator (fasta.rs:164-180)
                const size_t j = i * nsel + slot_of[k];
                *w++ = '>';
                memcpy(w, B.heads[i].first, B.heads[i].second);
                w += B.heads[i].second;
                if (named) {
                    *w++ = '|';
                    for (const char* q = names[k]; *q; ++q) *w++ = *q;
                }
                *w++ = '\n';
                const size_t len = aoff[j + 1] - aoff[j];
                memcpy(w, aa.data() + aoff[j], len);
                w += len;
                if (len) *w++ = '\n';
            }
        if (w != out.data() && fwrite(out.data(), 1, w - out.data(), stdout) != (size_t)(w - out.data())) fail("failed writing output");
    }
    return 0;
}

// ---- prot2kmer2lca ------------------------------------------------------------------------------
void stream_prot2kmer2lca(FILE* in, FILE* out_f, const umgap_index* idx, bool one_on_one) {
    // blocks of whole records parsed in place, one lookup call per block, the ids printed into one buffer
    BlockReader br(in, 8u << 20);
    UnwrappedBlock B;
    std::vector<uint32_t> taxa;
    std::vector<uint64_t> toff;
    std::vector<uint8_t> kept;
    std::vector<char> out;
    const char *p, *end;
    while (br.next(p, end)) {
        B.parse(p, end);
        const size_t n = B.size();
        if (!n) continue;
        taxa.resize(umgap_kmer_lookup_bound(B.seq.size(), n));
        toff.resize(n + 1);
        kept.resize(n + 1);
        check(umgap_kmer_lookup(idx, (const uint8_t*)B.seq.data(), B.off.data(), n, one_on_one, taxa.data(), toff.data(), kept.data()));
        size_t need = 11 * toff[n];
        for (size_t i = 0; i < n; ++i)
            if (kept[i]) need += B.heads[i].second + 2;
        out.resize(need);
        char* w = out.data();
        for (size_t i = 0; i < n; ++i) {
            if (!kept[i]) continue;  // prot2kmer2lca.rs:172
            *w++ = '>';
            memcpy(w, B.heads[i].first, B.heads[i].second);
            w += B.heads[i].second;
            *w++ = '\n';
            for (uint64_t j = toff[i]; j < toff[i + 1]; ++j) {
                w = put_u32(w, taxa[j]);
                *w++ = '\n';
            }
        }
        if (w != out.data() && fwrite(out.data(), 1, w - out.data(), out_f) != (size_t)(w - out.data())) fail("failed writing output");
        fflush(out_f);
    }
}

int cmd_prot2kmer2lca(int argc, char** argv) {
    Args a = parse(argc, argv, 2, {{'k', "length", true}, {'o', "one-on-one", false}, {'s', "socket", true},
                                   {'m', "in-memory", false}, {'c', "chunksize", true}});
    if (a.pos.size() != 1) fail("The following required arguments were not provided:\n    <fst-file>");
    const int k = (int)parse_usize(a.get("length", "9"));
    (void)parse_usize(a.get("chunksize", "240"));  // accepted; batches are sized for the GPU
    IndexHandle idx;
    check(umgap_index_load_fst(a.pos[0].c_str(), k, 0, 0.0, &idx.p));
    if (!a.has("socket")) {
        stream_prot2kmer2lca(stdin, stdout, idx.p, a.has("one-on-one"));
        return 0;
    }
    // socket server mode (prot2kmer2lca.rs:116-137): one connection at a time, until killed
    const std::string path = a.get("socket", "");
    const int srv = socket(AF_UNIX, SOCK_STREAM, 0);
    if (srv < 0) fail(std::string("socket: ") + strerror(errno));
    sockaddr_un addr{};
    addr.sun_family = AF_UNIX;
    if (path.size() >= sizeof addr.sun_path) fail("socket path too long");
    strcpy(addr.sun_path, path.c_str());
    if (bind(srv, (sockaddr*)&addr, sizeof addr) != 0) fail(std::string("bind: ") + strerror(errno));
    if (listen(srv, 16) != 0) fail(std::string("listen: ") + strerror(errno));
    printf("Socket created, listening for connections.\n");
    fflush(stdout);
    for (;;) {
        const int c = accept(srv, nullptr, nullptr);
        if (c < 0) {
            if (errno == EINTR) continue;
            fail(std::string("accept: ") + strerror(errno));
        }
        printf("Connection accepted.\n");
        fflush(stdout);
        FILE* in = fdopen(c, "r");
        FILE* out = fdopen(dup(c), "w");
        try {
            stream_prot2kmer2lca(in, out, idx.p, a.has("one-on-one"));
            printf("Connection finished succesfully.\n");
        } catch (const Fail& e) {
            printf("Connection died with an error: %s\n", e.what());
        }
        fflush(stdout);
        fclose(out);
        fclose(in);
    }
}

// ---- prot2tryp2lca ------------------------------------------------------------------------------
int cmd_prot2tryp2lca(int argc, char** argv) {
    Args a = parse(argc, argv, 2, {{'o', "one-on-one", false}, {'m', "in-memory", false}, {'c', "chunksize", true},
                                   {'p', "pattern", true}, {'l', "minlen", true}, {'L', "maxlen", true},
                                   {'k', "keep", true}, {'d', "drop", true}});
    if (a.pos.size() != 1) fail("The following required arguments were not provided:\n    <fst-file>");
    if (a.get("pattern", "([KR])([^P])") != "([KR])([^P])")
        fail("only the default cleavage pattern ([KR])([^P]) is implemented on the GPU path");
    const int minlen = (int)parse_usize(a.get("minlen", "5")), maxlen = (int)parse_usize(a.get("maxlen", "50"));
    const std::string keep = a.get("keep", ""), drop = a.get("drop", "");
    IndexHandle idx;
    check(umgap_index_load_fst(a.pos[0].c_str(), 0, 0, 0.0, &idx.p));
    FastaReader rd(stdin, false);
    std::vector<Record> recs;
    Record r;
    bool more = true;
    while (more) {
        recs.clear();
        while (recs.size() < kBatchRecords && (more = rd.next(r))) recs.push_back(r);
        if (recs.empty()) break;
        std::string aa;
        std::vector<uint64_t> off(1, 0);
        for (auto& rec : recs)
            for (auto& line : rec.seq) {
                aa += line;
                off.push_back(aa.size());
            }
        const uint64_t nlines = off.size() - 1;
        std::vector<uint32_t> taxa(umgap_tryp_lookup_bound(aa.size(), nlines));
        std::vector<uint64_t> toff(nlines + 1);
        check(umgap_tryp_lookup(idx.p, (const uint8_t*)aa.data(), off.data(), nlines, minlen, maxlen, keep.c_str(), drop.c_str(),
                                a.has("one-on-one"), taxa.data(), toff.data()));
        std::string out;
        uint64_t l = 0;
        for (auto& rec : recs) {
            out += '>';
            out += rec.header;
            out += '\n';  // the header is always written (prot2tryp2lca.rs:108)
            for (size_t s = 0; s < rec.seq.size(); ++s, ++l)
                for (uint64_t j = toff[l]; j < toff[l + 1]; ++j) {
                    out += std::to_string(taxa[j]);
                    out += '\n';
                }
        }
        put(stdout, out);
    }
    return 0;
}

// ---- seedextend / taxa2agg ---------------------------------------------------------------------
int cmd_seedextend(int argc, char** argv) {
    Args a = parse(argc, argv, 2, {{'s', "min-seed-size", true}, {'g', "max-gap-size", true}, {'r', "ranked", true}, {'p', "penalty", true}});
    const int s = (int)parse_usize(a.get("min-seed-size", "2")), g = (int)parse_usize(a.get("max-gap-size", "0"));
    const int penalty = (int)parse_usize(a.get("penalty", "5"));
    TaxHandle tax;  // -r: only the extended seed with the highest rank score is kept (seedextend.rs:84-90,151-164)
    if (a.has("ranked")) check(umgap_taxonomy_load(a.get("ranked", "").c_str(), 0, &tax.p));
    BlockReader br(stdin);
    IdBatch b;
    std::vector<uint32_t> out_ids;
    std::vector<uint64_t> ooff;
    std::string out;
    auto flush = [&]() {
        if (!b.size()) return;
        out_ids.assign(b.ids.size() + 1, 0);
        ooff.assign(b.size() + 1, 0);
        b.ids.push_back(0);  // never read: keeps data() valid for an all-empty batch
        if (tax.p)
            check(umgap_seedextend_ranked(tax.p, b.ids.data(), b.off.data(), b.size(), s, g, penalty, out_ids.data(), ooff.data()));
        else
            check(umgap_seedextend(0, b.ids.data(), b.off.data(), b.size(), s, g, out_ids.data(), ooff.data()));
        out.clear();
        for (size_t i = 0; i < b.size(); ++i) {  // fasta.rs:164-180: header, then one id per line
            out += '>';
            out.append(b.harena, b.hoff[i], b.hoff[i + 1] - b.hoff[i]);
            out += '\n';
            for (uint64_t j = ooff[i]; j < ooff[i + 1]; ++j) {
                append_u32(out, out_ids[j]);
                out += '\n';
            }
        }
        put(stdout, out);
        b.clear();
    };
    const char *p, *end;
    while (br.next(p, end))
        while (p < end) {
            p = parse_id_records(p, end, b, kBatchRecords);
            if (b.size() >= kBatchRecords) flush();
        }
    flush();
    return 0;
}

int parse_strategy(const std::string& method, const std::string& strategy) {
    const bool tree = method == "tree", rmq = method == "rmq";
    if (!tree && !rmq) fail("Invalid method: " + method);
    int st;
    if (strategy == "lca*") st = UMGAP_AGG_LCA_STAR;
    else if (strategy == "hybrid") st = UMGAP_AGG_HYBRID;
    else if (strategy == "mrtl") st = UMGAP_AGG_MRTL;
    else fail("Invalid strategy: " + strategy);
    if (tree && st == UMGAP_AGG_MRTL)  // taxa2agg.rs:134-138
        fail("Invalid invocation: Tree and MaximumRootToLeafPath cannot be combined");
    if (rmq && st == UMGAP_AGG_HYBRID) {  // taxa2agg.rs:117-124
        fputs("Warning: this is a hybrid between LCA/MRTL, not LCA*/MRTL\n", stderr);
        return UMGAP_AGG_RMQ_HYBRID;
    }
    // -m rmq -a lca*: the fold of rmq/lca.rs:60-90 gives tree/lca.rs:34-40's answer in every order of the record's taxa
    // (oracle/rmq.py restates it with its Euler tour and RMQ; tests/test_oracle_golden.py checks all orders): same kernel
    return st;
}

int cmd_taxa2agg(int argc, char** argv) {
    Args a = parse(argc, argv, 2, {{'s', "scored", false}, {'r', "ranked", false}, {'m', "method", true}, {'a', "aggregate", true},
                                   {'f', "factor", true}, {'l', "lower-bound", true}});
    if (a.pos.size() != 1) fail("The following required arguments were not provided:\n    <taxon-file>");
    const int st = parse_strategy(a.get("method", "tree"), a.get("aggregate", "hybrid"));
    const float factor = parse_f32(a.get("factor", "0.25")), lb = parse_f32(a.get("lower-bound", "0"));
    TaxHandle tax;
    check(umgap_taxonomy_load(a.pos[0].c_str(), 0, &tax.p));
    if (a.has("scored")) {  // -s: every line is "taxon=score" (taxa2agg.rs:141-148) -> umgap_aggregate_scored
        FastaReader rd(stdin, false);
        Record r;
        std::vector<std::string> heads;
        std::vector<uint32_t> ids, res;
        std::vector<float> scores;
        std::vector<uint64_t> off{0};
        std::string out;
        auto flush = [&]() {
            if (heads.empty()) return;
            res.assign(heads.size(), 0);
            ids.push_back(0);
            scores.push_back(0);
            check(umgap_aggregate_scored(tax.p, ids.data(), scores.data(), off.data(), heads.size(), st, factor, lb, a.has("ranked"), res.data()));
            out.clear();
            for (size_t i = 0; i < heads.size(); ++i) {
                out += '>';
                out += heads[i];
                out += '\n';
                append_u32(out, res[i]);
                out += '\n';
            }
            put(stdout, out);
            heads.clear();
            ids.clear();
            scores.clear();
            off.assign(1, 0);
        };
        while (rd.next(r)) {
            for (const std::string& pair : r.seq) {
                const size_t eq = pair.find('=');
                if (eq == std::string::npos || pair.find('=', eq + 1) != std::string::npos) fail("Taxon without score");
                const std::string t = pair.substr(0, eq);
                if (t.empty() || t.find_first_not_of("0123456789") != std::string::npos) fail("invalid digit found in string");
                errno = 0;
                const unsigned long long v = strtoull(t.c_str(), nullptr, 10);
                if (errno || v >= 0xFFFFFFFFull) fail("taxon id out of range: " + t);
                ids.push_back((uint32_t)v);
                scores.push_back(parse_f32(pair.substr(eq + 1)));
            }
            off.push_back(ids.size());
            heads.push_back(r.header);
            if (heads.size() >= kBatchRecords) flush();
        }
        flush();
        return 0;
    }
    BlockReader br(stdin);
    IdBatch b;
    std::vector<uint32_t> res;
    std::vector<float> ones;
    std::string out;
    auto flush = [&]() {
        if (!b.size()) return;
        res.assign(b.size(), 0);
        b.ids.push_back(0);
        if (st == UMGAP_AGG_RMQ_HYBRID) {  // the serial kernel of the scored mode, every score 1.0 (taxa2agg.rs:150-152)
            ones.assign(b.ids.size(), 1.0f);
            check(umgap_aggregate_scored(tax.p, b.ids.data(), ones.data(), b.off.data(), b.size(), st, factor, lb, a.has("ranked"), res.data()));
        } else {
            check(umgap_aggregate(tax.p, b.ids.data(), b.off.data(), b.size(), st, factor, lb, a.has("ranked"), res.data()));
        }
        out.clear();
        for (size_t i = 0; i < b.size(); ++i) {
            out += '>';
            out.append(b.harena, b.hoff[i], b.hoff[i + 1] - b.hoff[i]);
            out += '\n';
            append_u32(out, res[i]);
            out += '\n';
        }
        put(stdout, out);
        b.clear();
    };
    const char *p, *end;
    while (br.next(p, end))
        while (p < end) {
            p = parse_id_records(p, end, b, kBatchRecords);
            if (b.size() >= kBatchRecords) flush();
        }
    flush();
    return 0;
}

// ---- uniq / fastq2fasta (host only) ---------------------------------------------------------------
int cmd_uniq(int argc, char** argv) {
    Args a = parse(argc, argv, 2, {{'s', "separator", true}, {'w', "wrap", false}, {'d', "delimiter", true}});
    const std::string sep = a.get("separator", "\n");
    const bool wrap = a.has("wrap"), has_delim = a.has("delimiter");
    const std::string delim = a.get("delimiter", "");
    if (sep == "\n" && !wrap && !getenv("UMGAP_UNIQ_RECORDS")) {
        // The form every pipeline uses (items stay one per line): the stream is joined in place, block by block.  A group's
        // text is ">header\n", its items joined by "\n", and a final "\n" unless the joined string is empty
        // (fasta.rs:164-180) -- one item that is empty, or none.
        BlockReader br(stdin);
        std::string last, out;
        bool have_last = false, nonempty = false;
        size_t nitems = 0;
        const char *p, *end;
        while (br.next(p, end)) {
            while (p < end) {
                const char* ls;
                size_t ll;
                take_line(p, end, ls, ll);
                const char* hs = ls + 1;
                size_t hl = ll - 1;
                if (has_delim) {  // the header up to the first occurrence of the delimiter (uniq.rs:61-68)
                    const void* m = delim.empty() ? hs : memmem(hs, hl, delim.data(), delim.size());
                    if (m) hl = (const char*)m - hs;
                }
                if (!(have_last && last.size() == hl && memcmp(last.data(), hs, hl) == 0)) {
                    if (have_last && (nitems >= 2 || nonempty)) out += '\n';
                    out += '>';
                    out.append(hs, hl);
                    out += '\n';
                    last.assign(hs, hl);
                    have_last = true;
                    nitems = 0;
                    nonempty = false;
                }
                while (p < end && *p != '>') {
                    take_line(p, end, ls, ll);
                    if (nitems) out += '\n';
                    out.append(ls, ll);
                    ++nitems;
                    nonempty |= ll > 0;
                }
                if (out.size() > (1 << 20)) {
                    put(stdout, out);
                    out.clear();
                }
            }
        }
        if (have_last && (nitems >= 2 || nonempty)) out += '\n';
        put(stdout, out);
        return 0;
    }
    FastaReader rd(stdin, false);
    Record r, last;
    bool have_last = false;
    std::string out;
    while (rd.next(r)) {  // uniq.rs:59-82
        if (has_delim) {
            const size_t p = r.header.find(delim);
            if (p != std::string::npos) r.header.resize(p);
        }
        if (have_last && last.header == r.header) {
            last.seq.insert(last.seq.end(), r.seq.begin(), r.seq.end());
        } else {
            if (have_last) write_record(out, last.header, last.seq, sep, wrap);
            last = r;
            have_last = true;
        }
        if (out.size() > (1 << 20)) {
            put(stdout, out);
            out.clear();
        }
    }
    if (have_last) write_record(out, last.header, last.seq, sep, wrap);
    put(stdout, out);
    return 0;
}

// Lines of a FASTQ stream as views into a refilled buffer (valid until the next call), "\n" / "\r\n" stripped.
class FastqLines {
  public:
    explicit FastqLines(FILE* f) : f_(f), buf_(4u << 20) {}
    bool next(const char*& ls, size_t& ll) {
        for (;;) {
            const char* nl = (const char*)memchr(buf_.data() + pos_, '\n', len_ - pos_);
            if (nl) {
                ls = buf_.data() + pos_;
                ll = nl - ls;
                pos_ = (nl - buf_.data()) + 1;
                if (ll && ls[ll - 1] == '\r') --ll;
                return true;
            }
            if (eof_) {  // the last line has no line end
                if (pos_ == len_) return false;
                ls = buf_.data() + pos_;
                ll = len_ - pos_;
                pos_ = len_;
                return true;
            }
            memmove(buf_.data(), buf_.data() + pos_, len_ - pos_);
            len_ -= pos_;
            pos_ = 0;
            if (len_ == buf_.size()) buf_.resize(buf_.size() * 2);  // one line longer than the buffer
            const size_t n = fread(buf_.data() + len_, 1, buf_.size() - len_, f_);
            if (n == 0) eof_ = true;
            len_ += n;
        }
    }

  private:
    FILE* f_;
    std::vector<char> buf_;
    size_t pos_ = 0, len_ = 0;
    bool eof_ = false;
};

// The next record of a FASTQ stream (src/io/fastq.rs:26-87) appended to `out` as fasta::Writer prints it with an empty
// separator (fasta.rs:164-180): ">header\n", the sequence lines joined, "\n" unless the sequence is empty.  The quality
// lines -- as many as there were sequence lines -- are skipped.  False when the stream is exhausted.
bool fastq_record_to_fasta(FastqLines& src, std::string& out) {
    const char* ls;
    size_t ll;
    if (!src.next(ls, ll)) return false;
    if (ll == 0 || ls[0] != '@') fail("Expected @ at beginning of fastq header.");
    out += '>';
    out.append(ls + 1, ll - 1);
    out += '\n';
    size_t n = 0, seq = 0;
    while (src.next(ls, ll) && !(ll && ls[0] == '+')) {
        out.append(ls, ll);
        seq += ll;
        ++n;
    }
    if (seq) out += '\n';
    for (size_t i = 0; i < n; ++i)
        if (!src.next(ls, ll)) fail("Expected as many quality lines as sequence lines.");
    return true;
}

int cmd_fastq2fasta(int argc, char** argv) {
    Args a = parse(argc, argv, 2, {});
    if (a.pos.empty()) fail("The following required arguments were not provided:\n    <input>...");
    std::vector<FILE*> files;
    std::vector<std::unique_ptr<FastqLines>> rd;
    for (auto& p : a.pos) {
        FILE* f = fopen(p.c_str(), "r");
        if (!f) fail(p + ": " + strerror(errno));
        files.push_back(f);
        rd.emplace_back(new FastqLines(f));
    }
    std::string out;
    for (;;) {  // one record from every file, stop when any is exhausted (fastq2fasta.rs:62-84)
        const size_t row = out.size();
        bool ok = true;
        for (size_t i = 0; i < rd.size() && ok; ++i) ok = fastq_record_to_fasta(*rd[i], out);
        if (!ok) {
            out.resize(row);  // a row with a file short of a record is not written
            break;
        }
        if (out.size() > (1 << 20)) {
            put(stdout, out);
            out.clear();
        }
    }
    put(stdout, out);
    for (FILE* f : files) fclose(f);
    return 0;
}

// ---- classify: the whole preset in one process (extension) ---------------------------------------
// Host side of the fused command.  The stream is cut into blocks of whole uniq groups; K parser threads turn blocks
// into the arrays the library takes (no per-record strings; fasta.rs:38-67 with unwrap: a record is its header line
// and the concatenation of the lines up to the next line that starts with '>'); one thread per GPU classifies parsed
// blocks on its replica of the index (--gpus / UMGAP_DEVICES: index and taxonomy replicated, reads partitioned,
// SURVEY 8(e) mode 1); the writer prints the blocks in input order.
// Page-locked vectors (umgap_host_alloc): the library's copies of the ranges of a batch overlap its kernels only from
// and into such memory.
template <class T>
struct PinnedAlloc {
    using value_type = T;
    PinnedAlloc() = default;
    template <class U>
    PinnedAlloc(const PinnedAlloc<U>&) {}
    T* allocate(size_t n) {
        void* p = umgap_host_alloc(n * sizeof(T));
        if (!p) throw std::bad_alloc();
        return (T*)p;
    }
    void deallocate(T* p, size_t) { umgap_host_free(p); }
    template <class U>
    bool operator==(const PinnedAlloc<U>&) const { return true; }
    template <class U>
    bool operator!=(const PinnedAlloc<U>&) const { return false; }
};
template <class T>
using PinnedVec = std::vector<T, PinnedAlloc<T>>;

namespace classify_cli {

// The nucleotides of a block, in page-locked memory (umgap_host_alloc): the library copies them to the GPU asynchronously.
struct PinnedBytes {
    char* p = nullptr;
    size_t n = 0, cap = 0;
    ~PinnedBytes() { umgap_host_free(p); }
    void clear() { n = 0; }
    size_t size() const { return n; }
    void resize(size_t m) { n = m; }  // shrink only
    void reserve(size_t m) {
        if (m <= cap) return;
        char* q = (char*)umgap_host_alloc(m + m / 8 + 64);
        if (!q) fail(umgap_last_error());
        if (n) memcpy(q, p, n);
        umgap_host_free(p);
        p = q;
        cap = m + m / 8 + 64;
    }
    void append(const char* src, size_t len) {
        if (n + len > cap) reserve(n + len);
        memcpy(p + n, src, len);
        n += len;
    }
};

struct Batch {
    PinnedBytes nt;
    std::string harena;
    PinnedVec<uint64_t> roff, goff;  // page-locked like nt: an asynchronous batch is copied from where it lies
    std::vector<uint64_t> hoff;      // hoff[g] .. hoff[g+1]: header of group g in harena
    void reset() {
        nt.clear();
        harena.clear();
        roff.assign(1, 0);
        goff.clear();
        hoff.assign(1, 0);
    }
    size_t groups() const { return hoff.size() - 1; }
};

struct Job {
    std::vector<char> text;
    const char* view = nullptr;  // the block's bytes: text.data(), or a window of the memory-mapped input file
    size_t len = 0;
    uint64_t seq = 0;
    int stage = 0;  // what the pool threads do with it next: 0 parse, 1 format the classified groups
    off_t file_off = -1;  // >= 0: the block is read from the input file at this offset into `text` by the thread that parses it
    Batch batch;
    PinnedVec<uint32_t> res;
    std::string out;
};

template <class T>
class Queue {  // unbounded FIFO; the number of jobs in flight bounds it
  public:
    void push(T v) {
        {
            std::lock_guard<std::mutex> lk(mu_);
            q_.push_back(v);
        }
        cv_.notify_one();
    }
    bool pop(T& v) {  // false once closed and drained
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return !q_.empty() || closed_; });
        if (q_.empty()) return false;
        v = q_.front();
        q_.erase(q_.begin());
        return true;
    }
    bool try_pop(T& v) {  // false when nothing is queued right now
        std::lock_guard<std::mutex> lk(mu_);
        if (q_.empty()) return false;
        v = q_.front();
        q_.erase(q_.begin());
        return true;
    }
    void close() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            closed_ = true;
        }
        cv_.notify_all();
    }

  private:
    std::mutex mu_;
    std::condition_variable cv_;
    std::vector<T> q_;
    bool closed_ = false;
};

// Start of the record that ends at `end` (the position of a '>' that follows a '\n', or the end of the data).
inline size_t record_start(const char* buf, size_t end) {
    for (size_t q = end; q > 0;) {
        const void* m = memrchr(buf, '>', q);
        if (!m) break;
        const size_t at = (const char*)m - buf;
        if (at == 0 || buf[at - 1] == '\n') return at;
        q = at;
    }
    return 0;
}
// Header (truncated at the uniq delimiter) and sequence length of the record in [rs, re).
inline void record_view(const char* buf, size_t rs, size_t re, const std::string& delim, const char*& hs, size_t& hl, size_t& seq_len) {
    const char* e = (const char*)memchr(buf + rs, '\n', re - rs);
    const char* hend = e ? e : buf + re;
    hs = buf + rs + 1;
    hl = hend - hs;
    if (e && hl && hs[hl - 1] == '\r') --hl;
    if (!delim.empty()) {
        const void* m = memmem(hs, hl, delim.data(), delim.size());
        if (m) hl = (const char*)m - hs;
    }
    seq_len = 0;
    for (const char* p = e ? e + 1 : buf + re; p < buf + re; ++p) seq_len += (*p != '\n' && *p != '\r');
}
// Where to cut [0, len): the start of a complete record that takes part in uniq's grouping (long enough for a frame,
// prot2kmer2lca.rs:172) and whose group differs from that of the last such record before it -- so that no uniq
// group (uniq.rs:56-84) spans two blocks.  0: no such place in the buffer yet.
size_t find_cut(const char* buf, size_t len, const std::string& delim, size_t span) {
    size_t r_end = record_start(buf, len);  // the last record may be incomplete: it stays behind the cut
    while (r_end > 0) {
        const size_t at = record_start(buf, r_end);
        const char *h2, *h1;
        size_t l2, l1, n2, n1;
        record_view(buf, at, r_end, delim, h2, l2, n2);
        if (n2 >= span) {
            bool joined = false;
            for (size_t pe = at; pe > 0;) {  // the last grouped record before the candidate
                const size_t ps = record_start(buf, pe);
                record_view(buf, ps, pe, delim, h1, l1, n1);
                if (n1 >= span) {
                    joined = l1 == l2 && memcmp(h1, h2, l1) == 0;
                    break;
                }
                pe = ps;
            }
            if (!joined && at > 0) return at;
        }
        r_end = at;
    }
    return 0;
}

void parse_block(const char* p, const char* end, const std::string& delim, size_t span, Batch* B) {
    B->reset();
    B->nt.reserve(end - p);
    if (B->roff.capacity() < (size_t)(end - p) / 48 + 16) {  // page-locked vectors grow dearly: sized once for blocks of short reads
        B->roff.reserve((size_t)(end - p) / 48 + 16);
        B->goff.reserve((size_t)(end - p) / 48 + 16);
    }
    while (p < end) {
        const char* e = (const char*)memchr(p, '\n', end - p);  // header line
        const char* hend = e ? e : end;
        const char* hs = p + 1;
        size_t hl = hend - hs;
        if (e && hl && hs[hl - 1] == '\r') --hl;
        p = e ? e + 1 : end;
        const size_t nt0 = B->nt.size();
        while (p < end && *p != '>') {  // sequence lines
            const char* le = (const char*)memchr(p, '\n', end - p);
            const char* lend = le ? le : end;
            size_t ll = lend - p;
            if (le && ll && p[ll - 1] == '\r') --ll;
            B->nt.append(p, ll);
            p = le ? le + 1 : end;
        }
        // reads too short for any frame emit no record at all and so do not take part in uniq's grouping
        // (prot2kmer2lca.rs:172 drops them before uniq sees them)
        if (B->nt.size() - nt0 < span) {
            B->nt.resize(nt0);
            continue;
        }
        if (!delim.empty()) {  // uniq -d: the header up to the first occurrence of the delimiter (uniq.rs:61-68)
            const void* m = memmem(hs, hl, delim.data(), delim.size());
            if (m) hl = (const char*)m - hs;
        }
        const size_t ng = B->groups();
        const bool same = ng && B->hoff[ng] - B->hoff[ng - 1] == hl && memcmp(B->harena.data() + B->hoff[ng - 1], hs, hl) == 0;
        if (!same) {
            B->goff.push_back(B->roff.size() - 1);
            B->harena.append(hs, hl);
            B->hoff.push_back(B->harena.size());
        }
        B->roff.push_back(B->nt.size());
    }
    B->goff.push_back(B->roff.size() - 1);
}

// Peptide records (what the gene predictor prints): every physical line is an item of its record
// (prot2tryp2lca.rs:105-118), records with the same header up to the delimiter form a group; nt = the residues,
// roff = line offsets, goff = lines before each group.  A record without lines still names its group.
void parse_peptide_block(const char* p, const char* end, const std::string& delim, Batch* B) {
    B->reset();
    B->nt.reserve(end - p);
    if (B->roff.capacity() < (size_t)(end - p) / 16 + 16) {
        B->roff.reserve((size_t)(end - p) / 16 + 16);
        B->goff.reserve((size_t)(end - p) / 16 + 16);
    }
    while (p < end) {
        const char* e = (const char*)memchr(p, '\n', end - p);  // header line
        const char* hend = e ? e : end;
        const char* hs = p + 1;
        size_t hl = hend - hs;
        if (e && hl && hs[hl - 1] == '\r') --hl;
        p = e ? e + 1 : end;
        if (!delim.empty()) {
            const void* m = memmem(hs, hl, delim.data(), delim.size());
            if (m) hl = (const char*)m - hs;
        }
        const size_t ng = B->groups();
        const bool same = ng && B->hoff[ng] - B->hoff[ng - 1] == hl && memcmp(B->harena.data() + B->hoff[ng - 1], hs, hl) == 0;
        if (!same) {
            B->goff.push_back(B->roff.size() - 1);
            B->harena.append(hs, hl);
            B->hoff.push_back(B->harena.size());
        }
        while (p < end && *p != '>') {
            const char* le = (const char*)memchr(p, '\n', end - p);
            const char* lend = le ? le : end;
            size_t ll = lend - p;
            if (le && ll && p[ll - 1] == '\r') --ll;
            B->nt.append(p, ll);
            B->roff.push_back(B->nt.size());
            p = le ? le + 1 : end;
        }
    }
    B->goff.push_back(B->roff.size() - 1);
}

}  // namespace classify_cli

// What the block pipeline below runs: reads (translate | prot2kmer2lca | seedextend | uniq | taxa2agg) or peptide
// records (prot2tryp2lca | uniq | taxa2agg).
struct BlockPipeline {
    bool peptides = false;
    umgap_pipeline_opts o;
    umgap_tryp_opts to;
    std::string delim;
    int k = 9;
};

// Blocks of whole uniq groups cut from stdin, parsed on K threads straight into the library's arrays, classified on
// one thread per GPU, formatted on the pool threads, printed in input order.
int run_block_pipeline(const Args& a, const BlockPipeline& cfg) {
    using namespace classify_cli;
    const umgap_pipeline_opts& o = cfg.o;
    const std::string& delim = cfg.delim;
    const int k = cfg.k;
    const bool peptides = cfg.peptides;
    // devices: UMGAP_DEVICES=0,2,3 or --gpus N (devices 0..N-1); one by default
    std::vector<int> devices;
    if (const char* e = getenv("UMGAP_DEVICES")) {
        for (const char* p = e; *p;) {
            char* q = nullptr;
            devices.push_back((int)strtol(p, &q, 10));
            if (q == p) fail("UMGAP_DEVICES: expected a comma-separated list of device numbers");
            p = *q == ',' ? q + 1 : q;
        }
    }
    if (a.has("gpus")) {
        const int n = (int)parse_usize(a.get("gpus", "1"));
        if (n < 1) fail("--gpus needs a positive number");
        if (devices.empty()) for (int i = 0; i < n; ++i) devices.push_back(i);
        devices.resize(std::min<size_t>(devices.size(), n));
    }
    if (devices.empty()) devices.push_back(0);
    const size_t G = devices.size();
    std::vector<IndexHandle> idx(G);
    std::vector<TaxHandle> tax(G);
    check(umgap_index_load_fst(a.pos[0].c_str(), k, devices[0], 0.0, &idx[0].p));
    check(umgap_taxonomy_load(a.pos[1].c_str(), devices[0], &tax[0].p));
    for (size_t g = 1; g < G; ++g) {
        if (peptides) check(umgap_index_load_fst(a.pos[0].c_str(), 0, devices[g], 0.0, &idx[g].p));  // a peptide table is streamed again
        else check(umgap_index_replicate(idx[0].p, devices[g], &idx[g].p));
        check(umgap_taxonomy_replicate(tax[0].p, devices[g], &tax[g].p));
    }
    const auto t_loaded = std::chrono::steady_clock::now();
    std::atomic<uint64_t> n_reads{0}, n_bytes{0};
    // the steady state's own clock: from the moment every job's buffers have been through one block (page-locked
    // allocations and first-touch page faults behind it) -- reads counted as they are classified
    std::chrono::steady_clock::time_point t_warm{};
    uint64_t warm_reads = 0, warm_blocks = 0;
    // block size in bytes (UMGAP_CLI_BLOCK overrides it, for tests of the block seams)
    // 4 MB: a block's text and nucleotides stay in its parser thread's cache (8 MB blocks ran at 60-160 M reads/s from one run
    // to the next, 32 MB blocks at 47-69 M, 4 MB blocks at 150-185 M: profiles/r02_cli_probe*.log)
    const size_t block = getenv("UMGAP_CLI_BLOCK") ? std::max<size_t>(16, strtoull(getenv("UMGAP_CLI_BLOCK"), nullptr, 10)) : (size_t)4 << 20;
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const size_t P = a.has("parser-threads") ? std::max<size_t>(1, parse_usize(a.get("parser-threads", "1"))) : std::min<size_t>(std::max(2u, hw - 2), 12);
    // blocks in flight per GPU (UMGAP_CLI_DEPTH), and how the parser threads get at a mapped input file's bytes
    // (UMGAP_CLI_READ): populate = madvise(MADV_POPULATE_READ) on the block, then parse in place (a fault per page, twelve
    // threads on one address space: 103 M reads/s against 153 M); pread = into the job's own buffer (80 M); mmap = as it
    // comes.  Measured in profiles/README.md section 8.
    const size_t depth = getenv("UMGAP_CLI_DEPTH") ? std::max<size_t>(1, std::min<size_t>(8, strtoull(getenv("UMGAP_CLI_DEPTH"), nullptr, 10))) : 2;
    const std::string read_mode = getenv("UMGAP_CLI_READ") ? getenv("UMGAP_CLI_READ") : "populate";
#ifdef MADV_POPULATE_READ
    const bool populate = read_mode == "populate";
#endif
    const size_t njobs = P + P / 2 + (depth + 1) * G + 2;
    const size_t span = peptides ? 0 : 3 * (size_t)k;  // every peptide record takes part in uniq's grouping

    std::vector<std::unique_ptr<Job>> jobs(njobs);
    Queue<Job*> free_q, parse_q, classify_q;
    for (auto& j : jobs) {
        j.reset(new Job());
        free_q.push(j.get());
    }
    std::mutex mu;  // guards done / error
    std::condition_variable done_cv;
    std::map<uint64_t, Job*> done;
    std::string error;
    bool reader_finished = false;
    uint64_t total_blocks = 0;
    auto set_error = [&](const std::string& m) {
        {
            std::lock_guard<std::mutex> lk(mu);
            if (error.empty()) error = m;
        }
        done_cv.notify_all();
        free_q.close();
        parse_q.close();
        classify_q.close();
    };

    std::vector<std::thread> parsers, classifiers;
    // pool threads: parse a block, or format the groups of a classified one (both are text work; the classifier
    // threads only drive their GPU)
    auto format_job = [&](Job* j) {
        Batch& B = j->batch;
        const size_t ng = B.groups();
        j->out.clear();
        j->out.reserve(B.harena.size() + 12 * ng);
        for (size_t x = 0; x < ng; ++x) {
            if (j->res[x] == UMGAP_ABSENT && !peptides) continue;
            j->out += '>';
            j->out.append(B.harena, B.hoff[x], B.hoff[x + 1] - B.hoff[x]);
            j->out += '\n';
            append_u32(j->out, j->res[x] == UMGAP_ABSENT ? 1u : j->res[x]);  // a peptide record without lines aggregates to the literal 1
            j->out += '\n';
        }
        {
            std::lock_guard<std::mutex> lk(mu);
            done[j->seq] = j;
        }
        done_cv.notify_all();
    };
    for (size_t t = 0; t < P; ++t)
        parsers.emplace_back([&] {
            Job* j;
            while (parse_q.pop(j)) {
                try {
                    if (j->stage == 0) {
                        if (j->file_off >= 0) {  // the block's bytes, read by the thread that parses them
                            if (j->text.size() < j->len) j->text.resize(j->len);
                            for (size_t have = 0; have < j->len;) {
                                const ssize_t n = pread(0, j->text.data() + have, j->len - have, j->file_off + (off_t)have);
                                if (n < 0 && errno == EINTR) continue;
                                if (n <= 0) fail("failed reading input");
                                have += (size_t)n;
                            }
                            j->view = j->text.data();
                        }
#ifdef MADV_POPULATE_READ
                        else if (populate) {  // a mapped block: its pages in one call instead of a fault per page
                            const uintptr_t lo = (uintptr_t)j->view & ~(uintptr_t)4095;
                            (void)madvise((void*)lo, (uintptr_t)j->view + j->len - lo, MADV_POPULATE_READ);
                        }
#endif
                        if (peptides) parse_peptide_block(j->view, j->view + j->len, delim, &j->batch);
                        else parse_block(j->view, j->view + j->len, delim, span, &j->batch);
                        classify_q.push(j);
                    } else {
                        format_job(j);
                    }
                } catch (const std::exception& e) {
                    set_error(e.what());
                    break;
                }
            }
        });
    // classifier threads: one per GPU, `depth` blocks in flight (the next block's upload and first kernels run in the
    // tail of the one before it: umgap_classify_reads_async)
    for (size_t g = 0; g < G; ++g)
        classifiers.emplace_back([&, g] {
            std::vector<std::pair<Job*, umgap_pending*>> fl;
            auto finish = [&]() {
                Job* j = fl.front().first;
                umgap_pending* t = fl.front().second;
                fl.erase(fl.begin());
                if (t) check(umgap_pending_wait(t));
                n_reads += j->batch.roff.size() - 1;
                n_bytes += j->len;
                j->stage = 1;
                parse_q.push(j);
            };
            try {
                for (;;) {
                    Job* j = nullptr;
                    const bool got = fl.size() >= depth ? false : fl.empty() ? classify_q.pop(j) : classify_q.try_pop(j);
                    if (!got) {
                        if (fl.empty()) break;  // closed and drained
                        finish();
                        continue;
                    }
                    Batch& B = j->batch;
                    const size_t ng = B.groups();
                    j->res.resize(ng);
                    umgap_pending* t = nullptr;
                    if (ng && peptides)  // ranges of the batch overlap inside the call
                        check(umgap_classify_peptides(idx[g].p, tax[g].p, &cfg.to, (const uint8_t*)B.nt.p, B.roff.data(), B.roff.size() - 1,
                                                      B.goff.data(), ng, j->res.data()));
                    else if (ng)
                        check(umgap_classify_reads_async(idx[g].p, tax[g].p, &o, (const uint8_t*)B.nt.p, B.roff.data(), B.roff.size() - 1,
                                                         B.goff.data(), ng, j->res.data(), &t));
                    fl.emplace_back(j, t);
                }
            } catch (const std::exception& e) {
                for (auto& f : fl)  // every ticket is waited for before its index goes
                    if (f.second) (void)umgap_pending_wait(f.second);
                set_error(e.what());
            }
        });
    std::thread writer([&] {
        uint64_t next = 0;
        for (;;) {
            Job* j = nullptr;
            {
                std::unique_lock<std::mutex> lk(mu);
                done_cv.wait(lk, [&] { return !error.empty() || done.count(next) || (reader_finished && next >= total_blocks); });
                if (!error.empty()) return;
                if (!done.count(next)) {  // every block is written: the pool and the classifiers can go
                    lk.unlock();
                    parse_q.close();
                    classify_q.close();
                    return;
                }
                j = done[next];
                done.erase(next);
            }
            try {
                put(stdout, j->out);
            } catch (const std::exception& e) {
                set_error(e.what());
                return;
            }
            ++next;
            if (next == 2 * njobs) {
                t_warm = std::chrono::steady_clock::now();
                warm_reads = n_reads.load();
                warm_blocks = next;
            }
            free_q.push(j);
        }
    });

    // reader: this thread.  A regular file on stdin is mapped and cut in place (the parser threads read the file's pages
    // themselves); anything else (a pipe) is read block by block into the jobs' buffers.
    try {
        uint64_t seq = 0;
        struct stat sb;
        const char* map = nullptr;
        size_t map_len = 0;
        off_t map_at = 0;
        if (fstat(0, &sb) == 0 && S_ISREG(sb.st_mode) && sb.st_size > 0 && !getenv("UMGAP_CLI_NO_MMAP")) {
            const off_t at = lseek(0, 0, SEEK_CUR);
            void* m = mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, 0, 0);
            if (m != MAP_FAILED && at >= 0 && at < sb.st_size) {
                map = (const char*)m + at;
                map_len = (size_t)(sb.st_size - at);
                map_at = at;
            }
        }
        if (map) {
            if (map[0] != '>') fail("Expected > at beginning of fasta header.");  // fasta.rs:44-49
            size_t pos = 0;
            while (pos < map_len) {
                size_t want = block, cut = 0;
                for (;;) {
                    if (pos + want >= map_len) {
                        cut = map_len - pos;
                        break;
                    }
                    cut = find_cut(map + pos, want, delim, span);
                    if (cut) break;
                    want *= 2;  // one group larger than the block
                }
                Job* j;
                if (!free_q.pop(j)) break;
                j->stage = 0;
                j->view = map + pos;
                j->file_off = read_mode == "pread" ? map_at + (off_t)pos : (off_t)-1;
                j->len = cut;
                j->seq = seq++;
                parse_q.push(j);
                pos += cut;
            }
        } else {
#ifdef F_SETPIPE_SZ
            (void)fcntl(0, F_SETPIPE_SZ, 1 << 20);  // a larger pipe buffer, where stdin is a pipe and the kernel allows it
#endif
            std::vector<char> carry;  // what the previous block left behind its cut
            bool eof = false, first = true;
            while (!eof || !carry.empty()) {
                Job* j;
                if (!free_q.pop(j)) break;  // closed: an error elsewhere
                j->stage = 0;
                j->file_off = -1;
                if (j->text.size() < block + carry.size()) j->text.resize(block + carry.size());
                memcpy(j->text.data(), carry.data(), carry.size());
                size_t have = carry.size();
                carry.clear();
                size_t cut = 0;
                for (;;) {
                    while (!eof && have < j->text.size()) {
                        const ssize_t n = read(0, j->text.data() + have, j->text.size() - have);
                        if (n < 0) {
                            if (errno == EINTR) continue;
                            fail("failed reading input");
                        }
                        if (n == 0) eof = true;
                        have += (size_t)n;
                    }
                    if (first && have) {
                        if (j->text[0] != '>') fail("Expected > at beginning of fasta header.");  // fasta.rs:44-49
                        first = false;
                    }
                    if (eof) {
                        cut = have;
                        break;
                    }
                    cut = find_cut(j->text.data(), have, delim, span);
                    if (cut) break;
                    j->text.resize(j->text.size() * 2);  // one group larger than the block: read on
                }
                carry.assign(j->text.data() + cut, j->text.data() + have);
                j->view = j->text.data();
                j->len = cut;
                if (cut == 0) {
                    free_q.push(j);
                    continue;
                }
                j->seq = seq++;
                parse_q.push(j);
            }
        }
        {
            std::lock_guard<std::mutex> lk(mu);
            total_blocks = seq;
            reader_finished = true;
        }
        done_cv.notify_all();
    } catch (const std::exception& e) {
        set_error(e.what());
    }
    for (auto& t : parsers) t.join();
    for (auto& t : classifiers) t.join();
    writer.join();
    if (!error.empty()) fail(error);
    if (getenv("UMGAP_CLI_VERBOSE")) {  // the stream's own rate, start-up (CUDA context, index and taxonomy load) left out
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_loaded).count();
        fprintf(stderr, "umgap classify: %llu reads, %llu bytes of FASTA in %.3f s after the index was loaded: %.2f M reads/s, %.2f GB/s "
                "(%zu parser threads, %zu GPU(s))\n", (unsigned long long)n_reads.load(), (unsigned long long)n_bytes.load(), dt,
                n_reads.load() / dt / 1e6, n_bytes.load() / dt / 1e9, P, G);
        if (warm_blocks) {
            const double ds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_warm).count();
            fprintf(stderr, "umgap classify steady state: %.2f M reads/s over the %.3f s after the first %llu blocks\n",
                    (n_reads.load() - warm_reads) / ds / 1e6, ds, (unsigned long long)warm_blocks);
        }
    }
    return 0;
}

int cmd_classify(int argc, char** argv) {
    using namespace classify_cli;
    Args a = parse(argc, argv, 2, {{'t', "table", true}, {'M', "methionine", false}, {'k', "length", true}, {'O', "omit-misses", false},
                                   {'S', "no-seedextend", false}, {'s', "min-seed-size", true}, {'g', "max-gap-size", true},
                                   {'d', "delimiter", true}, {'r', "ranked", false}, {'m', "method", true}, {'a', "aggregate", true},
                                   {'f', "factor", true}, {'l', "lower-bound", true}, {'G', "gpus", true}, {'P', "parser-threads", true}});
    if (a.pos.size() != 2) fail("usage: umgap classify [flags] [--gpus N] <fst-file> <taxon-file> < reads.fa");
    umgap_pipeline_opts o;
    umgap_pipeline_opts_default(&o);
    o.table = (int)parse_usize(a.get("table", "1"));
    o.methionine = a.has("methionine");
    o.one_on_one = !a.has("omit-misses");
    o.seedextend = !a.has("no-seedextend");
    o.min_seed_size = (int)parse_usize(a.get("min-seed-size", "2"));
    o.max_gap_size = (int)parse_usize(a.get("max-gap-size", "0"));
    o.strategy = parse_strategy(a.get("method", a.get("aggregate", "hybrid") == "mrtl" ? "rmq" : "tree"), a.get("aggregate", "hybrid"));
    o.factor = parse_f32(a.get("factor", "0.25"));
    o.lower_bound = parse_f32(a.get("lower-bound", "0"));
    o.ranked_only = a.has("ranked");
    const std::string delim = a.get("delimiter", "/");  // the presets join the mates with `uniq -d /`
    const int k = (int)parse_usize(a.get("length", "9"));
    BlockPipeline cfg;
    cfg.o = o;
    cfg.delim = delim;
    cfg.k = k;
    return run_block_pipeline(a, cfg);
}

// ---- classify-peptides: the tryptic presets in one process (extension) -----------------------------
// prot2tryp2lca | uniq -d / | taxa2agg (scripts/umgap-analyse.sh:291-300) behind the gene predictor: peptide records
// on stdin, `>header\n<taxon>\n` per group of records on stdout.
int cmd_classify_peptides(int argc, char** argv) {
    Args a = parse(argc, argv, 2, {{'l', "minlen", true}, {'L', "maxlen", true}, {'k', "keep", true}, {'d', "drop", true},
                                   {'D', "delimiter", true}, {'r', "ranked", false}, {'m', "method", true}, {'a', "aggregate", true},
                                   {'f', "factor", true}, {'b', "lower-bound", true}, {'G', "gpus", true}, {'P', "parser-threads", true}});
    if (a.pos.size() != 2) fail("usage: umgap classify-peptides [flags] [--gpus N] <tryptic-fst-file> <taxon-file> < peptides.fa");
    BlockPipeline cfg;
    cfg.peptides = true;
    umgap_pipeline_opts_default(&cfg.o);
    umgap_tryp_opts_default(&cfg.to);
    cfg.to.minlen = (int)parse_usize(a.get("minlen", "5"));
    cfg.to.maxlen = (int)parse_usize(a.get("maxlen", "50"));
    const std::string keep = a.get("keep", ""), drop = a.get("drop", "");
    cfg.to.keep = keep.c_str();
    cfg.to.drop = drop.c_str();
    cfg.to.strategy = parse_strategy(a.get("method", a.get("aggregate", "mrtl") == "mrtl" ? "rmq" : "tree"), a.get("aggregate", "mrtl"));
    cfg.to.factor = parse_f32(a.get("factor", "0.25"));
    cfg.to.lower_bound = parse_f32(a.get("lower-bound", "0"));
    cfg.to.ranked_only = a.has("ranked");
    cfg.delim = a.get("delimiter", "/");
    cfg.k = 0;
    return run_block_pipeline(a, cfg);
}

void usage(FILE* f) {
    fputs("umgap 1.1.1 (umgap-b200: GPU implementation of the per-read classification path)\n\n"
          "USAGE:\n    umgap <SUBCOMMAND>\n\nFLAGS:\n    -h, --help       Prints help information\n    -V, --version    Prints version information\n\n"
          "SUBCOMMANDS:\n    translate        Translates DNA on stdin into amino acid sequences (six frames)\n"
          "    prot2kmer2lca    Maps all k-mers of peptides to taxon ids through an fst index\n"
          "    prot2tryp2lca    Digests peptides tryptically and maps them to taxon ids\n"
          "    seedextend       Selects promising regions in sequences of taxon ids\n"
          "    uniq             Joins consecutive FASTA records with the same header\n"
          "    taxa2agg         Aggregates taxon ids per record (lca*, hybrid, mrtl)\n"
          "    fastq2fasta      Interleaves FASTQ files into FASTA\n"
          "    buildindex       Writes an fst index from sorted `key<TAB>taxon id` lines\n"
          "    printindex       Prints the key/value pairs of an fst index\n"
          "    snaptaxon        Snaps taxon ids to a rank or to listed taxa\n"
          "    taxa2freq        Counts taxon ids per ranked ancestor (CSV)\n"
          "    bestof           Picks the best record of every group of frames\n"
          "    classify         translate | prot2kmer2lca | seedextend | uniq | taxa2agg in one process [--gpus N]\n"
          "    classify-peptides  prot2tryp2lca | uniq | taxa2agg in one process [--gpus N]\n", f);
}

// Lines of a stream without their line ends, as (pointer, length) views that stay valid until the next call.
class BlockLines {
  public:
    explicit BlockLines(FILE* f) : src_(f) {}
    bool next(const char*& ls, size_t& ll) {
        if (!src_.next(line_)) return false;
        ls = line_.data();
        ll = line_.size();
        return true;
    }

  private:
    LineSource src_;
    std::string line_;
};

// ---- buildindex / printindex (host only) ----------------------------------------------------------------------------
// buildindex.rs:32-48: TSV `key \t taxon id` on stdin, ordered by key, to an fst Map on stdout.
int cmd_buildindex(int argc, char** argv) {
    Args a = parse(argc, argv, 2, {});
    if (!a.pos.empty()) fail("Found argument '" + a.pos[0] + "' which wasn't expected, or isn't valid in this context");
    umgap_fst_writer* w = nullptr;
    check(umgap_fst_writer_open(nullptr, &w));
    try {
        BlockLines lines(stdin);
        const char* ls;
        size_t ll;
        while (lines.next(ls, ll)) {
            const char* tab = (const char*)memchr(ls, '\t', ll);
            if (!tab || memchr(tab + 1, '\t', ll - (tab + 1 - ls))) fail("CSV deserialize error: expected two tab-separated fields (key, taxon id)");
            const uint64_t v = parse_usize(std::string(tab + 1, ls + ll - (tab + 1)));
            check(umgap_fst_writer_insert(w, (const uint8_t*)ls, (size_t)(tab - ls), v));
        }
    } catch (...) {
        umgap_fst_writer_abort(w);
        throw;
    }
    check(umgap_fst_writer_finish(w));
    return 0;
}

// printindex.rs:38-51: every key of the index with its value, TSV.
int cmd_printindex(int argc, char** argv) {
    Args a = parse(argc, argv, 2, {});
    if (a.pos.size() != 1) fail("The following required arguments were not provided:\n    <fst-file>");
    std::string out;
    struct Ctx {
        std::string* out;
    } ctx{&out};
    check(umgap_fst_stream(a.pos[0].c_str(),
                           [](const uint8_t* key, size_t len, uint64_t value, void* user) -> int {
                               std::string& o = *((Ctx*)user)->out;
                               // csv::Writer quotes a field that holds the delimiter, a quote or a line break
                               bool quote = len == 0;
                               for (size_t i = 0; i < len; ++i) quote |= key[i] == '\t' || key[i] == '"' || key[i] == '\n' || key[i] == '\r';
                               if (quote) {
                                   o += '"';
                                   for (size_t i = 0; i < len; ++i) {
                                       if (key[i] == '"') o += '"';
                                       o += (char)key[i];
                                   }
                                   o += '"';
                               } else {
                                   o.append((const char*)key, len);
                               }
                               o += '\t';
                               o += std::to_string(value);
                               o += '\n';
                               if (o.size() > (1u << 20)) {
                                   if (fwrite(o.data(), 1, o.size(), stdout) != o.size()) return 1;
                                   o.clear();
                               }
                               return 0;
                           },
                           &ctx, nullptr));
    put(stdout, out);
    return 0;
}

// ---- reporting commands on the host: snaptaxon, taxa2freq, bestof ---------------------------------------------------
// The taxonomy as these commands use it (taxon.rs:89-163, 224-301): the TSV, the children map, and filter_ancestors.
struct HostTaxonomy {
    struct Taxon {
        uint64_t id, parent;
        int rank;
        bool valid;
        std::string name;
    };
    std::vector<Taxon> taxa;
    std::vector<int64_t> row_of;  // by id; -1 = no such taxon
    uint64_t root = 0, max_id = 0;

    static int rank_index(const std::string& s) {
        static const char* names[] = {"no rank", "superkingdom", "domain", "realm", "kingdom", "subkingdom", "superphylum", "phylum",
                                      "subphylum", "superclass", "class", "subclass", "infraclass", "superorder", "order", "suborder",
                                      "infraorder", "parvorder", "superfamily", "family", "subfamily", "tribe", "subtribe", "genus",
                                      "subgenus", "species group", "species subgroup", "species", "subspecies", "varietas", "forma", "strain"};
        for (int i = 0; i < 32; ++i)
            if (s == names[i]) return i;
        return -1;
    }
    explicit HostTaxonomy(const std::string& path) {
        FILE* f = fopen(path.c_str(), "rb");
        if (!f) fail("Failed opening taxon file.");
        BlockLines lines(f);
        const char* ls;
        size_t ll;
        while (lines.next(ls, ll)) {
            std::string line(ls, ll);
            while (!line.empty() && isspace((unsigned char)line.back())) line.pop_back();  // trim_end (taxon.rs:89)
            std::vector<std::string> col;
            size_t p = 0;
            for (;;) {
                const size_t t = line.find('\t', p);
                col.push_back(line.substr(p, t == std::string::npos ? std::string::npos : t - p));
                if (t == std::string::npos) break;
                p = t + 1;
            }
            // trim_end also eats a trailing "\x00"?  No: NUL is not whitespace; the valid byte survives it.
            if (col.size() != 5) {
                fclose(f);
                fail("Taxon requires five fields");
            }
            Taxon t;
            t.id = parse_usize(col[0]);
            t.name = col[1];
            t.rank = rank_index(col[2]);
            if (t.rank < 0) {
                fclose(f);
                fail("Matching variant not found");
            }
            t.parent = parse_usize(col[3]);
            if (col[4] == "\x01") t.valid = true;
            else if (col[4] == std::string(1, '\0')) t.valid = false;
            else {
                fclose(f);
                fail("Couldn't parse the valid byte");
            }
            taxa.push_back(t);
        }
        fclose(f);
        if (taxa.empty()) fail("There's no root!");
        for (const Taxon& t : taxa) max_id = std::max(max_id, t.id);
        row_of.assign(max_id + 1, -1);
        std::vector<bool> is_child(max_id + 1, false);
        for (size_t i = 0; i < taxa.size(); ++i) {
            row_of[taxa[i].id] = (int64_t)i;
            if (taxa[i].id != taxa[i].parent) is_child[taxa[i].id] = true;
        }
        // TaxonTree::new (taxon.rs:224-247): the ids never listed with another parent; exactly one may remain
        int nroots = 0;
        std::vector<bool> seen(max_id + 1, false);
        for (const Taxon& t : taxa)
            if (!is_child[t.id] && !seen[t.id]) {
                seen[t.id] = true;
                if (nroots++ == 0) root = t.id;
            }
        if (nroots > 1) fail("More than one root!");
        if (nroots == 0) fail("There's no root!");
    }
    // TaxonTree::filter_ancestors (taxon.rs:251-286): by id, the nearest ancestor-or-self that passes; the walk starts at
    // the root with Some(root); ids the walk does not reach stay None (-1).
    template <class F>
    std::vector<int64_t> filter_ancestors(F pass) const {
        std::vector<std::vector<uint64_t>> children(max_id + 1);
        for (const Taxon& t : taxa)
            if (t.id != t.parent && t.parent <= max_id) children[t.parent].push_back(t.id);
        std::vector<int64_t> out(max_id + 1, -1);
        std::vector<std::pair<uint64_t, int64_t>> stack{{root, (int64_t)root}};
        std::vector<bool> done(max_id + 1, false);
        while (!stack.empty()) {
            auto [cur, anc] = stack.back();
            stack.pop_back();
            if (done[cur]) continue;
            done[cur] = true;
            const int64_t mine = pass(cur) ? (int64_t)cur : anc;
            out[cur] = mine;
            for (uint64_t c : children[cur]) stack.push_back({c, mine});
        }
        return out;
    }
};

// snaptaxon.rs:66-108
int cmd_snaptaxon(int argc, char** argv) {
    Args a = parse(argc, argv, 2, {{'r', "rank", true}, {'t', "taxons", true}, {'i', "invalid", false}});
    // `-t 1239 2`: structopt's Vec option takes every following value, so bare numbers after -t are taxa as well
    std::vector<uint64_t> wanted;
    if (a.has("taxons"))
        for (const std::string& v : a.opt["taxons"]) wanted.push_back(parse_usize(v));
    std::string taxon_file;
    for (const std::string& p : a.pos) {
        const bool numeric = !p.empty() && p.find_first_not_of("0123456789") == std::string::npos;
        if (numeric && a.has("taxons") && !taxon_file.empty()) wanted.push_back(parse_usize(p));
        else if (taxon_file.empty()) taxon_file = p;
        else fail("Found argument '" + p + "' which wasn't expected, or isn't valid in this context");
    }
    if (taxon_file.empty()) fail("The following required arguments were not provided:\n    <taxon-file>");
    int rank = -1;
    if (a.has("rank")) {
        rank = HostTaxonomy::rank_index(a.get("rank", ""));
        if (rank < 0) fail("'" + a.get("rank", "") + "' isn't a valid value for '--rank <rank>'");
        if (rank == 0) fail("Snap to an actual rank.");
    }
    HostTaxonomy tax(taxon_file);
    const bool invalid = a.has("invalid");
    const std::vector<int64_t> snap = tax.filter_ancestors([&](uint64_t tid) {
        if (std::find(wanted.begin(), wanted.end(), tid) != wanted.end()) return true;
        const HostTaxonomy::Taxon& t = tax.taxa[tax.row_of[tid]];
        return (invalid || t.valid) && rank >= 0 && t.rank == rank;
    });
    BlockLines lines(stdin);
    const char* ls;
    size_t ll;
    std::string out;
    while (lines.next(ls, ll)) {
        if (ll && ls[0] == '>') {
            out.append(ls, ll);
        } else {
            const uint64_t t = parse_usize(std::string(ls, ll));
            if (t >= snap.size()) fail("index out of bounds: the len is " + std::to_string(snap.size()) + " but the index is " + std::to_string(t));
            out += std::to_string(snap[t] < 0 ? 0 : snap[t]);
        }
        out += '\n';
        if (out.size() > (1u << 20)) {
            put(stdout, out);
            out.clear();
        }
    }
    put(stdout, out);
    return 0;
}

// taxa2freq.rs:86-169
int cmd_taxa2freq(int argc, char** argv) {
    Args a = parse(argc, argv, 2, {{'r', "rank", true}, {'f', "frequency", true}});
    if (a.pos.empty()) fail("The following required arguments were not provided:\n    <taxon-file>");
    const int rank = HostTaxonomy::rank_index(a.get("rank", "species"));
    if (rank < 0) fail("'" + a.get("rank", "") + "' isn't a valid value for '--rank <rank>'");
    if (rank == 0) fail("Snap to an actual rank.");
    const uint64_t min_frequency = parse_usize(a.get("frequency", "1"));
    HostTaxonomy tax(a.pos[0]);
    const std::vector<std::string> files(a.pos.begin() + 1, a.pos.end());
    const size_t ncol = std::max<size_t>(1, files.size());
    const std::vector<int64_t> snap = tax.filter_ancestors([&](uint64_t tid) { return tax.taxa[tax.row_of[tid]].rank == rank; });
    std::string out = "taxon id,taxon name";
    if (files.empty()) out += ",stdin";
    for (const std::string& f : files) out += "," + f;
    out += '\n';
    std::map<uint64_t, std::vector<uint64_t>> counts;
    auto count_file = [&](FILE* f, size_t col) {
        BlockLines lines(f);
        const char* ls;
        size_t ll;
        while (lines.next(ls, ll)) {
            // lines that do not parse as a taxon id (FASTA headers) are skipped (taxa2freq.rs:160)
            size_t i = (ll && ls[0] == '+') ? 1 : 0;
            if (i >= ll) continue;
            uint64_t v = 0;
            bool ok = true;
            for (; i < ll && ok; ++i) {
                const unsigned d = (unsigned)(ls[i] - '0');
                ok = d <= 9 && v <= (UINT64_MAX - d) / 10;
                v = v * 10 + d;
            }
            if (!ok) continue;
            if (v >= snap.size()) fail("index out of bounds: the len is " + std::to_string(snap.size()) + " but the index is " + std::to_string(v));
            std::vector<uint64_t>& row = counts[snap[v] < 0 ? 0 : (uint64_t)snap[v]];
            if (row.empty()) row.assign(ncol, 0);
            row[col]++;
        }
    };
    if (files.empty()) {
        count_file(stdin, 0);
    } else {
        for (size_t i = 0; i < files.size(); ++i) {
            FILE* f = fopen(files[i].c_str(), "rb");
            if (!f) fail("No such file or directory (os error 2)");
            count_file(f, i);
            fclose(f);
        }
    }
    // rows by descending sum; the reference sorts a HashMap's entries by the sum (stable) and prints them reversed, so
    // rows with equal sums come in hash order there -- here by descending taxon id, a fixed choice among those orders
    std::vector<std::pair<uint64_t, std::vector<uint64_t>>> rows(counts.begin(), counts.end());
    auto total = [](const std::vector<uint64_t>& r) {
        uint64_t s = 0;
        for (uint64_t x : r) s += x;
        return s;
    };
    std::stable_sort(rows.begin(), rows.end(), [&](const auto& x, const auto& y) { return total(x.second) < total(y.second); });
    for (size_t i = rows.size(); i-- > 0;) {
        const uint64_t tid = rows[i].first;
        if (tid > tax.max_id || tax.row_of[tid] < 0) fail("LCA taxon id not in taxon list. Check compatibility with index.");
        if (total(rows[i].second) > min_frequency) {  // strictly greater, as written (taxa2freq.rs:141)
            const HostTaxonomy::Taxon& t = tax.taxa[tax.row_of[tid]];
            out += std::to_string(t.id) + "," + t.name;
            for (uint64_t c : rows[i].second) out += "," + std::to_string(c);
            out += '\n';
        }
    }
    put(stdout, out);
    return 0;
}

// bestof.rs:50-79: of every `frames` consecutive records the one with the most ids other than 0 and 1 (the last of
// equal maxima, Iterator::max_by_key).  As written, a group is emitted when its last record arrives and only the
// frames - 1 records before it are candidates: the record that completes the group is read but never pushed.
int cmd_bestof(int argc, char** argv) {
    Args a = parse(argc, argv, 2, {{'f', "frames", true}});
    const uint64_t frames = parse_usize(a.get("frames", "6"));
    if (frames == 0) fail("attempt to subtract with overflow");
    FastaReader rd(stdin, false);
    std::vector<Record> chunk;
    Record r;
    std::string out;
    while (rd.next(r)) {
        if (chunk.size() < frames - 1) {
            chunk.push_back(r);
            continue;
        }
        if (chunk.empty()) fail("called `Option::unwrap()` on a `None` value");  // -f 1: max_by_key of nothing
        size_t best = 0, best_n = 0;
        for (size_t i = 0; i < chunk.size(); ++i) {
            size_t n = 0;
            for (const std::string& tid : chunk[i].seq) {
                // tid.parse::<TaxonId>().unwrap_or(0), then everything but 0 and 1 counts
                bool ok = !tid.empty();
                size_t k = (ok && tid[0] == '+') ? 1 : 0;
                ok = ok && k < tid.size();
                unsigned __int128 v = 0;
                for (; ok && k < tid.size(); ++k) {
                    ok = tid[k] >= '0' && tid[k] <= '9';
                    v = v * 10 + (unsigned)(tid[k] - '0');
                    if (v > (unsigned __int128)UINT64_MAX) ok = false;
                }
                if (ok && v != 0 && v != 1) ++n;
            }
            if (n >= best_n) {
                best = i;
                best_n = n;
            }
        }
        write_record(out, chunk[best].header, chunk[best].seq, "\n", false);
        chunk.clear();
        if (out.size() > (1u << 20)) {
            put(stdout, out);
            out.clear();
        }
    }
    put(stdout, out);
    return 0;
}

}  // namespace


int main(int argc, char** argv) {
    try {
        if (argc < 2) {
            usage(stderr);
            return 1;
        }
        const std::string sub = argv[1];
        if (sub == "-V" || sub == "--version") {
            puts("umgap 1.1.1");
            return 0;
        }
        if (sub == "-h" || sub == "--help" || sub == "help") {
            usage(stdout);
            return 0;
        }
        if (sub == "translate") return cmd_translate(argc, argv);
        if (sub == "prot2kmer2lca") return cmd_prot2kmer2lca(argc, argv);
        if (sub == "prot2tryp2lca") return cmd_prot2tryp2lca(argc, argv);
        if (sub == "seedextend") return cmd_seedextend(argc, argv);
        if (sub == "taxa2agg") return cmd_taxa2agg(argc, argv);
        if (sub == "uniq") return cmd_uniq(argc, argv);
        if (sub == "fastq2fasta") return cmd_fastq2fasta(argc, argv);
        if (sub == "buildindex") return cmd_buildindex(argc, argv);
        if (sub == "printindex") return cmd_printindex(argc, argv);
        if (sub == "snaptaxon") return cmd_snaptaxon(argc, argv);
        if (sub == "taxa2freq") return cmd_taxa2freq(argc, argv);
        if (sub == "bestof") return cmd_bestof(argc, argv);
        if (sub == "classify") return cmd_classify(argc, argv);
        if (sub == "classify-peptides") return cmd_classify_peptides(argc, argv);
        fail("Found argument '" + sub + "' which wasn't expected, or isn't valid in this context");
    } catch (const Fail& e) {
        fprintf(stderr, "Error: %s\n", e.what());
        return 1;
    } catch (const std::exception& e) {
        fprintf(stderr, "Error: %s\n", e.what());
        return 1;
    }
}
