/*
 * umgap_ref.c -- CPU restatement ("port") of UMGAP's per-read classification path in plain C.
 *
 * TEST INFRASTRUCTURE ONLY: built into oracle/c/libumgap_ref.so and loaded by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  Nothing in the
 * product (umgap_b200/) links or calls it.
 *
 * It follows the reference's algorithm, not the GPU design: an in-memory `fst` image walked one
 * node per key byte (fst::Map::get, call site prot2kmer2lca.rs:176), codon-table translation
 * (dna/translation.rs:125-144, translate.rs:114-133), the seedextend state machine
 * (seedextend.rs:94-149), the uniq join (uniq.rs:56-84) and induced-tree aggregation
 * (tree/mod.rs:29-101, tree/lca.rs:34-40, tree/mix.rs:43-64, rmq/rtl.rs:39-57,
 * taxa2agg.rs:159-181), chunk-parallel over 240-record chunks like the reference's rayon
 * par_bridge (prot2kmer2lca.rs:163-166) -- here applied to the whole per-read chain, which is at
 * least as parallel as the reference's process pipeline.
 *
 * The `fst` crate (0.3.5, Cargo.toml:23) is not vendored in the reference; its format v2 node
 * encoding is restated from the published format.  Parity of the byte
 * format with files written by the real crate is UNPINNED (see oracle/__init__.py).
 */
#define _POSIX_C_SOURCE 200809L
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ fst */

static const char COMMON_INV[] =
    "te/oasripcnw.hlm-du012g=:bf3y5&_4v9678k%?xCDASFIBEjPTzRNM+LOqHGWUV,YKJZXQ;)(~[]$!'*@";

static int common_idx(uint8_t b) { /* 0 = not encodable, else index+1 (<= 63) */
    for (int i = 0; i < 63; ++i)
        if ((uint8_t)COMMON_INV[i] == b) return i + 1;
    return 0;
}

static inline uint64_t unpack(const uint8_t* p, unsigned n) {
    uint64_t v = 0;
    for (unsigned i = 0; i < n; ++i) v |= (uint64_t)p[i] << (8 * i);
    return v;
}
static unsigned pack_size(uint64_t v) {
    unsigned s = 1;
    while (s < 8 && v >= (1ull << (8 * s))) ++s;
    return s;
}

/* fst::Map::get: one node hop per key byte, outputs summed.  Returns 1 when found. */
int ref_fst_get(const uint8_t* d, uint64_t size, const uint8_t* key, uint32_t len, uint64_t* value) {
    if (size < 32) return 0;
    uint64_t addr = unpack(d + size - 8, 8);
    uint64_t out = 0;
    for (uint32_t kpos = 0; kpos <= len; ++kpos) {
        /* decode the node at addr far enough to find key[kpos] (or its final state at the end) */
        if (addr == 0) { /* EMPTY_ADDRESS: final, no transitions */
            if (kpos == len) { *value = out; return 1; }
            return 0;
        }
        const uint8_t s = d[addr];
        const unsigned kind = s >> 6;
        if (kind >= 2) { /* OneTransNext (11) / OneTrans (10): never final */
            if (kpos == len) return 0;
            const unsigned c = s & 0x3F, il = c == 0;
            const uint8_t inp = il ? d[addr - 1] : (uint8_t)COMMON_INV[c - 1];
            if (inp != key[kpos]) return 0;
            if (kind == 3) { addr = addr - il - 1; continue; }
            const uint8_t z = d[addr - il - 1];
            const unsigned tsz = z >> 4, osz = z & 15;
            const uint64_t dpos = addr - il - 1 - tsz;
            const uint64_t delta = unpack(d + dpos, tsz);
            const uint64_t start = dpos - osz;
            if (osz) out += unpack(d + start, osz);
            addr = delta ? start - delta : 0;
            continue;
        }
        const int final = (s & 0x40) != 0;
        unsigned n = s & 0x3F, nl = 0;
        if (n == 0) { nl = 1; n = d[addr - 1]; if (n == 1) n = 256; }
        const uint64_t base = addr - nl - 1;
        const uint8_t z = d[base];
        const unsigned tsz = z >> 4, osz = z & 15, isz = n > 32 ? 256 : 0;
        const uint64_t start = base - isz - n - (uint64_t)n * tsz - (uint64_t)n * osz - (final ? osz : 0);
        if (kpos == len) {
            if (!final) return 0;
            *value = out + (osz ? unpack(d + start, osz) : 0);
            return 1;
        }
        unsigned i;
        if (isz) {
            i = d[base - 256 + key[kpos]];
            if (i >= n) return 0;
        } else {
            for (i = 0; i < n; ++i)
                if (d[base - 1 - i] == key[kpos]) break;
            if (i == n) return 0;
        }
        const uint64_t delta = unpack(d + base - isz - n - (uint64_t)(i + 1) * tsz, tsz);
        if (osz) out += unpack(d + base - isz - n - (uint64_t)n * tsz - (uint64_t)(i + 1) * osz, osz);
        addr = delta ? start - delta : 0;
    }
    return 0;
}

/* Builder: the crate's incremental construction for sorted keys (shared prefixes, outputs pushed
 * down so that a path sums to the value), without the suffix-sharing registry -- minimisation only
 * changes the file size, never a lookup result. */
typedef struct { uint8_t inp; uint64_t out, addr; } BTrans;
typedef struct {
    int is_final;
    uint64_t final_out;
    BTrans* tr; int ntr, cap;
    int has_last; uint8_t last_inp; uint64_t last_out;
} BNode;
typedef struct {
    uint8_t* buf; uint64_t len, cap;
    BNode* stack; int depth, stack_cap; /* stack[0] = root; stack[i] reached by key[0..i) */
    uint8_t* prev; uint32_t prev_len, prev_cap;
    uint64_t last_addr, nkeys;
    int error;
} Builder;

static void b_put(Builder* b, const void* p, uint64_t n) {
    if (b->len + n > b->cap) {
        while (b->len + n > b->cap) b->cap = b->cap * 2 + 4096;
        b->buf = (uint8_t*)realloc(b->buf, b->cap);
    }
    memcpy(b->buf + b->len, p, n);
    b->len += n;
}
static void b_put_int(Builder* b, uint64_t v, unsigned n) {
    uint8_t t[8];
    for (unsigned i = 0; i < n; ++i) t[i] = (uint8_t)(v >> (8 * i));
    b_put(b, t, n);
}
static void node_reset(BNode* n) { n->is_final = 0; n->final_out = 0; n->ntr = 0; n->has_last = 0; }
static void node_push(BNode* n, uint8_t inp, uint64_t out, uint64_t addr) {
    if (n->ntr == n->cap) { n->cap = n->cap ? n->cap * 2 : 4; n->tr = (BTrans*)realloc(n->tr, n->cap * sizeof(BTrans)); }
    n->tr[n->ntr].inp = inp; n->tr[n->ntr].out = out; n->tr[n->ntr].addr = addr; n->ntr++;
}

/* Writes one node, returns its address (index of its state byte). */
static uint64_t emit_node(Builder* b, const BNode* n) {
    if (n->is_final && n->ntr == 0 && n->final_out == 0) return 0; /* EMPTY_ADDRESS */
    const uint64_t start = b->len;
    if (!n->is_final && n->ntr == 1) {
        const BTrans* t = &n->tr[0];
        const int ci = common_idx(t->inp);
        if (t->out == 0 && t->addr == b->last_addr && t->addr != 0 && start == t->addr + 1) { /* OneTransNext */
            if (!ci) b_put(b, &t->inp, 1);
            const uint8_t st = 0xC0 | (uint8_t)ci;
            b_put(b, &st, 1);
            return b->len - 1;
        }
        const uint64_t delta = t->addr ? start - t->addr : 0;
        const unsigned osz = t->out ? pack_size(t->out) : 0, tsz = pack_size(delta);
        if (osz) b_put_int(b, t->out, osz);
        b_put_int(b, delta, tsz);
        const uint8_t sizes = (uint8_t)((tsz << 4) | osz);
        b_put(b, &sizes, 1);
        if (!ci) b_put(b, &t->inp, 1);
        const uint8_t st = 0x80 | (uint8_t)ci;
        b_put(b, &st, 1);
        return b->len - 1;
    }
    /* AnyTrans */
    unsigned osz = 0, tsz = 1;
    int any_out = n->is_final && n->final_out;
    for (int i = 0; i < n->ntr; ++i) {
        if (n->tr[i].out) any_out = 1;
        const uint64_t delta = n->tr[i].addr ? start - n->tr[i].addr : 0;
        const unsigned s = pack_size(delta);
        if (s > tsz) tsz = s;
    }
    if (any_out) {
        osz = n->is_final ? pack_size(n->final_out) : 1;
        for (int i = 0; i < n->ntr; ++i) { const unsigned s = pack_size(n->tr[i].out); if (s > osz) osz = s; }
    }
    if (n->is_final && osz) b_put_int(b, n->final_out, osz);
    if (osz) for (int i = n->ntr - 1; i >= 0; --i) b_put_int(b, n->tr[i].out, osz);
    for (int i = n->ntr - 1; i >= 0; --i) b_put_int(b, n->tr[i].addr ? start - n->tr[i].addr : 0, tsz);
    for (int i = n->ntr - 1; i >= 0; --i) b_put(b, &n->tr[i].inp, 1);
    if (n->ntr > 32) {
        uint8_t index[256];
        memset(index, 255, 256);
        for (int i = 0; i < n->ntr; ++i) index[n->tr[i].inp] = (uint8_t)i;
        b_put(b, index, 256);
    }
    const uint8_t sizes = (uint8_t)((tsz << 4) | osz);
    b_put(b, &sizes, 1);
    uint8_t st = n->is_final ? 0x40 : 0;
    if (n->ntr >= 1 && n->ntr <= 63) st |= (uint8_t)n->ntr;
    else { const uint8_t cnt = n->ntr == 256 ? 1 : (uint8_t)n->ntr; b_put(b, &cnt, 1); }
    b_put(b, &st, 1);
    return b->len - 1;
}

/* Freezes stack[from+1 ..] bottom-up, linking each into its parent's pending transition. */
static void compile_from(Builder* b, int from) {
    while (b->depth > from) {
        BNode* n = &b->stack[b->depth];
        if (n->has_last) { b->error = 1; return; }
        const uint64_t addr = emit_node(b, n);
        if (addr) b->last_addr = addr;
        node_reset(n);
        b->depth--;
        BNode* par = &b->stack[b->depth];
        node_push(par, par->last_inp, par->last_out, addr);
        par->has_last = 0;
    }
}

static void builder_insert(Builder* b, const uint8_t* key, uint32_t len, uint64_t val) {
    if (b->nkeys) { /* keys must be strictly increasing (MapBuilder::insert, buildindex.rs:41-43) */
        const uint32_t mn = len < b->prev_len ? len : b->prev_len;
        const int c = memcmp(b->prev, key, mn);
        if (c > 0 || (c == 0 && b->prev_len >= len)) { b->error = 2; return; }
    }
    /* common prefix with the unfinished path, pushing output prefixes down (min under addition) */
    uint32_t p = 0;
    uint64_t out = val;
    while (p < len && (int)p < b->depth && b->stack[p].has_last && b->stack[p].last_inp == key[p]) {
        BNode* n = &b->stack[p];
        const uint64_t common = n->last_out < out ? n->last_out : out;
        const uint64_t push = n->last_out - common;
        if (push) {
            n->last_out = common;
            BNode* c = &b->stack[p + 1];
            if (c->is_final) c->final_out += push;
            for (int i = 0; i < c->ntr; ++i) c->tr[i].out += push;
            if (c->has_last) c->last_out += push;
        }
        out -= common;
        ++p;
    }
    compile_from(b, (int)p);
    if (b->error) return;
    if ((int)(len + 1) > b->stack_cap) {
        const int nc = (int)len + 16;
        b->stack = (BNode*)realloc(b->stack, nc * sizeof(BNode));
        memset(b->stack + b->stack_cap, 0, (nc - b->stack_cap) * sizeof(BNode));
        b->stack_cap = nc;
    }
    if (p == len) { b->error = 2; return; } /* duplicate key */
    for (uint32_t i = p; i < len; ++i) {
        BNode* n = &b->stack[i];
        n->has_last = 1; n->last_inp = key[i]; n->last_out = (i == p) ? out : 0;
        node_reset(&b->stack[i + 1]);
    }
    b->depth = (int)len;
    b->stack[len].is_final = 1;
    if (len > b->prev_cap) { b->prev_cap = len + 16; b->prev = (uint8_t*)realloc(b->prev, b->prev_cap); }
    memcpy(b->prev, key, len);
    b->prev_len = len;
    b->nkeys++;
}

/* Builds an fst Map image from sorted keys.  Returns a malloc'ed buffer (free with ref_free). */
uint8_t* ref_fst_build(const uint8_t* keys, const uint64_t* key_off, const uint64_t* values, uint64_t n,
                       uint64_t* out_size) {
    Builder b;
    memset(&b, 0, sizeof b);
    b.stack_cap = 64;
    b.stack = (BNode*)calloc(b.stack_cap, sizeof(BNode));
    b_put_int(&b, 2, 8); /* version */
    b_put_int(&b, 0, 8); /* type: Map */
    if (n && key_off[1] == key_off[0]) { /* the empty key */
        b.stack[0].is_final = 1;
        b.stack[0].final_out = values[0];
        b.nkeys = 1;
        for (uint64_t i = 1; i < n && !b.error; ++i)
            builder_insert(&b, keys + key_off[i], (uint32_t)(key_off[i + 1] - key_off[i]), values[i]);
    } else {
        for (uint64_t i = 0; i < n && !b.error; ++i)
            builder_insert(&b, keys + key_off[i], (uint32_t)(key_off[i + 1] - key_off[i]), values[i]);
    }
    uint8_t* result = NULL;
    if (!b.error) {
        compile_from(&b, 0);
        uint64_t root = emit_node(&b, &b.stack[0]);
        if (root == 0 && b.nkeys > 0) { /* a lone empty key with value 0: still needs a root byte */ }
        b_put_int(&b, b.nkeys, 8);
        b_put_int(&b, root, 8);
        result = b.buf;
        *out_size = b.len;
    } else {
        free(b.buf);
        *out_size = 0;
    }
    for (int i = 0; i < b.stack_cap; ++i) free(b.stack[i].tr);
    free(b.stack);
    free(b.prev);
    return result;
}
void ref_free(void* p) { free(p); }

/* ------------------------------------------------------------------------------------ taxonomy */

typedef struct {
    uint64_t max_id, root, n;
    int64_t* parent;   /* by id; -1 = no such taxon (TaxonList::ancestry, taxon.rs:158-163) */
    uint32_t* snap_valid;  /* by id; 0xFFFFFFFF = None (TaxonTree::snapping, taxon.rs:251-301) */
    uint32_t* snap_ranked;
} RefTax;

void* ref_tax_new(const uint64_t* ids, const uint64_t* parents, const uint8_t* rank, const uint8_t* valid, uint64_t n) {
    RefTax* t = (RefTax*)calloc(1, sizeof(RefTax));
    t->n = n;
    for (uint64_t i = 0; i < n; ++i) if (ids[i] > t->max_id) t->max_id = ids[i];
    const uint64_t m = t->max_id + 1;
    t->parent = (int64_t*)malloc(m * sizeof(int64_t));
    uint8_t* ok_v = (uint8_t*)calloc(m, 1), *ok_r = (uint8_t*)calloc(m, 1), *is_child = (uint8_t*)calloc(m, 1);
    for (uint64_t i = 0; i < m; ++i) t->parent[i] = -1;
    for (uint64_t i = 0; i < n; ++i) {
        t->parent[ids[i]] = (int64_t)parents[i];
        ok_v[ids[i]] = valid[i] != 0;
        ok_r[ids[i]] = valid[i] != 0 && rank[i] != 0;
        if (ids[i] != parents[i]) is_child[ids[i]] = 1;
    }
    int nroots = 0;
    for (uint64_t i = 0; i < m; ++i) if (t->parent[i] >= 0 && !is_child[i]) { if (!nroots++) t->root = i; }
    /* children lists (CSR) for the snapping DFS */
    uint64_t* cnt = (uint64_t*)calloc(m + 1, sizeof(uint64_t));
    for (uint64_t i = 0; i < n; ++i) if (ids[i] != parents[i] && parents[i] < m) cnt[parents[i] + 1]++;
    for (uint64_t i = 0; i < m; ++i) cnt[i + 1] += cnt[i];
    uint64_t* fill = (uint64_t*)malloc(m * sizeof(uint64_t));
    memcpy(fill, cnt, m * sizeof(uint64_t));
    uint64_t* child = (uint64_t*)malloc((n + 1) * sizeof(uint64_t));
    for (uint64_t i = 0; i < n; ++i) if (ids[i] != parents[i] && parents[i] < m) child[fill[parents[i]]++] = ids[i];
    t->snap_valid = (uint32_t*)malloc(m * sizeof(uint32_t));
    t->snap_ranked = (uint32_t*)malloc(m * sizeof(uint32_t));
    memset(t->snap_valid, 0xFF, m * sizeof(uint32_t));
    memset(t->snap_ranked, 0xFF, m * sizeof(uint32_t));
    if (nroots == 1) {
        uint64_t* st = (uint64_t*)malloc((n + 2) * sizeof(uint64_t));
        uint64_t sp = 0;
        st[sp++] = t->root;
        t->snap_valid[t->root] = ok_v[t->root] ? (uint32_t)t->root : (uint32_t)t->root;
        t->snap_ranked[t->root] = (uint32_t)t->root;
        uint8_t* seen = (uint8_t*)calloc(m, 1);
        seen[t->root] = 1;
        while (sp) {
            const uint64_t cur = st[--sp];
            for (uint64_t c = cnt[cur]; c < cnt[cur + 1]; ++c) {
                const uint64_t ch = child[c];
                if (seen[ch]) continue;
                seen[ch] = 1;
                t->snap_valid[ch] = ok_v[ch] ? (uint32_t)ch : t->snap_valid[cur];
                t->snap_ranked[ch] = ok_r[ch] ? (uint32_t)ch : t->snap_ranked[cur];
                st[sp++] = ch;
            }
        }
        free(st);
        free(seen);
    } else {
        t->n = 0; /* "More than one root!" / "There's no root!" */
    }
    free(ok_v); free(ok_r); free(is_child); free(cnt); free(fill); free(child);
    return t;
}
void ref_tax_free(void* p) {
    RefTax* t = (RefTax*)p;
    if (!t) return;
    free(t->parent); free(t->snap_valid); free(t->snap_ranked); free(t);
}

/* ----------------------------------------------------------------------------------- translate */

static const char* table_aas(int table, const char** starts) {
    static const struct { int id; const char* a; const char* s; } T[] = {
        {1, "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG", "---M---------------M---------------M----------------------------"},
        {2, "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNKKSS**VVVVAAAADDEEGGGG", "--------------------------------MMMM---------------M------------"},
        {3, "FFLLSSSSYY**CCWWTTTTPPPPHHQQRRRRIIMMTTTTNNKKSSRRVVVVAAAADDEEGGGG", "----------------------------------MM----------------------------"},
        {4, "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG", "--MM---------------M------------MMMM---------------M------------"},
        {5, "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNKKSSSSVVVVAAAADDEEGGGG", "---M----------------------------MMMM---------------M------------"},
        {6, "FFLLSSSSYYQQCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG", "-----------------------------------M----------------------------"},
        {9, "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNNKSSSSVVVVAAAADDEEGGGG", "-----------------------------------M---------------M------------"},
        {10, "FFLLSSSSYY**CCCWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG", "-----------------------------------M----------------------------"},
        {11, "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG", "---M---------------M------------MMMM---------------M------------"},
        {12, "FFLLSSSSYY**CC*WLLLSPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG", "-------------------M---------------M----------------------------"},
        {13, "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNKKSSGGVVVVAAAADDEEGGGG", "---M------------------------------MM---------------M------------"},
        {14, "FFLLSSSSYYY*CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNNKSSSSVVVVAAAADDEEGGGG", "-----------------------------------M----------------------------"},
        {15, "FFLLSSSSYY*QCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG", "-----------------------------------M----------------------------"},
        {16, "FFLLSSSSYY*LCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG", "-----------------------------------M----------------------------"},
        {21, "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNNKSSSSVVVVAAAADDEEGGGG", "-----------------------------------M---------------M------------"},
        {22, "FFLLSS*SYY*LCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG", "-----------------------------------M----------------------------"},
        {23, "FF*LSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG", "--------------------------------M--M---------------M------------"},
    };
    for (unsigned i = 0; i < sizeof T / sizeof T[0]; ++i)
        if (T[i].id == table) { *starts = T[i].s; return T[i].a; }
    return NULL;
}

static inline int nt_ord(uint8_t c) { return c == 'T' ? 0 : c == 'C' ? 1 : c == 'A' ? 2 : c == 'G' ? 3 : 4; }

/* translate_frame (translation.rs:136-144) of strand `fwd`/`rev` from offset f; returns length */
static uint32_t translate_frame(const uint8_t* lut65, const uint8_t* nt, uint32_t n, int rev, uint32_t f, uint8_t* out) {
    if (n <= f) return 0;
    const uint32_t plen = (n - f) / 3;
    for (uint32_t j = 0; j < plen; ++j) {
        int a, b, c;
        if (!rev) { a = nt_ord(nt[f + 3 * j]); b = nt_ord(nt[f + 3 * j + 1]); c = nt_ord(nt[f + 3 * j + 2]); }
        else {
            const uint32_t p = n - 1 - f - 3 * j; /* reverse strand = complement of the reversed read */
            a = nt_ord(nt[p]); b = nt_ord(nt[p - 1]); c = nt_ord(nt[p - 2]);
            if (a < 4) a ^= 2;
            if (b < 4) b ^= 2;
            if (c < 4) c ^= 2;
        }
        out[j] = (a | b | c) & 4 ? lut65[64] : lut65[16 * a + 4 * b + c];
    }
    return plen;
}

/* ---------------------------------------------------------------------------------- seedextend */

/* seedextend.rs:94-149 statement by statement; t has len+1 entries (sentinel 0 appended by caller).
 * Appends the selected ids to out; returns the new count. */
static uint32_t seedextend(const uint32_t* t, uint32_t tlen, uint32_t S, uint32_t G, uint32_t* out, uint32_t m) {
    uint32_t start = 0, end = 1, last = t[0], same = 1, smax = 1;
    while (end < tlen) {
        if (last == t[end]) { same++; end++; continue; }
        if (last == 0 && same > G) {
            if (smax >= S && end - same > start) for (uint32_t i = start; i < end - same; ++i) out[m++] = t[i];
            start = end; last = t[end]; same = 1; smax = 1; end++;
            continue;
        }
        if (last == 0 && end - start == same) { end++; start = end; continue; }
        if (last != 0 && same > smax) smax = same;
        last = t[end]; same = 1; end++;
    }
    if (smax >= S) {
        if (last == 0) end -= same;
        for (uint32_t i = start; i < end; ++i) out[m++] = t[i];
    }
    return m;
}

/* ----------------------------------------------------------------------------------- aggregate */

typedef struct { uint32_t id; float value; int first_child, next_sibling, nchildren, linked; } TNode;
typedef struct {
    uint32_t* keys; int* vals; uint32_t cap; /* open-addressing id -> node index */
    TNode* nodes; int nnodes, nodes_cap;
    uint32_t* queue; int qcap;
} AggScratch;

static int map_get(AggScratch* s, uint32_t id) {
    uint32_t h = (id * 2654435761u) & (s->cap - 1);
    while (s->keys[h] != 0xFFFFFFFFu) { if (s->keys[h] == id) return s->vals[h]; h = (h + 1) & (s->cap - 1); }
    return -1;
}
static void map_put(AggScratch* s, uint32_t id, int v) {
    uint32_t h = (id * 2654435761u) & (s->cap - 1);
    while (s->keys[h] != 0xFFFFFFFFu) h = (h + 1) & (s->cap - 1);
    s->keys[h] = id; s->vals[h] = v;
}
static int node_new(AggScratch* s, uint32_t id, float value) {
    if (s->nnodes == s->nodes_cap) { s->nodes_cap *= 2; s->nodes = (TNode*)realloc(s->nodes, s->nodes_cap * sizeof(TNode)); }
    TNode* n = &s->nodes[s->nnodes];
    n->id = id; n->value = value; n->first_child = -1; n->next_sibling = -1; n->nchildren = 0; n->linked = 0;
    map_put(s, id, s->nnodes);
    return s->nnodes++;
}

/* subtree sums after collapsing, computed on the fly: value(collapsed child) = sum of its subtree */
static float subtree_sum(const AggScratch* s, int n) {
    float v = s->nodes[n].value;
    for (int c = s->nodes[n].first_child; c >= 0; c = s->nodes[c].next_sibling) v += subtree_sum(s, c);
    return v;
}

/* taxa2agg for one record (taxa2agg.rs:159-181).  ids: the record's taxon ids (zeros allowed).
 * Returns the snapped taxon id, 1 for an empty record, 0xFFFFFFFE on UnknownTaxon (bad in *bad). */
static uint32_t aggregate_record(const RefTax* tax, AggScratch* s, const uint32_t* ids, uint32_t n, int strategy,
                                 float factor, float lower_bound, int ranked, uint32_t* bad) {
    /* count + filter (agg/mod.rs:27-44) into the node table: input taxa first */
    uint32_t need = 64;
    while (need < 8 * (n + 8)) need <<= 1;
    if (need > s->cap) {
        s->cap = need;
        s->keys = (uint32_t*)realloc(s->keys, s->cap * sizeof(uint32_t));
        s->vals = (int*)realloc(s->vals, s->cap * sizeof(int));
    }
    memset(s->keys, 0xFF, s->cap * sizeof(uint32_t));
    s->nnodes = 0;
    for (uint32_t i = 0; i < n; ++i) {
        if (ids[i] == 0) continue;
        const int k = map_get(s, ids[i]);
        if (k >= 0) s->nodes[k].value += 1.0f; else node_new(s, ids[i], 1.0f);
    }
    /* lower bound: drop by zeroing and compacting */
    int kept = 0;
    for (int i = 0; i < s->nnodes; ++i) if (s->nodes[i].value >= lower_bound) kept++;
    if (kept == 0) return 1u; /* literal "1" (taxa2agg.rs:174-175) */
    if (kept != s->nnodes) {
        memset(s->keys, 0xFF, s->cap * sizeof(uint32_t));
        int w = 0;
        for (int i = 0; i < s->nnodes; ++i)
            if (s->nodes[i].value >= lower_bound) { s->nodes[w] = s->nodes[i]; map_put(s, s->nodes[w].id, w); ++w; }
        s->nnodes = w;
    }
    const int ninput = s->nnodes;
    uint32_t result;
    if (strategy == 2) { /* MRTL, rmq/rtl.rs:39-57: last maximum in iteration order wins */
        float best = -1.0f;
        result = 0;
        for (int i = 0; i < ninput; ++i) {
            float c = s->nodes[i].value;
            uint64_t next = s->nodes[i].id;
            for (;;) {
                if (next == tax->root) break;
                if (next > tax->max_id || tax->parent[next] < 0) break;
                const uint64_t anc = (uint64_t)tax->parent[next];
                const int k = anc <= 0xFFFFFFFEull ? map_get(s, (uint32_t)anc) : -1;
                if (k >= 0 && k < ninput) c += s->nodes[k].value;
                if (anc == next) break;
                next = anc;
            }
            if (next != tax->root) { *bad = (uint32_t)next; return 0xFFFFFFFEu; }
            if (c >= best) { best = c; result = s->nodes[i].id; }
        }
    } else {
        /* Tree::new (tree/mod.rs:29-48): link every taxon to its parent, queueing new ancestors */
        if (s->qcap < 64) { s->qcap = 64; s->queue = (uint32_t*)realloc(s->queue, s->qcap * sizeof(uint32_t)); }
        int qh = 0, qt = 0;
        for (int i = 0; i < ninput; ++i) {
            if (qt == s->qcap) { s->qcap *= 2; s->queue = (uint32_t*)realloc(s->queue, s->qcap * sizeof(uint32_t)); }
            s->queue[qt++] = s->nodes[i].id;
        }
        while (qh < qt) {
            const uint32_t id = s->queue[qh++];
            if (id > tax->max_id || tax->parent[id] < 0) { *bad = id; return 0xFFFFFFFEu; }
            const uint64_t par = (uint64_t)tax->parent[id];
            if (par == id) continue;
            if (s->nnodes * 4 >= (int)s->cap) { /* grow the map */
                s->cap *= 2;
                s->keys = (uint32_t*)realloc(s->keys, s->cap * sizeof(uint32_t));
                s->vals = (int*)realloc(s->vals, s->cap * sizeof(int));
                memset(s->keys, 0xFF, s->cap * sizeof(uint32_t));
                for (int i = 0; i < s->nnodes; ++i) map_put(s, s->nodes[i].id, i);
            }
            int pk = map_get(s, (uint32_t)par);
            if (pk < 0) pk = node_new(s, (uint32_t)par, 0.0f);
            if (s->nodes[pk].nchildren == 0) { /* !tree.contains_key(parent): visit the parent too */
                if (qt == s->qcap) { s->qcap *= 2; s->queue = (uint32_t*)realloc(s->queue, s->qcap * sizeof(uint32_t)); }
                s->queue[qt++] = (uint32_t)par;
            }
            const int me = map_get(s, id);
            if (!s->nodes[me].linked) { /* siblings is a HashSet: inserting twice is a no-op */
                s->nodes[me].linked = 1;
                s->nodes[me].next_sibling = s->nodes[pk].first_child;
                s->nodes[pk].first_child = me;
                s->nodes[pk].nchildren++;
            }
        }
        int root = map_get(s, (uint32_t)tax->root);
        if (root < 0) root = node_new(s, (uint32_t)tax->root, 0.0f); /* cannot happen for a valid tree */
        /* collapse from the root (tree/mod.rs:71-86) */
        int base = root;
        while (s->nodes[base].nchildren == 1) base = s->nodes[base].first_child;
        if (strategy == 1) { /* hybrid, tree/mix.rs:43-64 */
            float bval = subtree_sum(s, root);
            for (;;) {
                int best = -1;
                float bm = 0.0f;
                for (int c = s->nodes[base].first_child; c >= 0; c = s->nodes[c].next_sibling) {
                    const float v = subtree_sum(s, c);
                    if (best < 0 || v >= bm) { best = c; bm = v; }
                }
                if (best < 0) break;
                if (bm / bval < factor) break;
                base = best;
                while (s->nodes[base].nchildren == 1) base = s->nodes[base].first_child;
                bval = bm;
            }
        }
        result = s->nodes[base].id;
    }
    const uint32_t sn = ranked ? tax->snap_ranked[result] : tax->snap_valid[result];
    if (sn == 0xFFFFFFFFu) { *bad = result; return 0xFFFFFFFEu; }
    return sn;
}

/* ------------------------------------------------------------------------------------ pipeline */

typedef struct {
    int table, methionine, one_on_one, seedextend, min_seed_size, max_gap_size, strategy;
    float factor, lower_bound;
    int ranked_only, k;
} RefOpts;

typedef struct {
    const uint8_t* img; uint64_t img_size;
    const RefTax* tax;
    RefOpts o;
    const uint8_t* nt; const uint64_t* read_off; const uint64_t* group_off; uint64_t ngroups;
    uint32_t* out;
    uint64_t next; pthread_mutex_t* mu;
    uint64_t lookups, hits;
    uint8_t lut[65];
    int error; uint32_t bad;
    int lookups_only;
} Job;

static void* worker(void* arg) {
    Job* job = (Job*)arg;
    const RefOpts* o = &job->o;
    AggScratch s;
    memset(&s, 0, sizeof s);
    s.nodes_cap = 256;
    s.nodes = (TNode*)malloc(s.nodes_cap * sizeof(TNode));
    uint32_t cap = 4096;
    uint8_t* pep = (uint8_t*)malloc(cap);
    uint32_t* ids = (uint32_t*)malloc((cap + 1) * sizeof(uint32_t));
    uint32_t kept_cap = 8192, *kept = (uint32_t*)malloc(kept_cap * sizeof(uint32_t));
    uint64_t lookups = 0, hits = 0;
    const uint64_t chunk = 20; /* 240 records = 20 pairs x 12 frame records (prot2kmer2lca.rs:97-104) */
    for (;;) {
        pthread_mutex_lock(job->mu);
        const uint64_t g0 = job->next;
        job->next += chunk;
        pthread_mutex_unlock(job->mu);
        if (g0 >= job->ngroups) break;
        const uint64_t g1 = g0 + chunk < job->ngroups ? g0 + chunk : job->ngroups;
        for (uint64_t g = g0; g < g1; ++g) {
            uint32_t m = 0;
            int present = 0;
            for (uint64_t r = job->group_off[g]; r < job->group_off[g + 1]; ++r) {
                const uint8_t* nt = job->nt + job->read_off[r];
                const uint32_t n = (uint32_t)(job->read_off[r + 1] - job->read_off[r]);
                if (n / 3 + 2 > cap) {
                    cap = n / 3 + 64;
                    pep = (uint8_t*)realloc(pep, cap);
                    ids = (uint32_t*)realloc(ids, (cap + 1) * sizeof(uint32_t));
                }
                for (int fr = 0; fr < 6; ++fr) {
                    const uint32_t plen = translate_frame(job->lut, nt, n, fr >= 3, (uint32_t)(fr % 3), pep);
                    if (plen < (uint32_t)o->k) continue; /* record dropped, header and all (:172) */
                    present = 1;
                    uint32_t cnt = 0;
                    for (uint32_t i = 0; i + o->k <= plen; ++i) {
                        uint64_t v;
                        ++lookups;
                        if (ref_fst_get(job->img, job->img_size, pep + i, (uint32_t)o->k, &v)) { ids[cnt++] = (uint32_t)v; ++hits; }
                        else if (o->one_on_one) ids[cnt++] = 0;
                    }
                    if (job->lookups_only) continue;
                    if (m + cnt + 1 > kept_cap) { kept_cap = 2 * (m + cnt + 1); kept = (uint32_t*)realloc(kept, kept_cap * sizeof(uint32_t)); }
                    if (o->seedextend) {
                        ids[cnt] = 0; /* sentinel (seedextend.rs:99) */
                        m = seedextend(ids, cnt + 1, (uint32_t)o->min_seed_size, (uint32_t)o->max_gap_size, kept, m);
                    } else {
                        memcpy(kept + m, ids, cnt * sizeof(uint32_t));
                        m += cnt;
                    }
                }
            }
            if (job->lookups_only) continue;
            uint32_t res = 0xFFFFFFFFu; /* no record at all */
            if (present) {
                uint32_t bad = 0;
                res = aggregate_record(job->tax, &s, kept, m, o->strategy, o->factor, o->lower_bound, o->ranked_only, &bad);
                if (res == 0xFFFFFFFEu) { job->error = 1; job->bad = bad; res = 0xFFFFFFFFu; }
            }
            job->out[g] = res;
        }
    }
    pthread_mutex_lock(job->mu);
    job->lookups += lookups;
    job->hits += hits;
    pthread_mutex_unlock(job->mu);
    free(s.keys); free(s.vals); free(s.nodes); free(s.queue); free(pep); free(ids); free(kept);
    return NULL;
}

/* translate -a | prot2kmer2lca [-o] | [seedextend] | uniq -d | taxa2agg over `threads` threads.
 * Returns 0, -1 on an unknown table, -4 on UnknownTaxon (*bad_taxon set). */
int ref_classify(const uint8_t* img, uint64_t img_size, const void* tax, const RefOpts* opts, const uint8_t* nt,
                 const uint64_t* read_off, const uint64_t* group_off, uint64_t ngroups, uint32_t* out, int threads,
                 int lookups_only, uint64_t* n_lookups, uint64_t* n_hits, uint32_t* bad_taxon) {
    const char* starts;
    const char* aas = table_aas(opts->table, &starts);
    if (!aas) return -1;
    Job job;
    memset(&job, 0, sizeof job);
    for (int i = 0; i < 64; ++i) job.lut[i] = (opts->methionine && starts[i] == 'M') ? 'M' : (uint8_t)aas[i];
    job.lut[64] = '-';
    pthread_mutex_t mu;
    pthread_mutex_init(&mu, NULL);
    job.img = img; job.img_size = img_size; job.tax = (const RefTax*)tax; job.o = *opts;
    job.nt = nt; job.read_off = read_off; job.group_off = group_off; job.ngroups = ngroups; job.out = out;
    job.mu = &mu; job.lookups_only = lookups_only;
    if (threads < 1) threads = 1;
    pthread_t* th = (pthread_t*)malloc(threads * sizeof(pthread_t));
    for (int i = 0; i < threads; ++i) pthread_create(&th[i], NULL, worker, &job);
    for (int i = 0; i < threads; ++i) pthread_join(th[i], NULL);
    free(th);
    pthread_mutex_destroy(&mu);
    if (n_lookups) *n_lookups = job.lookups;
    if (n_hits) *n_hits = job.hits;
    if (job.error) { if (bad_taxon) *bad_taxon = job.bad; return -4; }
    return 0;
}

/* Single-record entry points for the parity tests of the port against the Python oracle. */
uint32_t ref_seedextend(const uint32_t* ids, uint32_t n, uint32_t S, uint32_t G, uint32_t* out) {
    uint32_t* t = (uint32_t*)malloc((n + 1) * sizeof(uint32_t));
    memcpy(t, ids, n * sizeof(uint32_t));
    t[n] = 0;
    const uint32_t m = seedextend(t, n + 1, S, G, out, 0);
    free(t);
    return m;
}
uint32_t ref_aggregate(const void* tax, const uint32_t* ids, uint32_t n, int strategy, float factor, float lower_bound,
                       int ranked, uint32_t* bad) {
    AggScratch s;
    memset(&s, 0, sizeof s);
    s.nodes_cap = 256;
    s.nodes = (TNode*)malloc(s.nodes_cap * sizeof(TNode));
    const uint32_t r = aggregate_record((const RefTax*)tax, &s, ids, n, strategy, factor, lower_bound, ranked, bad);
    free(s.keys); free(s.vals); free(s.nodes); free(s.queue);
    return r;
}
uint32_t ref_translate(int table, int methionine, const uint8_t* nt, uint32_t n, int frame /*0..5*/, uint8_t* out) {
    const char* starts;
    const char* aas = table_aas(table, &starts);
    if (!aas) return 0xFFFFFFFFu;
    uint8_t lut[65];
    for (int i = 0; i < 64; ++i) lut[i] = (methionine && starts[i] == 'M') ? 'M' : (uint8_t)aas[i];
    lut[64] = '-';
    return translate_frame(lut, nt, n, frame >= 3, (uint32_t)(frame % 3), out);
}

/* ------------------------------------------------------------------------------ peptide pipeline */
/* prot2tryp2lca | uniq -d / | taxa2agg (the tryptic presets, scripts/umgap-analyse.sh:291-300).
 * prot2tryp2lca.rs:105-134 per LINE: the default pattern ([KR])([^P]) applied twice and '*' -> newline is the
 * closed form "a peptide ends after K/R unless the next residue is P (or the line ends), and at '*'"
 * (oracle/lookup.py: tryptic_digest, checked against the two regex passes); empty pieces drop out (:119), the
 * byte length must lie in minlen..maxlen (:120-123), optional keep / drop residue sets (:124-129), every survivor
 * is one fst::Map::get over the whole peptide (:130); misses are omitted (no -o in the presets). */
typedef struct {
    int minlen, maxlen;
    const char* keep;
    const char* drop;
    int strategy;
    float factor, lower_bound;
    int ranked_only;
} RefTrypOpts;

typedef struct {
    const uint8_t* img; uint64_t img_size;
    const RefTax* tax;
    RefTrypOpts o;
    const uint8_t* aa; const uint64_t* line_off; const uint64_t* group_off; uint64_t ngroups;
    uint32_t* out;
    uint64_t next; pthread_mutex_t* mu;
    uint64_t lookups, hits;
    int error; uint32_t bad;
} PepJob;

static int pep_passes_sets(const uint8_t* p, uint32_t n, const char* keep, const char* drop) {
    if ((!keep || !*keep) && (!drop || !*drop)) return 1;
    uint8_t seen[256];
    memset(seen, 0, sizeof seen);
    for (uint32_t i = 0; i < n; ++i) seen[p[i]] = 1;
    if (keep) for (const char* c = keep; *c; ++c) if (!seen[(uint8_t)*c]) return 0;
    if (drop) for (const char* c = drop; *c; ++c) if (seen[(uint8_t)*c]) return 0;
    return 1;
}

static void* pep_worker(void* arg) {
    PepJob* job = (PepJob*)arg;
    const RefTrypOpts* o = &job->o;
    AggScratch s;
    memset(&s, 0, sizeof s);
    s.nodes_cap = 256;
    s.nodes = (TNode*)malloc(s.nodes_cap * sizeof(TNode));
    uint32_t kept_cap = 1024, *kept = (uint32_t*)malloc(kept_cap * sizeof(uint32_t));
    uint64_t lookups = 0, hits = 0;
    const uint64_t chunk = 120; /* 240 records = 120 pairs x 2 peptide records */
    for (;;) {
        pthread_mutex_lock(job->mu);
        const uint64_t g0 = job->next;
        job->next += chunk;
        pthread_mutex_unlock(job->mu);
        if (g0 >= job->ngroups) break;
        const uint64_t g1 = g0 + chunk < job->ngroups ? g0 + chunk : job->ngroups;
        for (uint64_t g = g0; g < g1; ++g) {
            uint32_t m = 0;
            for (uint64_t l = job->group_off[g]; l < job->group_off[g + 1]; ++l) {
                const uint8_t* line = job->aa + job->line_off[l];
                const uint32_t n = (uint32_t)(job->line_off[l + 1] - job->line_off[l]);
                uint32_t start = 0;
                for (uint32_t i = 0; i <= n; ++i) {
                    int cut = 0;
                    uint32_t end = i;
                    if (i == n) cut = 1;
                    else if (line[i] == '*') cut = 1;
                    else if ((line[i] == 'K' || line[i] == 'R') && i + 1 < n && line[i + 1] != 'P') { cut = 1; end = i + 1; }
                    if (!cut) continue;
                    const uint32_t len = end - start;
                    if (len && (int)len >= o->minlen && (int)len <= o->maxlen && pep_passes_sets(line + start, len, o->keep, o->drop)) {
                        uint64_t v;
                        ++lookups;
                        if (ref_fst_get(job->img, job->img_size, line + start, len, &v)) {
                            if (m + 1 > kept_cap) { kept_cap *= 2; kept = (uint32_t*)realloc(kept, kept_cap * sizeof(uint32_t)); }
                            kept[m++] = (uint32_t)v;
                            ++hits;
                        }
                    }
                    start = (i < n && line[i] == '*') ? i + 1 : end;
                }
            }
            uint32_t res = 0xFFFFFFFFu; /* a group without lines has no record */
            if (job->group_off[g + 1] > job->group_off[g]) {
                uint32_t bad = 0;
                res = aggregate_record(job->tax, &s, kept, m, o->strategy, o->factor, o->lower_bound, o->ranked_only, &bad);
                if (res == 0xFFFFFFFEu) { job->error = 1; job->bad = bad; res = 0xFFFFFFFFu; }
            }
            job->out[g] = res;
        }
    }
    pthread_mutex_lock(job->mu);
    job->lookups += lookups;
    job->hits += hits;
    pthread_mutex_unlock(job->mu);
    free(s.keys); free(s.vals); free(s.nodes); free(s.queue); free(kept);
    return NULL;
}

int ref_classify_peptides(const uint8_t* img, uint64_t img_size, const void* tax, const RefTrypOpts* opts, const uint8_t* aa,
                          const uint64_t* line_off, const uint64_t* group_off, uint64_t ngroups, uint32_t* out, int threads,
                          uint64_t* n_lookups, uint64_t* n_hits, uint32_t* bad_taxon) {
    PepJob job;
    memset(&job, 0, sizeof job);
    pthread_mutex_t mu;
    pthread_mutex_init(&mu, NULL);
    job.img = img; job.img_size = img_size; job.tax = (const RefTax*)tax; job.o = *opts;
    job.aa = aa; job.line_off = line_off; job.group_off = group_off; job.ngroups = ngroups; job.out = out; job.mu = &mu;
    if (threads < 1) threads = 1;
    pthread_t* th = (pthread_t*)malloc(threads * sizeof(pthread_t));
    for (int i = 0; i < threads; ++i) pthread_create(&th[i], NULL, pep_worker, &job);
    for (int i = 0; i < threads; ++i) pthread_join(th[i], NULL);
    free(th);
    pthread_mutex_destroy(&mu);
    if (n_lookups) *n_lookups = job.lookups;
    if (n_hits) *n_hits = job.hits;
    if (job.error) { if (bad_taxon) *bad_taxon = job.bad; return -4; }
    return 0;
}

/* ------------------------------------------------------------- the reference's process structure */
/* scripts/umgap-analyse.sh:276-290 runs five processes connected by pipes: translate -a | prot2kmer2lca -o
 * | seedextend | uniq -d / | taxa2agg.  Only the lookups are multi-threaded (rayon over 240-record chunks,
 * prot2kmer2lca.rs:163-166); every stage parses FASTA text and prints FASTA text (io/fasta.rs:30-67,164-180).
 * ref_pipeline_staged runs the same five stages one after the other over in-memory text, each as the
 * reference structures it, and reports the seconds each took: as concurrent processes the pipeline moves at
 * the pace of its slowest stage. */
#include <stdio.h>
#include <time.h>

typedef struct { char* p; size_t n, cap; } Buf;
static void buf_need(Buf* b, size_t extra) {
    if (b->n + extra > b->cap) {
        while (b->n + extra > b->cap) b->cap = b->cap ? b->cap * 2 : (1u << 20);
        b->p = (char*)realloc(b->p, b->cap);
    }
}
static void buf_put(Buf* b, const char* s, size_t n) { buf_need(b, n); memcpy(b->p + b->n, s, n); b->n += n; }
static void buf_putc(Buf* b, char c) { buf_need(b, 1); b->p[b->n++] = c; }
static void buf_put_u32(Buf* b, uint32_t v) {
    char tmp[12];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    buf_need(b, (size_t)n);
    while (n) b->p[b->n++] = tmp[--n];
}
static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
/* one record of a FASTA stream: header [h, h+hn), item lines [body, body_end) (fasta.rs:38-67) */
typedef struct { const char* h; size_t hn; const char* body; const char* body_end; } Rec;
static const char* next_record(const char* p, const char* end, Rec* r) {
    const char* e = (const char*)memchr(p, '\n', (size_t)(end - p));
    const char* hend = e ? e : end;
    r->h = p + 1;
    r->hn = (size_t)(hend - p - 1);
    p = e ? e + 1 : end;
    r->body = p;
    while (p < end && *p != '>') {
        const char* le = (const char*)memchr(p, '\n', (size_t)(end - p));
        p = le ? le + 1 : end;
    }
    r->body_end = p;
    return p;
}

typedef struct {
    const uint8_t* img; uint64_t img_size; int k;
    const Rec* recs; uint64_t nrecs; uint64_t next; pthread_mutex_t* mu;
    Buf* outs; /* one per chunk, concatenated in order afterwards */
    uint64_t lookups;
} KJob;
static void* k_worker(void* arg) {
    KJob* j = (KJob*)arg;
    uint64_t lookups = 0;
    for (;;) {
        pthread_mutex_lock(j->mu);
        const uint64_t c = j->next++;
        pthread_mutex_unlock(j->mu);
        const uint64_t r0 = c * 240;
        if (r0 >= j->nrecs) break;
        const uint64_t r1 = r0 + 240 < j->nrecs ? r0 + 240 : j->nrecs;
        Buf* o = &j->outs[c];
        for (uint64_t r = r0; r < r1; ++r) {
            const Rec* rc = &j->recs[r];
            /* unwrap = true: the item lines concatenated; translate prints one line, so it is the line itself */
            const char* s = rc->body;
            size_t n = (size_t)(rc->body_end - rc->body);
            while (n && (s[n - 1] == '\n' || s[n - 1] == '\r')) --n;
            if (n < (size_t)j->k) continue;                           /* :172 */
            buf_putc(o, '>'); buf_put(o, rc->h, rc->hn); buf_putc(o, '\n');
            for (size_t i = 0; i + (size_t)j->k <= n; ++i) {
                uint64_t v = 0;
                ++lookups;
                if (!ref_fst_get(j->img, j->img_size, (const uint8_t*)s + i, (uint32_t)j->k, &v)) v = 0;   /* -o */
                buf_put_u32(o, (uint32_t)v); buf_putc(o, '\n');
            }
        }
    }
    pthread_mutex_lock(j->mu);
    j->lookups += lookups;
    pthread_mutex_unlock(j->mu);
    return NULL;
}

static uint32_t parse_ids(const Rec* r, uint32_t** ids, uint32_t* cap) {
    uint32_t n = 0;
    const char* p = r->body;
    while (p < r->body_end) {
        uint32_t v = 0;
        while (p < r->body_end && *p >= '0' && *p <= '9') v = v * 10 + (uint32_t)(*p++ - '0');
        while (p < r->body_end && (*p == '\n' || *p == '\r')) ++p;
        if (n + 2 > *cap) { *cap = 2 * (*cap) + 64; *ids = (uint32_t*)realloc(*ids, *cap * sizeof(uint32_t)); }
        (*ids)[n++] = v;
    }
    return n;
}

/* fasta: `>h1\nACGT...\n>h2\n...`; out_taxa receives one taxon per uniq group (as many as it has room for);
 * stage_s[5] = seconds of translate, prot2kmer2lca, seedextend, uniq, taxa2agg.  Returns the number of groups,
 * or a negative error. */
int64_t ref_pipeline_staged(const uint8_t* img, uint64_t img_size, const void* taxp, const RefOpts* o, const char* fasta,
                            uint64_t fasta_len, int threads, uint32_t* out_taxa, uint64_t out_cap, double* stage_s,
                            uint64_t* n_lookups) {
    const RefTax* tax = (const RefTax*)taxp;
    const char* starts;
    const char* aas = table_aas(o->table, &starts);
    if (!aas) return -1;
    uint8_t lut[65];
    for (int i = 0; i < 64; ++i) lut[i] = (o->methionine && starts[i] == 'M') ? 'M' : (uint8_t)aas[i];
    lut[64] = '-';
    /* ---- translate -a (single thread) */
    double t0 = now_s();
    Buf tb = {0, 0, 0};
    {
        const char* p = fasta;
        const char* end = fasta + fasta_len;
        uint8_t* pep = (uint8_t*)malloc(1 << 16);
        while (p < end) {
            Rec r;
            p = next_record(p, end, &r);
            const char* s = r.body;
            size_t n = (size_t)(r.body_end - r.body);
            while (n && (s[n - 1] == '\n' || s[n - 1] == '\r')) --n;
            for (int fr = 0; fr < 6; ++fr) {
                const uint32_t plen = translate_frame(lut, (const uint8_t*)s, (uint32_t)n, fr >= 3, (uint32_t)(fr % 3), pep);
                buf_putc(&tb, '>'); buf_put(&tb, r.h, r.hn); buf_putc(&tb, '\n');
                if (plen) { buf_put(&tb, (const char*)pep, plen); buf_putc(&tb, '\n'); }
            }
        }
        free(pep);
    }
    stage_s[0] = now_s() - t0;
    /* ---- prot2kmer2lca -o (reader single-threaded, 240-record chunks over `threads` threads, output in order) */
    t0 = now_s();
    Buf kb = {0, 0, 0};
    {
        uint64_t nrecs = 0, cap = 1 << 16;
        Rec* recs = (Rec*)malloc(cap * sizeof(Rec));
        const char* p = tb.p;
        const char* end = tb.p + tb.n;
        while (p < end) {
            if (nrecs == cap) { cap *= 2; recs = (Rec*)realloc(recs, cap * sizeof(Rec)); }
            p = next_record(p, end, &recs[nrecs++]);
        }
        const uint64_t nchunks = (nrecs + 239) / 240;
        KJob job;
        memset(&job, 0, sizeof job);
        pthread_mutex_t mu;
        pthread_mutex_init(&mu, NULL);
        job.img = img; job.img_size = img_size; job.k = o->k; job.recs = recs; job.nrecs = nrecs; job.mu = &mu;
        job.outs = (Buf*)calloc(nchunks ? nchunks : 1, sizeof(Buf));
        if (threads < 1) threads = 1;
        pthread_t* th = (pthread_t*)malloc((size_t)threads * sizeof(pthread_t));
        for (int i = 0; i < threads; ++i) pthread_create(&th[i], NULL, k_worker, &job);
        for (int i = 0; i < threads; ++i) pthread_join(th[i], NULL);
        for (uint64_t c = 0; c < nchunks; ++c) { buf_put(&kb, job.outs[c].p, job.outs[c].n); free(job.outs[c].p); }
        if (n_lookups) *n_lookups = job.lookups;
        free(th); free(job.outs); free(recs);
        pthread_mutex_destroy(&mu);
    }
    stage_s[1] = now_s() - t0;
    free(tb.p);
    /* ---- seedextend (single thread) */
    t0 = now_s();
    Buf sb = {0, 0, 0};
    Buf* cur = &kb;
    uint32_t idcap = 256, *ids = (uint32_t*)malloc(idcap * sizeof(uint32_t)), *sel = NULL, selcap = 0;
    if (o->seedextend) {
        const char* p = kb.p;
        const char* end = kb.p + kb.n;
        while (p < end) {
            Rec r;
            p = next_record(p, end, &r);
            uint32_t n = parse_ids(&r, &ids, &idcap);
            ids[n] = 0;
            if (n + 2 > selcap) { selcap = 2 * n + 64; sel = (uint32_t*)realloc(sel, selcap * sizeof(uint32_t)); }
            const uint32_t m = seedextend(ids, n + 1, (uint32_t)o->min_seed_size, (uint32_t)o->max_gap_size, sel, 0);
            buf_putc(&sb, '>'); buf_put(&sb, r.h, r.hn); buf_putc(&sb, '\n');
            for (uint32_t i = 0; i < m; ++i) { buf_put_u32(&sb, sel[i]); buf_putc(&sb, '\n'); }
        }
        cur = &sb;
    }
    stage_s[2] = now_s() - t0;
    /* ---- uniq -d / (single thread) */
    t0 = now_s();
    Buf ub = {0, 0, 0};
    {
        const char* p = cur->p;
        const char* end = cur->p + cur->n;
        const char* last = NULL;
        size_t lastn = 0;
        while (p < end) {
            Rec r;
            p = next_record(p, end, &r);
            const char* d = (const char*)memchr(r.h, '/', r.hn);
            const size_t hn = d ? (size_t)(d - r.h) : r.hn;
            if (!last || lastn != hn || memcmp(last, r.h, hn) != 0) {
                buf_putc(&ub, '>'); buf_put(&ub, r.h, hn); buf_putc(&ub, '\n');
                last = r.h;   /* points into cur, which outlives this loop */
                lastn = hn;
            }
            buf_put(&ub, r.body, (size_t)(r.body_end - r.body));
        }
    }
    stage_s[3] = now_s() - t0;
    free(kb.p); free(sb.p);
    /* ---- taxa2agg (single thread) */
    t0 = now_s();
    int64_t ngroups = 0;
    int err = 0;
    {
        AggScratch s;
        memset(&s, 0, sizeof s);
        s.nodes_cap = 256;
        s.nodes = (TNode*)malloc(s.nodes_cap * sizeof(TNode));
        Buf ab = {0, 0, 0};
        const char* p = ub.p;
        const char* end = ub.p + ub.n;
        while (p < end) {
            Rec r;
            p = next_record(p, end, &r);
            const uint32_t n = parse_ids(&r, &ids, &idcap);
            uint32_t bad = 0;
            const uint32_t res = aggregate_record(tax, &s, ids, n, o->strategy, o->factor, o->lower_bound, o->ranked_only, &bad);
            if (res == 0xFFFFFFFEu) { err = 1; break; }
            buf_putc(&ab, '>'); buf_put(&ab, r.h, r.hn); buf_putc(&ab, '\n'); buf_put_u32(&ab, res); buf_putc(&ab, '\n');
            if ((uint64_t)ngroups < out_cap) out_taxa[ngroups] = res;
            ++ngroups;
        }
        free(ab.p);
        free(s.keys); free(s.vals); free(s.nodes); free(s.queue);
    }
    stage_s[4] = now_s() - t0;
    free(ub.p); free(ids); free(sel);
    return err ? -4 : ngroups;
}

/* ------------------------------------------------------------ synthetic workload (bench aid) ----
 * C mirror of oracle/synth.py (itself the numpy mirror of umgap_b200/csrc/synth.cu, checked bit for bit
 * against the device generator): the counter-based proteome, its 9-mer index as an fst image, and the
 * reads.  Lets bench.py's CPU legs hold an index of 1e8+ keys without minutes of numpy.  Not part of
 * the reference. */

static inline uint64_t sm64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
static inline uint64_t rnd3(uint64_t seed, uint64_t a, uint64_t b) { return sm64(sm64(seed ^ (a * 0xD6E8FEB86659FD93ull)) ^ b); }
static const uint32_t SYN_CUM[20] = {5407, 6305, 9877, 14300, 16830, 21463, 22951, 26832, 30638, 36969,
                                     38548, 41209, 44309, 46884, 50510, 54855, 58361, 62858, 63572, 65535};
static const char SYN_AAS[] = "ACDEFGHIKLMNPQRSTVWY";
static const char SYN_TABLE1[] = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
#define SYN_K_TAXON 0x7461786F6Eull
#define SYN_K_VALUE 0x76616C7565ull
#define SYN_K_MUT 0x6D7574ull
#define SYN_K_CODON 0x636F646F6Eull
static inline uint32_t syn_residue(uint64_t seed, uint64_t j, uint64_t p) {
    const uint32_t u = (uint32_t)(rnd3(seed, j, p) & 0xFFFF);
    uint32_t c = 0;
    for (int i = 0; i < 19; ++i) c += u > SYN_CUM[i];
    return c;
}

typedef struct {
    uint64_t seed, j0, j1; uint32_t plen, home_pct, anc_pct, ntax;
    const uint32_t* depth; const uint32_t* parent_dense;
    uint64_t* out; /* (key45 << 16 | dense) per window */
} SynGen;
static void* syn_gen_worker(void* arg) {
    SynGen* g = (SynGen*)arg;
    const uint32_t wpp = g->plen - 8;
    uint8_t* res = (uint8_t*)malloc(g->plen);
    for (uint64_t j = g->j0; j < g->j1; ++j) {
        for (uint32_t p = 0; p < g->plen; ++p) res[p] = (uint8_t)syn_residue(g->seed, j, p);
        const uint32_t home = (uint32_t)(rnd3(g->seed ^ SYN_K_TAXON, j, 0) % g->ntax);
        uint64_t key = 0;
        for (uint32_t p = 0; p < 8; ++p) key = key << 5 | res[p];
        for (uint32_t p = 0; p < wpp; ++p) {
            key = (key << 5 | res[p + 8]) & ((1ull << 45) - 1);
            const uint64_t rv = rnd3(g->seed ^ SYN_K_VALUE, j, p);
            const uint32_t u = (uint32_t)(rv % 100);
            const uint64_t hi = rv >> 32;
            uint32_t dense = home;
            if (u >= g->home_pct && u < g->home_pct + g->anc_pct) {
                const uint32_t d = (uint32_t)(hi % (g->depth[home] + 1ull));
                while (g->depth[dense] > d) dense = g->parent_dense[dense];
            } else if (u >= g->home_pct + g->anc_pct) {
                dense = (uint32_t)(hi % g->ntax);
            }
            g->out[(j - g->j0) * wpp + p] = key << 16 | dense;
        }
    }
    free(res);
    return NULL;
}

static void radix_sort_u64(uint64_t* a, uint64_t* tmp, uint64_t n, int lo_bit, int hi_bit) {
    uint64_t *src = a, *dst = tmp;
    for (int b = lo_bit; b < hi_bit; b += 8) {
        uint64_t cnt[257] = {0};
        for (uint64_t i = 0; i < n; ++i) cnt[((src[i] >> b) & 255) + 1]++;
        for (int i = 0; i < 256; ++i) cnt[i + 1] += cnt[i];
        for (uint64_t i = 0; i < n; ++i) dst[cnt[(src[i] >> b) & 255]++] = src[i];
        uint64_t* t = src; src = dst; dst = t;
    }
    if (src != a) memcpy(a, src, n * sizeof(uint64_t));
}
typedef struct { uint64_t* a; uint64_t* tmp; const uint64_t* start; int* next; pthread_mutex_t* mu; } SynSort;
static void* syn_sort_worker(void* arg) {
    SynSort* s = (SynSort*)arg;
    for (;;) {
        pthread_mutex_lock(s->mu);
        const int b = (*s->next)++;
        pthread_mutex_unlock(s->mu);
        if (b >= 256) break;
        radix_sort_u64(s->a + s->start[b], s->tmp + s->start[b], s->start[b + 1] - s->start[b], 16, 53);
    }
    return NULL;
}

/* All 9-mer windows of the proteome, sorted, equal k-mers merged by LCA, as an fst Map image (malloc'ed; ref_free).
 * id_of / depth / parent_dense: the taxonomy in the library's preorder numbering (oracle/synth.py: Preorder). */
uint8_t* ref_synth_fst(uint64_t seed, uint64_t n_prot, uint32_t plen, uint32_t home_pct, uint32_t anc_pct,
                       const uint64_t* id_of, const uint32_t* depth, const uint32_t* parent_dense, uint32_t ntax,
                       int threads, uint64_t* out_size, uint64_t* out_nkeys) {
    *out_size = 0;
    if (plen < 9 || ntax == 0 || ntax > 65536 || n_prot == 0) return NULL;
    if (threads < 1) threads = 1;
    const uint64_t wpp = plen - 8, n = n_prot * wpp;
    uint64_t* a = (uint64_t*)malloc(n * sizeof(uint64_t));
    uint64_t* tmp = (uint64_t*)malloc(n * sizeof(uint64_t));
    if (!a || !tmp) { free(a); free(tmp); return NULL; }
    pthread_t* th = (pthread_t*)malloc(threads * sizeof(pthread_t));
    SynGen* gs = (SynGen*)calloc(threads, sizeof(SynGen));
    for (int t = 0; t < threads; ++t) {
        SynGen g = {seed, n_prot * t / threads, n_prot * (t + 1) / threads, plen, home_pct, anc_pct, ntax, depth, parent_dense, NULL};
        g.out = a + g.j0 * wpp;
        gs[t] = g;
        pthread_create(&th[t], NULL, syn_gen_worker, &gs[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
    free(gs);
    /* partition by the top 8 key bits, then sort the 256 parts on the threads */
    uint64_t start[257] = {0};
    for (uint64_t i = 0; i < n; ++i) start[(a[i] >> 53) + 1]++;
    for (int i = 0; i < 256; ++i) start[i + 1] += start[i];
    {
        uint64_t fill[256];
        memcpy(fill, start, sizeof fill);
        for (uint64_t i = 0; i < n; ++i) tmp[fill[a[i] >> 53]++] = a[i];
        uint64_t* t = a; a = tmp; tmp = t;
    }
    pthread_mutex_t mu;
    pthread_mutex_init(&mu, NULL);
    int next = 0;
    SynSort ss = {a, tmp, start, &next, &mu};
    for (int t = 0; t < threads; ++t) pthread_create(&th[t], NULL, syn_sort_worker, &ss);
    for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
    pthread_mutex_destroy(&mu);
    free(th);
    free(tmp);
    Builder b;
    memset(&b, 0, sizeof b);
    b.stack_cap = 64;
    b.stack = (BNode*)calloc(b.stack_cap, sizeof(BNode));
    b_put_int(&b, 2, 8);
    b_put_int(&b, 0, 8);
    uint64_t i = 0;
    while (i < n && !b.error) {
        const uint64_t key = a[i] >> 16;
        uint32_t x = (uint32_t)(a[i] & 0xFFFF);
        for (++i; i < n && (a[i] >> 16) == key; ++i) { /* equal k-mers: LCA of their values */
            uint32_t y = (uint32_t)(a[i] & 0xFFFF);
            while (x != y) { if (depth[x] >= depth[y]) x = parent_dense[x]; else y = parent_dense[y]; }
        }
        uint8_t kb[9];
        for (int c = 0; c < 9; ++c) kb[c] = (uint8_t)SYN_AAS[(key >> (5 * (8 - c))) & 31];
        builder_insert(&b, kb, 9, id_of[x]);
    }
    free(a);
    uint8_t* result = NULL;
    if (!b.error) {
        compile_from(&b, 0);
        const uint64_t root = emit_node(&b, &b.stack[0]);
        b_put_int(&b, b.nkeys, 8);
        b_put_int(&b, root, 8);
        result = b.buf;
        *out_size = b.len;
        if (out_nkeys) *out_nkeys = b.nkeys;
    } else {
        free(b.buf);
    }
    for (int k = 0; k < b.stack_cap; ++k) free(b.stack[k].tr);
    free(b.stack);
    free(b.prev);
    return result;
}

typedef struct {
    uint64_t seed, n_prot, read_seed, first_pair, r0, r1; uint32_t plen, read_len, hit_pct;
    uint8_t* out; const uint8_t (*codon_tab)[6]; const uint8_t* ncodon;
} SynReads;
static void* syn_reads_worker(void* arg) {
    SynReads* s = (SynReads*)arg;
    const uint32_t L = s->read_len, ncod = L / 3;
    static const uint8_t base_of_tcag[4] = {3, 1, 0, 2};
    for (uint64_t r = s->r0; r < s->r1; ++r) {
        const uint64_t pair = s->first_pair + r / 2, mate = r & 1, rid = pair * 2 + mate;
        const uint64_t rp = rnd3(s->read_seed, pair, 0);
        const int hit = (rp % 100) < s->hit_pct && s->plen >= ncod;
        const uint64_t j = (rp >> 8) % s->n_prot;
        const uint64_t rm = rnd3(s->read_seed, pair, 1 + mate);
        const uint64_t o = hit ? rm % (s->plen - ncod + 1) : 0;
        const uint32_t sh = (uint32_t)((rm >> 32) % 3);
        uint8_t* dst = s->out + r * L;
        for (uint32_t x = 0; x < L; ++x) {
            const uint64_t rx = rnd3(s->read_seed ^ SYN_K_MUT, rid, x);
            uint32_t base = (uint32_t)((rx >> 16) & 3);
            if (hit && x >= sh && (x - sh) / 3 < (L - sh) / 3 && ((rx & 0xFFFF) % 100) != 0) {
                const uint32_t q = (x - sh) / 3, ph = (x - sh) % 3;
                const uint32_t aa = syn_residue(s->seed, j, o + q);
                const uint64_t rc = rnd3(s->read_seed ^ SYN_K_CODON, rid, q);
                const uint32_t codon = s->codon_tab[aa][rc % s->ncodon[aa]];
                base = base_of_tcag[(codon >> (2 * (2 - ph))) & 3];
            }
            const int is_n = ((rx >> 32) % 1000) == 0;
            if (mate) dst[L - 1 - x] = is_n ? 'N' : (uint8_t)"ACGT"[3 - base];
            else dst[x] = is_n ? 'N' : (uint8_t)"ACGT"[base];
        }
    }
    return NULL;
}
/* npairs * 2 reads of read_len nucleotides (mate 2 reverse-complemented) into out. */
void ref_synth_reads(uint64_t seed, uint64_t n_prot, uint32_t plen, uint64_t read_seed, uint64_t first_pair, uint64_t npairs,
                     uint32_t read_len, uint32_t hit_pct, uint8_t* out, int threads) {
    uint8_t codon_tab[20][6];
    uint8_t ncodon[20];
    memset(codon_tab, 0, sizeof codon_tab);
    for (int a = 0; a < 20; ++a) {
        ncodon[a] = 0;
        for (int c = 0; c < 64; ++c) if (SYN_TABLE1[c] == SYN_AAS[a]) codon_tab[a][ncodon[a]++] = (uint8_t)c;
    }
    if (threads < 1) threads = 1;
    pthread_t* th = (pthread_t*)malloc(threads * sizeof(pthread_t));
    SynReads* ss = (SynReads*)calloc(threads, sizeof(SynReads));
    const uint64_t nreads = 2 * npairs;
    for (int t = 0; t < threads; ++t) {
        SynReads s = {seed, n_prot, read_seed, first_pair, nreads * t / threads, nreads * (t + 1) / threads, plen, read_len, hit_pct,
                      out, (const uint8_t (*)[6])codon_tab, ncodon};
        ss[t] = s;
        pthread_create(&th[t], NULL, syn_reads_worker, &ss[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
    free(ss);
    free(th);
}
