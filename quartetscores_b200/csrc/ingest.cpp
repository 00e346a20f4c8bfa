// ingest.cpp — single-pass, multi-threaded Newick -> flat-tree ingest (host side of libqscuda).
//
// Replaces, for the EVALUATION trees only, the two serial genesis parses of the reference
// (src/QuartetScores.cpp:23-32 countEvalTrees and src/QuartetCounterLookup.hpp:202-221 countQuartets, which
// re-reads the file through NewickInputIterator) by one pass that goes straight from the text to the
// qs_add_trees encoding (include/qscuda.h): pre-order node numbering, children in Newick order, parent[i] < i,
// leaf_lookup_id >= 0 exactly for leaves.  No tree objects are built.
//
// Grammar: the subset of genesis' lexer that evaluation trees use
// (genesis/lib/genesis/tree/formats/newick/reader.cpp:296-387): '(' ')' ',' ';', names = printable characters
// except blanks and ":;()[],", quoted names in '...' or "..." with a doubled quote standing for the quote itself,
// [comments] (dropped), ":<number>" branch lengths (dropped), inner-node labels (dropped).  A taxon name that is not
// in the reference tree is an error naming the taxon — the reference throws std::out_of_range at the same place
// (QuartetCounterLookup.hpp:218).
//
// Threads: the text is cut into tree spans at top-level ';' by one sequential scan (quotes and comments
// respected), the spans are parsed by n_threads workers into thread-local arrays, and the pieces are concatenated
// in file order, so the result does not depend on the thread count.
#include "../../include/qscuda.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <string_view>
#include <thread>
#include <vector>

struct qs_flat_trees {
    std::vector<int64_t> node_offsets{0};
    std::vector<int32_t> parent, leaf_lookup_id;
};

namespace {

struct Span { size_t begin, end; };   // [begin, end): one tree WITHOUT its ';'

inline bool is_space(unsigned char c) { return c == ' ' || (c >= 9 && c <= 13); }
inline bool is_name_char(unsigned char c) {
    return c > 32 && c < 127 && c != ':' && c != ';' && c != '(' && c != ')' && c != '[' && c != ']' && c != ',';
}

// cut the text into trees; returns false on an unterminated quote/comment/tree
bool split_trees(const char* s, size_t n, std::vector<Span>& out, std::string& err) {
    size_t i = 0, start = 0;
    bool content = false;
    if (n && !memchr(s, '[', n) && !memchr(s, '\'', n) && !memchr(s, '"', n)) {
        // common case (no comments, no quoted labels): every ';' ends a tree, found at memchr speed
        while (start < n) {
            const char* p = (const char*)memchr(s + start, ';', n - start);
            const size_t end = p ? (size_t)(p - s) : n;
            size_t b = start;
            while (b < end && is_space((unsigned char)s[b])) ++b;
            if (b < end) {
                if (!p) { err = "tree not terminated by ';'"; return false; }
                out.push_back(Span{b, end});
            }
            start = end + 1;
        }
        return true;
    }
    while (i < n) {
        const unsigned char c = (unsigned char)s[i];
        if (c == ';') {
            if (content) out.push_back(Span{start, i});
            start = i + 1; content = false; ++i;
        } else if (c == '[') {
            const void* p = memchr(s + i, ']', n - i);
            if (!p) { err = "unterminated comment"; return false; }
            i = (size_t)((const char*)p - s) + 1;
        } else if (c == '\'' || c == '"') {
            content = true;
            size_t j = i + 1;
            while (true) {
                const void* p = memchr(s + j, c, n - j);
                if (!p) { err = "unterminated quoted label"; return false; }
                j = (size_t)((const char*)p - s);
                if (j + 1 < n && (unsigned char)s[j + 1] == c) j += 2; else break;
            }
            i = j + 1;
        } else {
            if (!is_space(c)) content = true;
            ++i;
        }
    }
    if (content) { err = "tree not terminated by ';'"; return false; }
    return true;
}

struct Piece {
    std::vector<int64_t> sizes;            // nodes per tree
    std::vector<int32_t> parent, leaf;
    std::string err;
    size_t err_tree = 0;
};

// taxon name -> lookup id: open addressing, power-of-two size, 8-bytes-at-a-time multiplicative hash.  Read-only
// while the workers run (one lookup per leaf is the parser's hot spot: ~1e6 lookups for 10,000 x 100-taxon trees).
struct NameMap {
    struct Slot { const char* p; uint32_t len; int32_t id; uint64_t h; };
    std::vector<Slot> slots;
    uint64_t mask = 0;
    static uint64_t hash(const char* p, size_t n) {
        uint64_t h = 0x9E3779B97F4A7C15ull ^ (n * 0xff51afd7ed558ccdull);
        while (n >= 8) { uint64_t w; memcpy(&w, p, 8); h = (h ^ w) * 0xff51afd7ed558ccdull; h ^= h >> 32; p += 8; n -= 8; }
        uint64_t w = 0;
        memcpy(&w, p, n);
        h = (h ^ w) * 0xc4ceb9fe1a85ec53ull;
        return h ^ (h >> 29);
    }
    void init(size_t n) {
        size_t cap = 16;
        while (cap < 4 * n) cap <<= 1;
        slots.assign(cap, Slot{nullptr, 0, -1, 0});
        mask = cap - 1;
    }
    bool insert(const char* p, int32_t id) {             // false: duplicate
        const size_t n = strlen(p);
        const uint64_t h = hash(p, n);
        for (uint64_t i = h & mask;; i = (i + 1) & mask) {
            Slot& s = slots[i];
            if (!s.p) { s = Slot{p, (uint32_t)n, id, h}; return true; }
            if (s.h == h && s.len == n && memcmp(s.p, p, n) == 0) return false;
        }
    }
    int32_t find(std::string_view k) const {             // -1: unknown
        const uint64_t h = hash(k.data(), k.size());
        for (uint64_t i = h & mask;; i = (i + 1) & mask) {
            const Slot& s = slots[i];
            if (!s.p) return -1;
            if (s.h == h && s.len == k.size() && memcmp(s.p, k.data(), k.size()) == 0) return s.id;
        }
    }
};

// parse one tree span, appending its nodes to P; false (P.err set) on a syntax error or unknown taxon
bool parse_tree(const char* s, Span sp, const NameMap& names, std::vector<int32_t>& stack, std::string& scratch, Piece& P) {
    const size_t base = P.parent.size();
    stack.clear();
    // cur_kind: what the last completed element was — 0 nothing pending (a leaf may still have to be created),
    // 1 a leaf/labelled element, 2 a closed inner node (a label may follow)
    int pending = 0;
    bool have_root = false, named = false;
    auto new_node = [&](int32_t leaf_id) {
        const int32_t idx = (int32_t)(P.parent.size() - base);
        P.parent.push_back(stack.empty() ? -1 : stack.back());
        P.leaf.push_back(leaf_id);
        return idx;
    };
    auto leaf_from_name = [&](std::string_view nm) -> bool {
        const int32_t id = names.find(nm);
        if (id < 0) { P.err = "taxon '" + std::string(nm) + "' of an evaluation tree is not in the reference tree"; return false; }
        if (stack.empty()) {
            if (have_root) { P.err = "more than one element at the top level of a tree"; return false; }
            have_root = true;
        }
        new_node(id);
        return true;
    };
    size_t i = sp.begin;
    while (i < sp.end) {
        const unsigned char c = (unsigned char)s[i];
        if (is_space(c)) { ++i; continue; }
        if (c == '[') { i = (size_t)((const char*)memchr(s + i, ']', sp.end - i) - s) + 1; continue; }
        if (c == '(') {
            if (pending != 0) { P.err = "'(' directly after a node"; return false; }
            if (stack.empty()) {
                if (have_root) { P.err = "more than one element at the top level of a tree"; return false; }
                have_root = true;
            }
            stack.push_back(new_node(-1));
            ++i;
        } else if (c == ',') {
            if (stack.empty()) { P.err = "',' outside of a tree"; return false; }
            if (pending == 0 && !leaf_from_name(std::string_view())) return false;      // empty leaf name
            pending = 0; named = false; ++i;
        } else if (c == ')') {
            if (stack.empty()) { P.err = "unbalanced ')'"; return false; }
            if (pending == 0 && !leaf_from_name(std::string_view())) return false;
            stack.pop_back();
            pending = 2; named = false; ++i;
        } else if (c == ':') {
            if (pending == 0) { if (!leaf_from_name(std::string_view())) return false; pending = 1; }
            ++i;
            while (i < sp.end && is_space((unsigned char)s[i])) ++i;
            while (i < sp.end && (is_name_char((unsigned char)s[i]))) ++i;                  // the length text is dropped
        } else if (c == '\'' || c == '"') {
            scratch.clear();
            size_t j = i + 1;
            while (true) {
                const char* p = (const char*)memchr(s + j, c, sp.end - j);
                scratch.append(s + j, (size_t)(p - (s + j)));
                j = (size_t)(p - s);
                if (j + 1 < sp.end && (unsigned char)s[j + 1] == c) { scratch.push_back((char)c); j += 2; } else break;
            }
            i = j + 1;
            if (pending == 2 && !named) named = true;                                     // inner-node label: dropped
            else if (pending == 0) { if (!leaf_from_name(scratch)) return false; pending = 1; named = true; }
            else { P.err = "two labels on one node"; return false; }
        } else if (is_name_char(c)) {
            size_t j = i;
            while (j < sp.end && is_name_char((unsigned char)s[j])) ++j;
            if (pending == 2 && !named) named = true;
            else if (pending == 0) { if (!leaf_from_name(std::string_view(s + i, j - i))) return false; pending = 1; named = true; }
            else { P.err = "two labels on one node"; return false; }
            i = j;
        } else {
            char b[64]; snprintf(b, sizeof b, "invalid character 0x%02x", (unsigned)c);
            P.err = b; return false;
        }
    }
    if (!stack.empty()) { P.err = "unbalanced '(' at ';'"; return false; }
    if (!have_root) { P.err = "empty tree"; return false; }
    P.sizes.push_back((int64_t)(P.parent.size() - base));
    return true;
}

void set_err(char* errbuf, size_t cap, const std::string& m) {
    if (errbuf && cap) { snprintf(errbuf, cap, "%s", m.c_str()); }
}

}  // namespace

extern "C" int qs_newick_flatten(const char* text, size_t text_len, int n_taxa, const char* const* taxon_names, int n_threads,
                                 qs_flat_trees** out, char* errbuf, size_t errbuf_len) {
    if (!out) return QS_E_ARG;
    *out = nullptr;
    if ((!text && text_len) || n_taxa < 0 || (n_taxa && !taxon_names)) { set_err(errbuf, errbuf_len, "bad argument"); return QS_E_ARG; }
    NameMap names;
    names.init((size_t)n_taxa);
    for (int i = 0; i < n_taxa; ++i) {
        if (!taxon_names[i]) { set_err(errbuf, errbuf_len, "null taxon name"); return QS_E_ARG; }
        if (!names.insert(taxon_names[i], i)) { set_err(errbuf, errbuf_len, std::string("duplicate taxon name '") + taxon_names[i] + "'"); return QS_E_ARG; }
    }
    std::vector<Span> spans;
    std::string err;
    if (!split_trees(text, text_len, spans, err)) { set_err(errbuf, errbuf_len, err); return QS_E_TREE; }
    const size_t T = spans.size();
    int nt = n_threads > 0 ? n_threads : (int)std::max(1u, std::thread::hardware_concurrency());
    nt = (int)std::max<size_t>(1, std::min<size_t>((size_t)nt, (T + 63) / 64));
    std::vector<Piece> pieces((size_t)nt);
    auto work = [&](int w) {
        const size_t b = T * (size_t)w / (size_t)nt, e = T * (size_t)(w + 1) / (size_t)nt;
        Piece& P = pieces[(size_t)w];
        size_t bytes = 0;
        for (size_t t = b; t < e; ++t) bytes += spans[t].end - spans[t].begin;
        P.parent.reserve(bytes / 4 + 16); P.leaf.reserve(bytes / 4 + 16); P.sizes.reserve(e - b);
        std::vector<int32_t> stack;
        std::string scratch;
        for (size_t t = b; t < e; ++t)
            if (!parse_tree(text, spans[t], names, stack, scratch, P)) { P.err_tree = t; return; }
    };
    if (nt == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int w = 0; w < nt; ++w) th.emplace_back(work, w);
        for (auto& t : th) t.join();
    }
    for (auto& P : pieces)
        if (!P.err.empty()) {                                   // the first failing tree in file order (pieces are in file order)
            set_err(errbuf, errbuf_len, "evaluation tree " + std::to_string(P.err_tree) + ": " + P.err);
            return QS_E_TREE;
        }
    auto* F = new qs_flat_trees();
    std::vector<size_t> node_base((size_t)nt + 1, 0), tree_base((size_t)nt + 1, 0);
    for (int w = 0; w < nt; ++w) { node_base[w + 1] = node_base[w] + pieces[w].parent.size(); tree_base[w + 1] = tree_base[w] + pieces[w].sizes.size(); }
    const size_t nodes = node_base[(size_t)nt];
    F->node_offsets.resize(T + 1); F->parent.resize(nodes); F->leaf_lookup_id.resize(nodes);
    F->node_offsets[0] = 0;
    auto gather = [&](int w) {                                  // every worker copies its own piece to its final place
        const Piece& P = pieces[(size_t)w];
        int64_t off = (int64_t)node_base[w];
        for (size_t k = 0; k < P.sizes.size(); ++k) { off += P.sizes[k]; F->node_offsets[tree_base[w] + k + 1] = off; }
        if (!P.parent.empty()) {
            memcpy(F->parent.data() + node_base[w], P.parent.data(), P.parent.size() * sizeof(int32_t));
            memcpy(F->leaf_lookup_id.data() + node_base[w], P.leaf.data(), P.leaf.size() * sizeof(int32_t));
        }
    };
    if (nt == 1) gather(0);
    else {
        std::vector<std::thread> th;
        for (int w = 0; w < nt; ++w) th.emplace_back(gather, w);
        for (auto& t : th) t.join();
    }
    *out = F;
    return QS_OK;
}

extern "C" int qs_flat_trees_view(const qs_flat_trees* f, int64_t* n_trees, int64_t* n_nodes, const int64_t** node_offsets,
                                  const int32_t** parent, const int32_t** leaf_lookup_id) {
    if (!f) return QS_E_ARG;
    if (n_trees) *n_trees = (int64_t)f->node_offsets.size() - 1;
    if (n_nodes) *n_nodes = (int64_t)f->parent.size();
    if (node_offsets) *node_offsets = f->node_offsets.data();
    if (parent) *parent = f->parent.data();
    if (leaf_lookup_id) *leaf_lookup_id = f->leaf_lookup_id.data();
    return QS_OK;
}

extern "C" void qs_flat_trees_free(qs_flat_trees* f) { delete f; }
