// QuartetScoreComputerB200.hpp — drop-in replacement of the reference's QuartetScoreComputer<CINT>
// (src/QuartetScoreComputer.hpp:43-51, :698-785) on top of libqscuda's C ABI (include/qscuda.h).
//
// Same template name, constructor arguments and methods as the reference class, so the reference's own
// main (src/QuartetScores.cpp:114-147) compiles against it unchanged — see integration/main_b200.cpp and
// INTEGRATION.md.  Host side only: genesis still parses the REFERENCE tree and numbers its nodes/edges (the annotated-Newick writer
// depends on that numbering); the evaluation trees go through libqscuda's parallel single-pass parser.  All counting and scoring happens
// in libqscuda.so; there is no CPU fallback (a missing GPU is a std::runtime_error).
//
// Multi-GPU: one process, one context per device ($QS_NUM_GPUS, default 1; or the ordinals in $QS_DEVICES), each owning a shard of the quartet
// rank space and driven by its own host thread; the per-shard partials are reduced on the host (min for
// LQ-IC, sum for the pair sums) — no collective is needed inside a single process.
#pragma once

#include "genesis/genesis.hpp"
#include "qscuda.h"

#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <iostream>
#include <limits>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace qsb200 {

using namespace genesis;
using namespace genesis::tree;

struct FlatTree {
    std::vector<int32_t> parent, parent_edge, leaf_id, first_child, next_sibling;
};

// genesis numbers nodes parent-first (root = 0), so parent[i] < i holds
inline void flatten_tree(Tree const& tree, FlatTree& f, bool want_children,
                         std::unordered_map<std::string, int32_t> const* name_to_id, std::vector<int32_t> const* node_to_id) {
    const size_t N = tree.node_count();
    f.parent.assign(N, -1); f.parent_edge.assign(N, -1); f.leaf_id.assign(N, -1);
    if (want_children) { f.first_child.assign(N, -1); f.next_sibling.assign(N, -1); }
    if (tree.root_node().index() != 0) throw std::runtime_error("flatten_tree: root node index is not 0");
    for (size_t i = 0; i < N; ++i) {
        TreeNode const& node = tree.node_at(i);
        const bool is_root = (i == 0);
        if (!is_root) {
            f.parent[i] = (int32_t)node.primary_link().outer().node().index();
            f.parent_edge[i] = (int32_t)node.primary_link().edge().index();
            if (f.parent[i] >= (int32_t)i) throw std::runtime_error("flatten_tree: node numbering is not parent-first");
        }
        if (node.is_leaf()) {
            if (node_to_id) f.leaf_id[i] = (*node_to_id)[i];
            else f.leaf_id[i] = name_to_id->at(node.data<DefaultNodeData>().name);   // std::out_of_range on an unknown taxon, as QuartetCounterLookup.hpp:218
        }
        if (want_children && !node.is_leaf()) {
            // children in Newick order: the links after the primary link; the root's primary link is its first child
            TreeLink const* first = is_root ? &tree.root_link() : &node.primary_link().next();
            TreeLink const* stop = is_root ? &tree.root_link() : &node.primary_link();
            int32_t prev = -1;
            TreeLink const* l = first;
            do {
                const int32_t ch = (int32_t)l->outer().node().index();
                if (prev < 0) f.first_child[i] = ch; else f.next_sibling[prev] = ch;
                prev = ch;
                l = &l->next();
            } while (l != stop);
        }
    }
}

inline void qs_check(qs_ctx* ctx, int rc, const char* what) {
    if (rc == QS_OK) return;
    std::string msg = std::string(what) + ": " + (ctx ? qs_last_error(ctx) : qs_strerror(rc));
    if (rc == QS_E_MEMORY) throw std::runtime_error("Insufficient memory! " + msg);      // wording of QuartetScoreComputer.hpp:735-737
    throw std::runtime_error(msg);
}

template<typename CINT>
class QuartetScoreComputer {
public:
    QuartetScoreComputer(Tree const& refTree, const std::string& evalTreesPath, size_t m, bool verboseOutput, bool savemem);
    ~QuartetScoreComputer() { for (auto c : ctxs) if (c) qs_destroy(c); }
    QuartetScoreComputer(QuartetScoreComputer const&) = delete;
    QuartetScoreComputer& operator=(QuartetScoreComputer const&) = delete;

    std::vector<double> getLQICScores() { return LQICScores; }
    std::vector<double> getQPICScores() { return QPICScores; }
    std::vector<double> getEQPICScores() { return EQPICScores; }
    void printRawQICScores(Tree const& refTree, const std::string& rawFilePath);

private:
    std::vector<qs_ctx*> ctxs;
    std::vector<std::string> taxa;     // label of every lookup id
    std::vector<double> LQICScores, QPICScores, EQPICScores;
    int count_scale = 1;
    bool verbose = false;
};

template<typename CINT>
QuartetScoreComputer<CINT>::QuartetScoreComputer(Tree const& refTree, const std::string& evalTreesPath, size_t m, bool verboseOutput, bool savemem)
    : verbose(verboseOutput) {
    (void)m;
    std::cout << "There are " << m << " evaluation trees.\n";
    // taxon (lookup) ids: position in the reference tree's Euler-tour leaf order (QuartetCounterLookup.hpp:249-258)
    std::vector<int32_t> refNodeToId(refTree.node_count(), -1);
    std::unordered_map<std::string, int32_t> nameToId;
    for (auto it : eulertour(refTree)) {
        if (it.node().is_leaf() && refNodeToId[it.node().index()] < 0) {
            const int32_t id = (int32_t)taxa.size();
            refNodeToId[it.node().index()] = id;
            nameToId[it.node().template data<DefaultNodeData>().name] = id;
            taxa.push_back(it.node().template data<DefaultNodeData>().name);
        }
    }
    const int n = (int)taxa.size();
    std::cout << "The reference tree has " << n << " taxa.\n";
    FlatTree ref;
    flatten_tree(refTree, ref, true, nullptr, &refNodeToId);

    // The reference's memory policy (QuartetScoreComputer.hpp:724-745) decides two different things; both are kept.
    //  (1) WHICH COUNTS ARE SCORED.  With -s, or when its n^4 table would exceed 0.9 x the host's RAM, the reference counts into
    //      the memory-efficient table, whose entries are doubled and wrap in CINT — visible in QP-IC /
    //      EQP-IC once the 32-bit accumulators wrap (App. B4).  The same test on the same host gives the same choice here
    //      (count_scale 2), and the same stdout lines, so outputs match the reference binary run on this machine.
    //  (2) WHERE THE TABLE LIVES.  Here that is a question of HBM, not host RAM: -s asks for no resident table
    //      (QS_MODE_TABLE_FREE); otherwise QS_MODE_AUTO keeps each shard's table on its GPU if it fits and falls back to
    //      table-free slabs if not — the reference's "Insufficient memory!" (:735-737) is only raised when even that fails.
    const size_t nn = (size_t)n;
    const size_t memoryLookupFast = nn * nn * nn * nn * sizeof(CINT);
    const size_t memoryLookup = (nn * (nn - 1) * (nn - 2) * (nn - 3) / 24) * 3 * sizeof(CINT) + sizeof(size_t);
    const size_t estimatedMemory = (size_t)sysconf(_SC_PHYS_PAGES) * (size_t)sysconf(_SC_PAGE_SIZE);       // getTotalSystemMemory (:610-621)
    std::cout << "Estimated memory usages (in bytes):" << std::endl;
    std::cout << "  Runtime-efficient Lookup table: " << memoryLookupFast << std::endl;
    std::cout << "  Memory-efficient Lookup table: " << memoryLookup << std::endl;
    std::cout << "  Estimated available memory: " << estimatedMemory << std::endl;
    const bool compact = savemem || memoryLookupFast > 0.9 * estimatedMemory;
    std::cout << (compact ? "Using memory-efficient Lookup table\n" : "Using runtime-efficient Lookup table\n");
    count_scale = compact ? 2 : 1;

    // one context per GPU
    int n_gpus = 1;
    if (const char* env = std::getenv("QS_NUM_GPUS")) n_gpus = std::max(1, std::atoi(env));
    const int mode = savemem ? QS_MODE_TABLE_FREE : QS_MODE_AUTO;
    // shard g runs on device g, or on the g-th entry of $QS_DEVICES (comma-separated ordinals; several shards may share a GPU)
    std::vector<int> devices;
    if (const char* env = std::getenv("QS_DEVICES")) {
        for (const char* p = env; *p;) { devices.push_back(std::atoi(p)); while (*p && *p != ',') ++p; if (*p == ',') ++p; }
        if (!devices.empty()) n_gpus = (int)devices.size();
    }
    ctxs.assign(n_gpus, nullptr);
    for (int g = 0; g < n_gpus; ++g) {
        qs_check(nullptr, qs_create(&ctxs[g], n, (int)sizeof(CINT), mode, devices.empty() ? g : devices[g], g, n_gpus), "qs_create");
        qs_check(ctxs[g], qs_set_count_scale(ctxs[g], count_scale), "qs_set_count_scale");
        qs_check(ctxs[g], qs_set_reference(ctxs[g], (int)refTree.node_count(), ref.parent.data(), ref.parent_edge.data(), ref.leaf_id.data(),
                                           ref.first_child.data(), ref.next_sibling.data()), "qs_set_reference");
    }

    // evaluation trees: one multi-threaded pass from the Newick text to the flat encoding (libqscuda's own parser,
    // qs_newick_flatten) instead of a second serial genesis parse (QuartetCounterLookup.hpp:202-221); an unknown taxon
    // is reported where the reference throws std::out_of_range (:218)
    std::chrono::steady_clock::time_point begin = std::chrono::steady_clock::now();
    {
        std::string text = utils::file_read(evalTreesPath);
        std::vector<const char*> names;
        for (auto const& t : taxa) names.push_back(t.c_str());
        qs_flat_trees* flat = nullptr;
        char err[512] = {0};
        int threads = 0;
        if (const char* env = std::getenv("OMP_NUM_THREADS")) threads = std::atoi(env);
        const int rc = qs_newick_flatten(text.data(), text.size(), n, names.data(), threads, &flat, err, sizeof err);
        if (rc != QS_OK) {
            if (std::string(err).find("is not in the reference tree") != std::string::npos) throw std::out_of_range(std::string("unordered_map::at: ") + err);
            throw std::runtime_error(std::string("evaluation trees: ") + err);
        }
        int64_t T = 0, N = 0;
        const int64_t* off = nullptr; const int32_t *par = nullptr, *leaf = nullptr;
        qs_flat_trees_view(flat, &T, &N, &off, &par, &leaf);
        for (auto c : ctxs) {
            const int r = qs_add_trees(c, (int)T, off, par, leaf);
            if (r != QS_OK) { qs_flat_trees_free(flat); qs_check(c, r, "qs_add_trees"); }
        }
        qs_flat_trees_free(flat);
    }
    for (auto c : ctxs) qs_check(c, qs_rebalance_shards(c, nullptr), "qs_rebalance_shards");      // shard ranges for the class mix of these trees
    std::cout << "Finished parsing evaluation trees.\n";

    // count + partial scores per shard (one host thread per GPU), reduce on the host
    int64_t n_pairs = 0;
    qs_check(ctxs[0], qs_score_num_pairs(ctxs[0], &n_pairs), "qs_score_num_pairs");
    const size_t E = refTree.edge_count();
    std::vector<std::vector<double>> lq(n_gpus, std::vector<double>(E));
    std::vector<std::vector<uint64_t>> sums(n_gpus, std::vector<uint64_t>((size_t)n_pairs * 3));
    std::vector<std::string> errors(n_gpus);
    std::vector<std::thread> workers;
    for (int g = 0; g < n_gpus; ++g)
        workers.emplace_back([&, g]() {
            try {
                qs_check(ctxs[g], qs_count(ctxs[g]), "qs_count");
                qs_check(ctxs[g], qs_score_partials(ctxs[g], count_scale, lq[g].data(), sums[g].data()), "qs_score_partials");
            } catch (std::exception const& e) { errors[g] = e.what(); }
        });
    for (auto& w : workers) w.join();
    for (auto& e : errors) if (!e.empty()) throw std::runtime_error(e);
    std::chrono::steady_clock::time_point end = std::chrono::steady_clock::now();
    std::cout << "Finished counting quartets.\nIt took: " << std::chrono::duration_cast<std::chrono::microseconds>(end - begin).count()
              << " microseconds.\n";
    begin = std::chrono::steady_clock::now();
    for (int g = 1; g < n_gpus; ++g) {
        for (size_t e = 0; e < E; ++e) lq[0][e] = std::min(lq[0][e], lq[g][e]);
        for (size_t k = 0; k < sums[0].size(); ++k) sums[0][k] += sums[g][k];
    }
    LQICScores.assign(E, std::numeric_limits<double>::infinity());
    QPICScores.assign(E, std::numeric_limits<double>::infinity());
    EQPICScores.assign(E, std::numeric_limits<double>::infinity());
    qs_check(ctxs[0], qs_score_finalize(ctxs[0], 0, lq[0].data(), sums[0].data(), LQICScores.data(), QPICScores.data(), EQPICScores.data()), "qs_score_finalize");
    if (!is_bifurcating(refTree)) {       // QuartetScoreComputer.hpp:760-765: only LQ-IC for a multifurcating reference
        std::cout << "The reference tree is multifurcating.\n";
        QPICScores.clear();
        EQPICScores.clear();
    } else {
        std::cout << "The reference tree is bifurcating.\n";
    }
    end = std::chrono::steady_clock::now();
    std::cout << "Finished computing scores.\nIt took: " << std::chrono::duration_cast<std::chrono::microseconds>(end - begin).count()
              << " microseconds.\n";
}

template<typename CINT>
void QuartetScoreComputer<CINT>::printRawQICScores(Tree const& refTree, const std::string& rawFilePath) {
    (void)refTree;
    std::vector<const char*> names;
    for (auto const& t : taxa) names.push_back(t.c_str());
    // any table type and any number of shards, as the reference (QuartetScores.cpp:120-122 calls it regardless of -s)
    qs_check(ctxs[0], qs_write_raw_qic_shards(ctxs.data(), (int)ctxs.size(), count_scale, names.data(), rawFilePath.c_str()), "qs_write_raw_qic_shards");
}

}  // namespace qsb200
