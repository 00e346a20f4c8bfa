// oracle/ref_dump.cpp — TEST INFRASTRUCTURE (oracle harness), not product code.
//
// Includes the UNMODIFIED reference headers from /root/reference/src (never copied into this repo)
// and dumps what the reference computes, in a machine-readable form, so that oracle/qs_oracle.c
// (the restatement) and the CUDA path can be pinned against the reference itself:
//
//   qs_ref_dump <ref.nwk> <eval.nwk> <out_prefix> [savemem(0|1)] [threads]
//
// writes
//   <out_prefix>.counts.u64  C(n,4) x 3 little-endian uint64, quartets a<b<c<d (lookup ids = position
//                            in the reference tree's Euler-tour leaf order) in QuartetLookupTable rank
//                            order (quartet_lookup_table.hpp:141-212); entry = counts of
//                            (ab|cd, ac|bd, ad|bc) as returned by
//                            QuartetCounterLookup::countQuartetOccurrences (QuartetCounterLookup.hpp:300-318)
//   <out_prefix>.scores.f64  3 x edge_count little-endian doubles: LQIC, QPIC, EQPIC by genesis edge index
//                            (QuartetScoreComputer.hpp:106-126); QPIC/EQPIC are +inf-filled when the
//                            reference tree is multifurcating (the reference leaves them empty)
//   <out_prefix>.meta.txt    n, m, edge_count, bifurcating flag, leaf names in lookup-id order, and per
//                            node: index, parent index, edge index above it, name
#include "genesis/genesis.hpp"
#include "QuartetScoreComputer.hpp"
#include <cstdio>
#include <fstream>
#include <iostream>
#include <limits>
#ifdef GENESIS_OPENMP
#include <omp.h>
#endif

using namespace genesis;
using namespace tree;

static size_t count_trees(const std::string& p) {
    size_t c = 0;
    utils::InputStream in(utils::make_unique<utils::FileInputSource>(p));
    auto it = NewickInputIterator(in);
    while (it) { ++c; ++it; }
    return c;
}

template <typename CINT>
static void run(Tree const& ref, std::string const& evalPath, size_t m, bool savemem, std::string const& prefix) {
    // leaves in Euler order -> lookup ids
    std::vector<size_t> leaves;
    for (auto it : eulertour(ref)) if (it.node().is_leaf()) leaves.push_back(it.node().index());
    size_t n = leaves.size();

    {
        QuartetCounterLookup<CINT> qcl(ref, evalPath, m, savemem);
        FILE* f = fopen((prefix + ".counts.u64").c_str(), "wb");
        for (size_t d = 3; d < n; ++d) for (size_t c = 2; c < d; ++c) for (size_t b = 1; b < c; ++b) for (size_t a = 0; a < b; ++a) {
            auto t = qcl.countQuartetOccurrences(leaves[a], leaves[b], leaves[c], leaves[d]);
            uint64_t v[3] = {(uint64_t)std::get<0>(t), (uint64_t)std::get<1>(t), (uint64_t)std::get<2>(t)};
            fwrite(v, 8, 3, f);
        }
        fclose(f);
    }
    QuartetScoreComputer<CINT> qsc(ref, evalPath, m, false, savemem);
    std::vector<double> lq = qsc.getLQICScores(), qp = qsc.getQPICScores(), eqp = qsc.getEQPICScores();
    size_t E = ref.edge_count();
    double inf = std::numeric_limits<double>::infinity();
    if (qp.empty()) qp.assign(E, inf);
    if (eqp.empty()) eqp.assign(E, inf);
    FILE* f = fopen((prefix + ".scores.f64").c_str(), "wb");
    fwrite(lq.data(), 8, E, f); fwrite(qp.data(), 8, E, f); fwrite(eqp.data(), 8, E, f);
    fclose(f);
    qsc.printRawQICScores(ref, prefix + ".rawqic.txt");

    std::ofstream meta(prefix + ".meta.txt");
    meta << "n " << n << "\nm " << m << "\nedges " << E << "\nbifurcating " << (is_bifurcating(ref) ? 1 : 0) << "\nleaves";
    for (size_t i = 0; i < n; ++i) meta << " " << ref.node_at(leaves[i]).data<DefaultNodeData>().name;
    meta << "\n";
    for (size_t i = 0; i < ref.node_count(); ++i) {
        auto const& nd = ref.node_at(i);
        long parent = -1, edge = -1;
        if (!nd.is_root()) { parent = (long)nd.primary_link().outer().node().index(); edge = (long)nd.primary_link().edge().index(); }
        meta << "node " << i << " " << parent << " " << edge << " " << (nd.is_leaf() ? 1 : 0) << " " << nd.data<DefaultNodeData>().name << "\n";
    }
}

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: qs_ref_dump ref.nwk eval.nwk out_prefix [savemem] [threads]\n"); return 2; }
    bool savemem = argc > 4 && atoi(argv[4]) != 0;
#ifdef GENESIS_OPENMP
    if (argc > 5 && atoi(argv[5]) > 0) omp_set_num_threads(atoi(argv[5]));
#endif
    Tree ref = DefaultTreeNewickReader().from_file(argv[1]);
    size_t m = count_trees(argv[2]);
    // same CINT dispatch as the reference main (QuartetScores.cpp:115-147)
    if (m < (size_t(1) << 8)) run<uint8_t>(ref, argv[2], m, savemem, argv[3]);
    else if (m < (size_t(1) << 16)) run<uint16_t>(ref, argv[2], m, savemem, argv[3]);
    else run<uint32_t>(ref, argv[2], m, savemem, argv[3]);
    return 0;
}
