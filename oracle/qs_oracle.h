/* oracle/qs_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the reference QuartetScores algorithm for the hot path
 * (quartet counting -> lookup table -> LQ-IC / QP-IC / EQP-IC).  It exists to CHECK the CUDA
 * path; nothing in the product (quartetscores_b200/, libqscuda.so) may include, link or call it.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg use it.
 *
 * Parity pin: this restatement is checked in tests/test_oracle.py against
 *   (1) the known-answer vectors the survey generated from the reference binary,
 *   (2) fixtures under tests/golden/ produced by oracle/_ref/qs_ref_dump, a harness that includes the
 *       UNMODIFIED reference headers (see oracle/ref_dump.cpp, tests/golden/make_golden.py).
 * The reference repository itself ships no tests for this path.
 *
 * Every function cites the reference file:line (relative to /root/reference) it follows.
 *
 * Tree encoding (same as the C ABI in include/qscuda.h): nodes 0..N-1, parent[i] < i for i > 0,
 * parent[0] = -1 (root), leaf_id[i] = lookup id of a leaf (position of the taxon in the reference
 * tree's Euler-tour leaf order, QuartetCounterLookup.hpp:249-258) or -1 for an inner node.
 * Child order (only relevant for the reference tree) = increasing `child_rank[i]`.
 */
#ifndef QS_ORACLE_H
#define QS_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QSO_MISSING 0xFFFFu

/* C(n,4) */
uint64_t qso_num_quartets(uint64_t n);
/* quartet_lookup_table.hpp:141-168,170-212: rank of {a,b,c,d} (any order) */
uint64_t qso_rank(uint64_t a, uint64_t b, uint64_t c, uint64_t d);
/* quartet_lookup_table.hpp:87-111: slot (0,1,2) of the pairing ab|cd inside the tuple of {a,b,c,d} */
int qso_tuple_index(uint64_t a, uint64_t b, uint64_t c, uint64_t d);

/* TreeInformation.hpp:40-43,71-113 applied to one (gene) tree: n x n topological distances in edges
 * between taxa, QSO_MISSING where a taxon is absent from the tree.  Returns max finite distance. */
int qso_distance_matrix(int n_nodes, const int32_t* parent, const int32_t* leaf_id, int n_taxa, uint16_t* D);

/* QuartetCounterLookup.hpp:66-238 (clade enumeration over Euler-tour leaf ranges), compact table
 * (savemem = true): table[rank*3 + slot] += 1 at BOTH ends of a resolved quartet's central path,
 * i.e. 2 per tree.  table has C(n,4)*3 uint32 entries, caller zeroes it. */
int qso_count_clades_compact(int n_taxa, int n_trees, const int64_t* node_off, const int32_t* parent,
                             const int32_t* leaf_id, uint32_t* table);
/* Same enumeration into the n^4 "fast" table (QuartetCounterLookup.hpp:90) followed by the 4-cell
 * lookup (QuartetCounterLookup.hpp:283-290,313-315) for every sorted quartet: canonical counts,
 * 1 per tree.  Only for n_taxa <= 64 (n^4 uint32). table = C(n,4)*3 uint32. */
int qso_count_clades_fast(int n_taxa, int n_trees, const int64_t* node_off, const int32_t* parent,
                          const int32_t* leaf_id, uint32_t* table);
/* Four-point restatement: per tree, per present quartet, unique minimum of the three
 * pair sums of topological distances.  Canonical counts, 1 per tree. */
int qso_count_fourpoint(int n_taxa, int n_trees, const int64_t* node_off, const int32_t* parent,
                        const int32_t* leaf_id, uint32_t* table);

/* The same four-point restatement for a SAMPLE of table entries (sizes where C(n,4) entries are out of reach for a
 * CPU checker): out[i*3 + slot] = canonical count of the quartet with rank ranks[i] (quartet_lookup_table.hpp:141-212
 * inverted by qso_unrank -> q[0] < q[1] < q[2] < q[3]); distances by climbing to the LCA (TreeInformation.hpp:40-43).
 * Caller zeroes out.  Returns 0, <0 on a malformed tree or a rank outside C(n_taxa,4). */
void qso_unrank(uint64_t rank, uint64_t q[4]);
int qso_count_fourpoint_ranks(int n_taxa, int n_trees, const int64_t* node_off, const int32_t* parent, const int32_t* leaf_id,
                              int64_t n_ranks, const uint64_t* ranks, uint32_t* out);

/* QuartetScoreComputer.hpp:135-159, same operation order, libm log. */
double qso_log_score(uint64_t q1, uint64_t q2, uint64_t q3);

/* Scoring.  ref tree: parent/leaf_id as above plus parent_edge[i] = index of the edge above node i
 * (-1 for the root) and child_rank (Newick order of siblings).  table = canonical counts
 * (C(n,4)*3 uint32).  count_scale = 1 (fast table semantics) or 2 (compact table, App. B1);
 * cint_bits = 8/16/32/64: stored counts are reduced mod 2^cint_bits after scaling (App. B2) and the
 * QP-IC accumulators wrap mod 2^32 (QuartetScoreComputer.hpp:382, App. B4).
 * Outputs have edge_count entries each, +inf where untouched.  If the tree is multifurcating
 * (some node has rank > 2, genesis functions.cpp:57-69) only lqic is filled (QuartetScoreComputer.hpp:513-593)
 * and *bifurcating = 0.  Returns 0, or <0 if the tree has an inner node with fewer than 3 links
 * while passing is_bifurcating (out of contract). */
int qso_score(int n_nodes, const int32_t* parent, const int32_t* leaf_id, const int32_t* parent_edge,
              const int32_t* child_rank, int n_taxa, const uint32_t* table, int count_scale, int cint_bits,
              double* lqic, double* qpic, double* eqpic, int* bifurcating);

/* QuartetScoreComputer.hpp:623-690: raw per-quartet QIC in the reference's line order
 * (lexicographic in sorted lookup ids a<b<c<d, a outermost).  For every reference-resolved quartet
 * writes topo[i] (0: ab|cd, 2: ad|bc in sorted ids) and qic[i]; unresolved quartets get topo = -1.
 * Arrays have C(n,4) entries indexed in that lexicographic order. */
int qso_raw_qic(int n_nodes, const int32_t* parent, const int32_t* leaf_id, int n_taxa,
                const uint32_t* table, int count_scale, int cint_bits, int8_t* topo, double* qic);

#ifdef __cplusplus
}
#endif
#endif
