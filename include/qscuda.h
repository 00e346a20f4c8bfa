/* include/qscuda.h — C ABI of libqscuda.so, the B200 (sm_100a) implementation of the QuartetScores
 * hot path: per-gene-tree distance matrices -> quartet topology counting -> LQ-IC / QP-IC / EQP-IC.
 *
 * The reference (lutteropp/QuartetScores @ f57c08b) has no FFI layer; the seam this ABI replaces is
 * the C++ template class boundary between main() and QuartetScoreComputer<CINT>
 * (src/QuartetScores.cpp:114-147 constructs it and calls four methods).  Each entry point below names
 * the reference interface it stands in for (paths relative to the reference repository).
 * INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions: plain pointers and sizes, no C++/torch types; every function returns 0 (QS_OK) or a
 * negative QS_E* code and never throws; qs_last_error(ctx) gives a human-readable message for the
 * last failure on that context.  Calls on one context must be serialised by the caller.  A context is
 * bound to one CUDA device and owns all its device memory.  There is no CPU fallback: without a
 * usable CUDA device qs_create fails with QS_E_CUDA.
 *
 * Taxon ("lookup") ids: position of the taxon in the reference tree's Euler-tour leaf order = left to
 * right in its Newick text (src/QuartetCounterLookup.hpp:249-258).
 * Table layout: QuartetLookupTable (src/quartet_lookup_table.hpp:135-212): entry of the quartet with
 * sorted ids s0<s1<s2<s3 is at rank C(s3,4)+C(s2,3)+C(s1,2)+s0 and holds three CINT counters
 * [#(s0s1|s2s3), #(s0s2|s1s3), #(s0s3|s1s2)].
 */
#ifndef QSCUDA_H
#define QSCUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QS_ABI_VERSION 1

enum {
    QS_OK = 0,
    QS_E_ARG = -1,          /* bad argument / call order */
    QS_E_CUDA = -2,         /* CUDA runtime failure (message has the CUDA error string) */
    QS_E_TREE = -3,         /* malformed tree encoding */
    QS_E_REFERENCE = -4,    /* reference tree not usable (ids not in planar order, degree-2 inner node, ...) */
    QS_E_MEMORY = -5,       /* table / matrices do not fit this device ("Insufficient memory!", QuartetScoreComputer.hpp:735-737) */
    QS_E_UNSUPPORTED = -6,  /* valid input outside what this build handles (message says what) */
    QS_E_STATE = -7         /* e.g. qs_score before qs_count */
};

/* mode for qs_create */
enum {
    QS_MODE_TABLE = 0,      /* keep the (sharded) count table resident: fast + qs_get_counts / qs_raw_qic possible */
    QS_MODE_TABLE_FREE = 1, /* -s / savemem analogue: counts are scored as they are produced, no table */
    QS_MODE_AUTO = 2        /* the reference constructor's memory policy (QuartetScoreComputer.hpp:724-745) applied to HBM:
                             * qs_count keeps the shard's table resident if it fits this device beside the distance matrices
                             * and falls back to table-free otherwise, instead of failing with QS_E_MEMORY (decided when the
                             * table size or the number of trees changes, not on every call) */
};

/* device for qs_create: a host-only context.  It makes no CUDA call and supports only the host-side pieces
 * of the multi-GPU path (qs_set_reference, qs_score_finalize, qs_score_num_pairs, qs_shard_range) — e.g. a
 * rank that only combines reduced partials, or CPU tests of that logic.  Every compute entry point fails
 * on it with QS_E_STATE: there is no CPU fallback. */
#define QS_DEVICE_NONE (-1)

typedef struct qs_ctx qs_ctx;

/* Replaces: construction of QuartetScoreComputer<CINT> (src/QuartetScoreComputer.hpp:698-745) — CINT
 * width selection (src/QuartetScores.cpp:115-147) is the caller's: cint_bytes in {1,2,4,8}.
 * device: CUDA ordinal.  shard_index/shard_count: this context owns the quartets whose largest
 * taxon id s3 lies in the shard's range (balanced split of the rank space by the outer index);
 * use 0/1 for a single GPU. */
int qs_create(qs_ctx** out, int n_taxa, int cint_bytes, int mode, int device, int shard_index, int shard_count);
int qs_destroy(qs_ctx* ctx);
const char* qs_last_error(const qs_ctx* ctx);
const char* qs_strerror(int code);
int qs_abi_version(void);

/* Launch all work of this context on the given cudaStream_t (passed as void*; NULL = the context's
 * own non-blocking stream; for the legacy default stream pass cudaStreamLegacy, (void*)1).  Lets a host
 * framework time the kernels with its own events and order its collectives with them. */
int qs_set_stream(qs_ctx* ctx, void* cuda_stream);

/* Table-free contexts score while they count, so the count semantics must be known before qs_count:
 * 1 = runtime-efficient table (default), 2 = the reference's -s table (doubled, CINT-wrapped counts). */
int qs_set_count_scale(qs_ctx* ctx, int count_scale);

/* Replaces: the reference-tree half of the QuartetScoreComputer / QuartetCounterLookup constructors
 * (QuartetScoreComputer.hpp:710-718, QuartetCounterLookup.hpp:249-258, TreeInformation.hpp:95-113).
 * Nodes are numbered as genesis numbers them (root 0, parent[i] < i), parent_edge[i] is the genesis
 * edge index above node i (-1 for the root) so edge-indexed outputs line up with the Newick writer,
 * children in Newick order through first_child/next_sibling, leaf_lookup_id[i] >= 0 exactly for leaves. */
int qs_set_reference(qs_ctx* ctx, int n_nodes, const int32_t* parent, const int32_t* parent_edge,
                     const int32_t* leaf_lookup_id, const int32_t* first_child, const int32_t* next_sibling);

/* Replaces: the per-tree part of QuartetCounterLookup::countQuartets (QuartetCounterLookup.hpp:202-221):
 * the host flattens each evaluation tree while it parses (unknown taxon names are rejected there,
 * where the reference throws, :218).  Tree t owns nodes [node_offsets[t], node_offsets[t+1]); inside a
 * tree parent[i] < i (local indices), parent = -1 for the root, leaf_lookup_id >= 0 exactly for leaves,
 * every id at most once per tree.  May be called repeatedly (streaming); host buffers are copied. */
int qs_add_trees(qs_ctx* ctx, int n_trees, const int64_t* node_offsets, const int32_t* parent,
                 const int32_t* leaf_lookup_id);
int qs_clear_trees(qs_ctx* ctx);
int qs_num_trees(const qs_ctx* ctx, int64_t* n_trees);

/* Replaces: QuartetCounterLookup::countQuartets + updateQuartets* (QuartetCounterLookup.hpp:66-238) and
 * the TreeInformation distance queries (TreeInformation.hpp:40-43) applied per gene tree.
 * Builds every tree's n x n distance matrix on the device and counts, for every quartet of this
 * shard and every tree, the displayed topology.  Limit: distances are held as exact small integers in fp16, so a gene
 * tree with a leaf-to-leaf path longer than 2,048 edges is refused with QS_E_UNSUPPORTED (the reference's distances are
 * `unsigned`, TreeInformation.hpp:40-43; such a tree needs more than 2,049 taxa or long chains of unary nodes).  QS_MODE_TABLE leaves the table resident;
 * QS_MODE_TABLE_FREE also accumulates the scoring partials (qs_count then implies the scan). */
int qs_count(qs_ctx* ctx);

/* Replaces: computeQuartetScoresBifurcating / processNodePair / computeQuartetScoresMultifurcating
 * (QuartetScoreComputer.hpp:379-593) and getLQICScores/getQPICScores/getEQPICScores (:106-126).
 * Each output has edge_count = n_nodes-1 doubles indexed by genesis edge index, +inf where untouched
 * (:763,772-774).  For a multifurcating reference only lqic is computed; qpic/eqpic (may be NULL) are
 * filled with +inf.  Limit: the LQ-IC selection packs three counts into 63 bits, so n_trees * count_scale must stay below
 * 2^21 (2,097,152) — beyond that qs_score fails with QS_E_UNSUPPORTED (the reference's log_score takes size_t and has
 * no such limit; QuartetScoreComputer.hpp:135).  count_scale: 1 = the reference's runtime-efficient table semantics, 2 = its
 * memory-efficient (-s) table, whose stored counts are doubled and wrap in CINT;
 * the 32-bit wrap of the QP-IC accumulators (QuartetScoreComputer.hpp:382) is always reproduced unless
 * exact_qp != 0.  With shard_count > 1 this returns the scores of this shard's quartets only; use
 * qs_score_partials + qs_score_finalize around an all-reduce instead. */
int qs_score(qs_ctx* ctx, int count_scale, int exact_qp, double* lqic, double* qpic, double* eqpic);

/* Multi-GPU: partial results of this shard.  lqic_partial: edge_count doubles (min over this shard's
 * quartets, +inf if none) -> all-reduce MIN.  pair_sums: 3 * n_pairs uint64 (sum over this shard's
 * quartets of the three topology counts of every inner-node pair) -> all-reduce SUM.
 * qs_score_finalize turns reduced partials into the three score vectors on the host. */
int qs_score_num_pairs(const qs_ctx* ctx, int64_t* n_pairs);
/* Layout of the per-pair arrays: n_pairs = I x I for the I inner nodes of the reference tree; the pair {u,v} of inner indices
 * u < v lives at u * I + v (the other half is unused).  Inner indices follow the planar leaf order (an inner node is ranked by
 * the gap between the last leaf of its first child and the next leaf; nodes with a single child come last):
 * qs_score_inner_nodes writes the reference-tree node id of every inner index (node_of_inner may be NULL to query I only). */
int qs_score_inner_nodes(const qs_ctx* ctx, int32_t* node_of_inner, int64_t capacity, int64_t* n_inner);
int qs_score_partials(qs_ctx* ctx, int count_scale, double* lqic_partial, uint64_t* pair_sums);
int qs_score_finalize(qs_ctx* ctx, int exact_qp, const double* lqic_reduced, const uint64_t* pair_sums_reduced,
                      double* lqic, double* qpic, double* eqpic);

/* Multi-GPU without the host in the data path (one process per GPU; NCCL over NVLink).  The per-inner-node-pair
 * partials stay on the device and the CALLER's collectives run on them in place, on the context's stream:
 *   qs_score_scan(ctx, count_scale)              scan this shard's table (table-free contexts: already done by qs_count);
 *                                                asynchronous, nothing is copied to the host
 *   qs_score_device_partials(...)                device pointers, n_pairs elements each (x3 for pair_sums), all int64-compatible:
 *        pair_sums   uint64 [n_pairs][3]  topology sums (QuartetScoreComputer.hpp:429-431)            -> all-reduce SUM
 *        pair_score  int64  [n_pairs]     order-preserving image of the pair's minimal QIC (:432), max = none -> all-reduce MIN
 *        pair_best   int64  [n_pairs]     count triple of the quartet that attains it, max = none
 *   qs_score_select_winners(ctx)                 after the MIN all-reduce of pair_score: shards that do not hold the
 *                                                winning score drop their triple                          -> all-reduce MIN of pair_best
 *   qs_score_finish(ctx, exact_qp, lqic, qpic, eqpic)   per-edge minima on the device, log_score of the selected triples / sums
 *                                                on the host (libm, reference operation order); same outputs as qs_score.
 * With one shard, qs_score == qs_score_scan + qs_score_finish.  Replaces the same reference lines as qs_score. */
int qs_score_scan(qs_ctx* ctx, int count_scale);
int qs_score_device_partials(qs_ctx* ctx, void** pair_sums, void** pair_score, void** pair_best, int64_t* n_pairs);
int qs_score_select_winners(qs_ctx* ctx);
int qs_score_finish(qs_ctx* ctx, int exact_qp, double* lqic, double* qpic, double* eqpic);

/* Parity hook for QuartetCounterLookup::countQuartetOccurrences (QuartetCounterLookup.hpp:300-318):
 * canonical per-tree counts (1 per tree) of the entries [rank_begin, rank_end) in table layout, 3*cint_bytes
 * per rank, copied to the host buffer `out`.  Ranks outside this shard read as zero.  A table-free context keeps no
 * table: it counts the slabs that cover the range again (slow, but the same numbers). */
int qs_get_counts(qs_ctx* ctx, uint64_t rank_begin, uint64_t rank_end, void* out);
/* Rank range [begin,end) owned by this shard. */
int qs_shard_range(const qs_ctx* ctx, uint64_t* rank_begin, uint64_t* rank_end);
/* The sharding rule itself (pure function, no context): shard g of G owns the quartets whose largest id s3
 * lies in [s3_begin, s3_end), i.e. the rank range [C(s3_begin,4), C(s3_end,4)); the boundaries equalise the counting
 * kernel's cost (work items including their padding), not the number of quartets.  Output pointers may be NULL. */
int qs_shard_bounds(int n_taxa, int shard_index, int shard_count, int* s3_begin, int* s3_end, uint64_t* rank_begin, uint64_t* rank_end);

/* The default ranges assume that every gene tree needs all three topology compares (class B).  Gene trees that contain all
 * taxa and are fully resolved (class A) skip one role, which moves the cost balance between low and high s3; once the trees
 * are added, qs_rebalance_shards classifies them (distance kernels only) and re-cuts the ranges for the observed class mix.
 * Every shard must call it with the same trees (they then agree without talking to each other); *changed = 1 if this
 * shard's range moved — counts and scores of earlier calls are invalid then, call qs_count again.  Optional: results do not
 * depend on it, only the balance.  qs_table_resident: 1 if the last qs_count left this shard's table in device memory
 * (QS_MODE_TABLE, or QS_MODE_AUTO when it fits), 0 if it ran table-free. */
int qs_rebalance_shards(qs_ctx* ctx, int* changed);
int qs_table_resident(const qs_ctx* ctx, int* resident);

/* Diagnostics of the counting kernel's host-built task table for the quartets with s3 in [s3_begin, s3_end)
 * (pure host function, no context, no GPU).  stats[17] = {X tasks, Y tasks, XO items, XD items, Y + Z items,
 * XO item slots, XD item slots, Y + Z item slots, staged rows summed over tasks, max rows of a task,
 * self-check violations (must be 0: tasks tile each enumeration, every item's rows are staged, the items of role X keep
 * every quartet of the range exactly once, and so do those of roles Y and Z together), quartets,
 * XR items, XR item slots, compares issued by role X per tree (2 x quartets of them are useful), Z items (role-Y work
 * on the d whose block of 8 the range cuts), compares issued by roles Y and Z per tree (1 x quartets of them are useful)}. */
int qs_plan_stats(int n_taxa, int s3_begin, int s3_end, int64_t* stats);

/* Parity hook for the distance kernel: tree t's matrix as n x n uint16 (0xFFFF = taxon missing). */
int qs_get_distances(qs_ctx* ctx, int64_t tree, uint16_t* out);

/* Replaces: printRawQICScores (QuartetScoreComputer.hpp:623-690), which the reference runs with either table type
 * (src/QuartetScores.cpp:120-122).  Writes "(a,b|c,d): qic" lines in the reference's order and formatting; taxon_names[id] is
 * the label of lookup id `id`.  The file is ordered by (a,b,c,d), the table by rank, so the whole table (C(n,4) x 3 x
 * cint_bytes) is gathered in host memory first; a table-free context counts its slabs again for that.
 * qs_write_raw_qic: one context that owns the whole rank space.  qs_write_raw_qic_shards: the contexts of all shards (same
 * process, shard 0 .. G-1 in order); the reference tree and the error message live in ctxs[0]. */
int qs_write_raw_qic(qs_ctx* ctx, int count_scale, const char* const* taxon_names, const char* path);
int qs_write_raw_qic_shards(qs_ctx* const* ctxs, int n_ctxs, int count_scale, const char* const* taxon_names, const char* path);

/* ---- host ingest -------------------------------------------------------------------------
 * Replaces: the two serial genesis parses of the evaluation-tree file (src/QuartetScores.cpp:23-32
 * countEvalTrees, src/QuartetCounterLookup.hpp:202-221 NewickInputIterator loop + the per-tree Euler leaf list,
 * :207-219) by ONE multi-threaded pass from Newick text straight to the qs_add_trees encoding.  taxon_names[id]
 * is the label of lookup id `id` (n_taxa entries, the reference tree's leaves left to right).  An evaluation-tree
 * taxon that is not among them fails with QS_E_TREE and a message naming it (the reference throws
 * std::out_of_range there, QuartetCounterLookup.hpp:218).  n_threads <= 0: all host cores.  The result does not
 * depend on the thread count.  No GPU and no context needed for qs_newick_flatten. */
typedef struct qs_flat_trees qs_flat_trees;
int qs_newick_flatten(const char* text, size_t text_len, int n_taxa, const char* const* taxon_names, int n_threads,
                      qs_flat_trees** out, char* errbuf, size_t errbuf_len);
int qs_flat_trees_view(const qs_flat_trees* f, int64_t* n_trees, int64_t* n_nodes, const int64_t** node_offsets,
                       const int32_t** parent, const int32_t** leaf_lookup_id);
void qs_flat_trees_free(qs_flat_trees* f);
/* flatten + qs_add_trees in one call, from memory or from a file (read whole; one tree or many, ';'-terminated) */
int qs_add_newick(qs_ctx* ctx, const char* text, size_t text_len, const char* const* taxon_names, int n_threads, int64_t* n_trees_added);
int qs_add_newick_file(qs_ctx* ctx, const char* path, const char* const* taxon_names, int n_threads, int64_t* n_trees_added);

/* ---- table persistence -------------------------------------------------------------------
 * The reference keeps its table only in RAM (the STXXL external-memory idea is commented out,
 * src/quartet_lookup_table.hpp:3,218-222).  qs_save_table writes this shard's counted table (QS_MODE_TABLE, after
 * qs_count) to `path`: a 64-byte header {magic "QSTBL001", n_taxa, cint_bytes, n_trees, s3_begin, s3_end,
 * rank_begin, rank_end} followed by the entries in QuartetLookupTable layout.  qs_load_table restores it into a
 * context created with the same n_taxa / cint_bytes / shard, which can then score against any reference tree
 * on the same taxon ids without recounting. */
int qs_save_table(qs_ctx* ctx, const char* path);
int qs_load_table(qs_ctx* ctx, const char* path);

/* Device timing of the last qs_count / scoring pass (CUDA events on the context's stream), and the
 * number of kernels this library launched since the context was created. */
int qs_last_timing(const qs_ctx* ctx, double* dist_ms, double* count_ms, double* score_ms);
int qs_launch_count(const qs_ctx* ctx, int64_t* n_launches);

/* Tree classes found by the last qs_count: class A = gene trees that contain all n taxa and have no node of
 * degree > 3, so they resolve every quartet and need two compares per quartet instead of three
 * (kernels/count_rows.cuh).  Reported so that bench.py can state the algorithmic work of a run. */
int qs_tree_classes(const qs_ctx* ctx, int64_t* n_class_a, int64_t* n_class_b);

/* Live micro-benchmark of the issue rate the counting kernel is bound by: packed fp16x2 compare
 * (HSET2 -> integer mask, two per three-input IADD3 accumulate) lane-operations per second on this
 * device, counting the HSET2 only, and the plain int32 LOP3/IADD3 rate, both measured at the clocks
 * the device sustains right now.  Used as the roofline denominator in bench.py. */
int qs_measure_alu_peak(qs_ctx* ctx, double* hset2_laneops_per_s, double* int32_laneops_per_s);

#ifdef __cplusplus
}
#endif
#endif /* QSCUDA_H */
