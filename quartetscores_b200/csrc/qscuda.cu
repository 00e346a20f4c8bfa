// qscuda.cu — host side of libqscuda.so: context, C ABI (include/qscuda.h), kernel orchestration.
//
// B200-native (sm_100a) re-design of the QuartetScores hot path (reference: lutteropp/QuartetScores,
// src/QuartetCounterLookup.hpp, src/quartet_lookup_table.hpp, src/TreeInformation.hpp,
// src/QuartetScoreComputer.hpp).  There is no CPU fallback in this file: every count and every
// per-quartet QIC selection happens in the CUDA kernels under kernels/.
#include "../../include/qscuda.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <limits>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "kernels/common.cuh"
#include "kernels/count_rows.cuh"
#include "kernels/dist.cuh"
#include "kernels/score.cuh"
#include "kernels/ubench.cuh"

using namespace qs;

namespace {


constexpr int kMaxHalfExact = 2048;     // integers up to 2048 are exact in fp16 (distances; the counters are integer)
constexpr size_t kTableSlack = 64;      // the scan reads the table in 48-byte groups: the last one may reach past the last entry

struct HostRef {
    int n_nodes = 0, n_inner = 0;
    bool bifurcating = false;
    std::vector<int32_t> parent, parent_edge, leaf_id, first_child, next_sib, depth, inner_index, inner_node, leaf_node;
    std::vector<uint16_t> lca;          // [n][n] inner index
    std::vector<uint16_t> idepth;       // [I]
    // scan kernel (kernels/score.cuh): rows of lca[][] run-length encoded, parents in inner-index space
    std::vector<int32_t> inner_parent, leaf_parent, inner_gap;
};

}  // namespace

struct qs_ctx {
    int device = 0, n = 0, n_pad = 0, cint_bytes = 2, mode = 0, shard_index = 0, shard_count = 1;
    bool host_only = false;      // QS_DEVICE_NONE: reference bookkeeping + qs_score_finalize only
    bool auto_mode = false;      // QS_MODE_AUTO: `mode` is decided by qs_count whenever the table size or the tree count changed (table if it fits this device, else table-free)
    int* h_flags = nullptr;      // pinned: {max distance, tree error, |A|} read back asynchronously by qs_count
    int d_begin = 0, d_end = 0, num_sms = 148, smem_optin = 0;
    uint64_t rank_begin = 0, rank_end = 0;
    std::string err;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int64_t launches = 0;
    double dist_ms = 0, count_ms = 0, score_ms = 0;
    bool score_timed = false;            // ev[4]/ev[5] bracket a scan whose time has not been read yet

    bool has_ref = false;
    HostRef ref;
    uint32_t* d_lcapd = nullptr;          // [n][n] lca inner index | depth << 16
    uint16_t* d_idepth = nullptr;
    int32_t *d_inner_parent = nullptr, *d_leaf_parent = nullptr, *d_inner_gap = nullptr;
    int32_t *d_inner_node = nullptr, *d_node_parent = nullptr, *d_node_depth = nullptr, *d_node_edge = nullptr, *d_node_inner = nullptr;
    long long* d_edge = nullptr;          // [4][E] edge minima / arg-minima of the per-edge reduction
    unsigned long long* d_edge_out = nullptr;   // [E][7]
    unsigned long long* d_scan_scratch = nullptr;
    size_t scan_scratch_bytes = 0;
    int* d_scan_counter = nullptr;
    int4* d_scan_items = nullptr;         // work items of the scan for the d-range [scan_dB, scan_dE)
    size_t scan_items_cap = 0;
    int64_t scan_n_items = 0;
    int scan_dB = -1, scan_dE = -1;

    // trees
    int64_t m = 0, total_nodes = 0, cap_nodes = 0, cap_trees = 0;
    int max_nodes = 0;
    std::vector<int64_t> h_off{0};
    int64_t* d_off = nullptr;
    int32_t* d_parent = nullptr;
    int32_t* d_leaf = nullptr;

    // distances
    __half* d_D = nullptr;
    size_t D_cap = 0;
    int* d_flags = nullptr;      // [0] max distance, [1] tree error
    bool dist_valid = false;

    // tree classes (kernels/dist.cuh): class A = complete and fully resolved
    int32_t* d_class = nullptr;
    int32_t* d_order = nullptr;
    int32_t* d_nA = nullptr;
    int64_t class_cap = 0, n_class_a = 0;
    int* d_counter = nullptr;    // dynamic task / tile scheduling

    // counting (kernels/count_rows.cuh)
    void* d_table = nullptr;     // QS_MODE_TABLE: this shard's table; QS_MODE_TABLE_FREE: the current slab
    size_t table_bytes = 0;
    RowTask* d_tasks = nullptr;  // task table of the d-range [plan_dB, plan_dE)
    size_t tasks_cap = 0;
    int plan_dB = -1, plan_dE = -1, plan_with_y = 0, plan_threads = 0, plan_nx = 0, plan_ny = 0, plan_max_rows = 1;
    std::vector<float> plan_xcost, plan_ycost;
    int64_t chunk_key[4] = {-1, -1, -1, -1}, chunk_choice = 1;      // cached chunk count for (plan generation, m, |A| hint, CTA slots)
    int64_t plan_generation = 0;
    int64_t* d_enum = nullptr;   // PXO | PXD | PY | CD | PXR | PZ prefix tables of the plan
    size_t enum_cap = 0;
    int plan_zF = 0, plan_zL = 0;
    bool counted = false;
    bool counted_once = false;   // n_class_a holds the class split of an earlier qs_count on this context

    // scoring
    unsigned long long* d_pair_sums = nullptr;      // [I*I][3]
    long long* d_pair_best = nullptr;               // [I*I] packed triple
    long long* d_pair_score = nullptr;              // [I*I] ordered image of the exact score of pair_best (all-reduced MIN across shards)
    long long* d_pair_score_local = nullptr;        // [I*I] this shard's own copy (qs_score_select_winners)
    std::vector<unsigned long long> h_pair_sums, h_edge_out;
    bool fused_partials_valid = false;   // table-free mode: partials accumulated by qs_count
    int fused_scale = 1;
    bool partials_ready = false;         // the device partials hold a finished scan (qs_score_scan / table-free qs_count)
    size_t auto_need = ~(size_t)0;       // QS_MODE_AUTO: the table size and tree count the current mode was chosen for
    int64_t auto_m = -1;

};

namespace {

#define QS_FAIL(ctx, code, ...)                       \
    do {                                              \
        char _b[512];                                 \
        snprintf(_b, sizeof(_b), __VA_ARGS__);        \
        (ctx)->err = _b;                              \
        return (code);                                \
    } while (0)

#define QS_CUDA(ctx, call)                                                                                   \
    do {                                                                                                     \
        cudaError_t _e = (call);                                                                             \
        if (_e != cudaSuccess) QS_FAIL(ctx, QS_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

template <typename T>
int dev_alloc(qs_ctx* c, T** p, size_t count) {
    if (*p) { cudaFree(*p); *p = nullptr; }
    if (count == 0) return QS_OK;
    cudaError_t e = cudaMalloc((void**)p, count * sizeof(T));
    if (e != cudaSuccess) {
        cudaGetLastError();
        QS_FAIL(c, QS_E_MEMORY, "Insufficient memory! cudaMalloc of %zu bytes failed: %s", count * sizeof(T), cudaGetErrorString(e));
    }
    return QS_OK;
}

// Split of the rank space by the outer index d into G contiguous ranges of (nearly) equal COUNTING COST — not equal
// quartet counts: what a shard pays is the work items of the counting kernel (kernels/count_rows.cuh), i.e. 8 x 8 blocks
// including their padding.  Per d: the role-X items of all c < d (full blocks at 1, ragged-block items by the share of the
// block they run over, diagonal blocks at 1 or 1/2; a (c,d) never costs less than the share of a task it blocks when the
// task's row budget, not its item count, ends the task — small c); per whole block of 8 consecutive d inside the shard: the
// role-Y items of all (b,c) below it; per d of a block the shard's range cuts: its role-Z items (z_split, declared below); both
// only for the class-B trees: weight 0.76 x class_b_fraction (measured item costs, round 1, class-B trees,
// profiles/r01_zj_*: 2.78e-5 ms per X item, 2.12e-5 ms per Y item).  The boundaries minimise the largest shard cost
// (greedy fill under a bisected threshold, exact for contiguous partitions of a monotone cost); a pure function of its
// arguments: every rank computes the same ranges.  (use_xo_diag / cr_threads_for are declared further down; their rules
// are repeated here.)
void z_split(int dB, int dE, int* z_first, int* z_last);

void shard_bounds(int n, int g, int G, double class_b_fraction, int* d_begin, int* d_end) {
    if (G <= 1) { *d_begin = 3; *d_end = n; return; }
    const int xo_diag = n > 160 ? 1 : 0;
    const double T = n <= 112 ? 512 : 256;                                              // thread-items per task
    const double R = (double)std::max<size_t>(4, std::min<size_t>((size_t)n, (24 * 1024) / ((size_t)((n + 7) / 8 * 8) * 2))) - 1;   // d rows a task can stage
    const double wy = 0.76 * std::min(1.0, std::max(0.0, class_b_fraction));
    std::vector<double> PX(n + 1, 0.0), PYb((n >> 3) + 2, 0.0), PZc(n + 2, 0.0);
    {
        std::vector<double> A(n + 1, 0.0), B(n + 1, 0.0);       // prefix over c of the X cost of one (c,d) / of the Y a-blocks of one (b,c)
        for (int c = 2; c < n; ++c) {
            const int nf = c >> 3, rg = c & 7, nd = ((c - 2) >> 3) + 1;
            double xo = nf * (nf - 1) / 2 + (xo_diag ? nf : 0), xr = 0, xd = xo_diag ? 0 : 0.5 * nd;
            if (rg) xr = (nf + ((xo_diag && rg >= 2) ? 1 : 0)) * (0.14 + 0.86 * rg / 8.0);
            if (xo > 0) xo = std::max(xo, T / R);
            if (xr > 0) xr = std::max(xr, (0.14 + 0.86 * rg / 8.0) * T / R);
            if (xd > 0) xd = std::max(xd, T / R);
            // every (c,d) also brings a matrix row of its own into its tasks' staging: an empirical 31e-6 n^2 items' worth per pair
            // (the first shard, whose small d hold many pairs with few blocks each, ran 3 % (n = 500) and 5-7 % (n = 1000) over the
            // others without it: profiles/r02_y_shards.txt, r02_z_bench_cfg4_n8.json)
            A[c + 1] = A[c] + xo + xr + xd + 31e-6 * n * n;
            double bc = 0;
            for (int bb = 1; bb < c; ++bb) bc += (bb + 7) >> 3;
            B[c + 1] = B[c] + bc;
        }
        for (int d = 3; d < n; ++d) PX[d + 1] = PX[d] + A[d];                          // X cost with this d: all c < d
        for (int k = 0; k <= (n - 1) >> 3; ++k) {                                       // Y items of d-block k: all (b,c) with c + 1 <= 8k + 7
            const int cmax = std::min(8 * k + 6, n - 2);
            PYb[k + 1] = PYb[k] + (cmax >= 2 ? wy * B[cmax + 1] : 0.0);
        }
    }
    for (int d = 0; d <= n; ++d) {                                                      // Z items of one d: all a <= d - 3
        double z = 0;
        if (d < n) for (int a = 0; a + 3 <= d; ++a) z += cr_nz(a, d);
        PZc[d + 1] = PZc[d] + 0.82 * wy * z;          // a Z item against a Y item: fewer rows staged per task (8 shards of n = 500, profiles/r02_y_shards.txt)
    }
    auto cost = [&](int b, int e) -> double {                                           // shard [b, e), b < e
        int zf, zl;
        z_split(b, e, &zf, &zl);
        return PX[e] - PX[b] + (PYb[zl >> 3] - PYb[zf >> 3]) + (PZc[zf] - PZc[b]) + (PZc[e] - PZc[zl]);
    };
    // smallest threshold for which greedy filling needs <= G shards
    auto fill = [&](double thr, std::vector<int>* out) -> int {
        int cnt = 0, b = 3;
        while (b < n) {
            int e = b + 1;
            while (e < n && cost(b, e + 1) <= thr) ++e;
            ++cnt; b = e;
            if (out) out->push_back(e);
            if (cnt > G) return cnt;
        }
        return cnt;
    };
    double lo = 0, hi = cost(3, n);
    for (int it = 0; it < 60; ++it) {
        const double mid = 0.5 * (lo + hi);
        if (fill(mid, nullptr) <= G) hi = mid; else lo = mid;
    }
    std::vector<int> ends;
    fill(hi, &ends);
    while ((int)ends.size() < G) ends.push_back(n);                                     // fewer d values than shards: the last shards are empty
    ends[G - 1] = n;
    *d_begin = g == 0 ? 3 : ends[g - 1];
    *d_end = ends[g];
    if (*d_end < *d_begin) *d_end = *d_begin;
}

double host_log_score(uint64_t q1, uint64_t q2, uint64_t q3) {
    // src/QuartetScoreComputer.hpp:135-159 — evaluated on the HOST with the same libm and operation
    // order as the reference; only applied to integer results produced on the GPU.
    if (q1 == 0 && q2 == 0 && q3 == 0) return 0;
    uint64_t sum = q1 + q2 + q3;
    double p1 = (double)q1 / sum, p2 = (double)q2 / sum, p3 = (double)q3 / sum;
    double qic = 1;
    if (p1 != 0) qic += p1 * log(p1) / log(3);
    if (p2 != 0) qic += p2 * log(p2) / log(3);
    if (p3 != 0) qic += p3 * log(p3) / log(3);
    if (q1 < q2 || q1 < q3) return qic * -1;
    return qic;
}

uint64_t cint_mask(int bytes) { return bytes >= 8 ? ~0ull : ((1ull << (8 * bytes)) - 1); }

void free_all(qs_ctx* c) {
    cudaFree(c->d_lcapd); cudaFree(c->d_idepth);
    cudaFree(c->d_inner_parent); cudaFree(c->d_leaf_parent); cudaFree(c->d_inner_gap);
    cudaFree(c->d_inner_node); cudaFree(c->d_node_parent); cudaFree(c->d_node_depth); cudaFree(c->d_node_edge); cudaFree(c->d_node_inner);
    cudaFree(c->d_edge); cudaFree(c->d_edge_out); cudaFree(c->d_scan_scratch); cudaFree(c->d_scan_counter); cudaFree(c->d_scan_items);
    cudaFree(c->d_off); cudaFree(c->d_parent); cudaFree(c->d_leaf);
    cudaFree(c->d_D); cudaFree(c->d_flags); cudaFree(c->d_table);
    cudaFree(c->d_class); cudaFree(c->d_order); cudaFree(c->d_nA); cudaFree(c->d_counter);
    cudaFree(c->d_tasks); cudaFree(c->d_enum);
    cudaFree(c->d_pair_sums); cudaFree(c->d_pair_best); cudaFree(c->d_pair_score); cudaFree(c->d_pair_score_local);
    if (c->h_flags) cudaFreeHost(c->h_flags);
    for (auto& e : c->ev) if (e) cudaEventDestroy(e);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
}

int run_distances(qs_ctx* c) {
    const size_t need = (size_t)c->m * c->n * c->n_pad;
    if (need > c->D_cap) {
        int r = dev_alloc(c, &c->d_D, need);
        if (r) { c->D_cap = 0; return r; }
        c->D_cap = need;
    }
    if (c->m > c->class_cap) {
        int r;
        if ((r = dev_alloc(c, &c->d_class, (size_t)c->m))) { c->class_cap = 0; return r; }
        if ((r = dev_alloc(c, &c->d_order, (size_t)c->m))) { c->class_cap = 0; return r; }
        c->class_cap = c->m;
    }
    // (no memset of D: the distance kernels write every entry of every matrix, NaN = 0xFFFF = "taxon missing" included)
    QS_CUDA(c, cudaMemsetAsync(c->d_flags, 0, 2 * sizeof(int), c->stream));
    DistArgs da;
    da.node_off = c->d_off; da.parent = c->d_parent; da.leaf_id = c->d_leaf;
    da.m = (int)c->m; da.n = c->n; da.n_pad = c->n_pad; da.max_nodes = c->max_nodes; da.D = c->d_D; da.max_dist = c->d_flags;
    da.tree_class = c->d_class;
    // QS_DIST_SMEM_MATRIX=1: the warp kernel keeps the matrix of a small tree in shared memory (kernels/dist.cuh).  Measured slower
    // (0.47 against 0.27 ms at cfg2: 8 warps per SM instead of 48), so it stays an opt-in
    bool smem_matrix = (size_t)c->n * c->n_pad * 2 <= 24 * 1024 && dist_warp_smem_per_warp(c->max_nodes, c->n, true) * DW_WARPS + 1024 <= (size_t)c->smem_optin;
    if (const char* env = getenv("QS_DIST_SMEM_MATRIX")) smem_matrix = smem_matrix && atoi(env) != 0;
    else smem_matrix = false;
    const size_t warp_smem = dist_warp_smem_per_warp(c->max_nodes, c->n, smem_matrix) * DW_WARPS;
    if (c->max_nodes <= 2048 && c->n <= 32767 && warp_smem <= (size_t)c->smem_optin) {
        // small trees: one warp per tree, as many trees in flight as shared memory allows
        int per_sm = std::max(1, std::min(8, (int)((size_t)c->smem_optin / (warp_smem + 1024))));
        int grid = (int)std::min<int64_t>((c->m + DW_WARPS - 1) / DW_WARPS, (int64_t)c->num_sms * per_sm);
        if (smem_matrix) {
            QS_CUDA(c, cudaFuncSetAttribute(qs_dist_warp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)warp_smem));
            qs_dist_warp_kernel<true><<<grid, 32 * DW_WARPS, warp_smem, c->stream>>>(da);
        } else {
            QS_CUDA(c, cudaFuncSetAttribute(qs_dist_warp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)warp_smem));
            qs_dist_warp_kernel<false><<<grid, 32 * DW_WARPS, warp_smem, c->stream>>>(da);
        }
    } else {
        size_t smem = (size_t)8 * 4 * c->max_nodes + (size_t)((c->n + 31) / 32) * 4;
        if (smem > (size_t)c->smem_optin) QS_FAIL(c, QS_E_UNSUPPORTED, "gene tree with %d nodes exceeds the distance kernel's shared-memory budget", c->max_nodes);
        QS_CUDA(c, cudaFuncSetAttribute(qs_dist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = std::max(1, std::min(8, (int)((size_t)c->smem_optin / std::max<size_t>(smem, 1024))));
        int grid = (int)std::min<int64_t>(c->m, (int64_t)c->num_sms * per_sm);
        qs_dist_kernel<<<grid, 256, smem, c->stream>>>(da);
    }
    c->launches++;
    QS_CUDA(c, cudaGetLastError());
    qs_order_kernel<<<1, 1024, 0, c->stream>>>(c->d_class, (int)c->m, c->d_order, c->d_nA);
    c->launches++;
    QS_CUDA(c, cudaGetLastError());
    c->dist_valid = true;
    return QS_OK;
}

// ---- task table of the counting kernel for the quartets with d in [dB, dE) (see kernels/count_rows.cuh) ----
struct HostEnum {
    std::vector<int64_t> PXO, PXD, PY, CD, PXR, PZ;
    int z_first = 0, z_last = 0;
    int xo_diag = 0;
    EnumTables view() const { return EnumTables{PXO.data(), PXD.data(), PY.data(), CD.data(), PXR.data(), PZ.data(), z_first, z_last, xo_diag}; }
    const int64_t* prefix(int kind) const { return kind == ITEM_XO ? PXO.data() : kind == ITEM_XD ? PXD.data() : kind == ITEM_XR ? PXR.data() : kind == ITEM_Y ? PY.data() : PZ.data(); }
    int64_t total(int kind) const { return kind == ITEM_XO ? PXO.back() : kind == ITEM_XD ? PXD.back() : kind == ITEM_XR ? PXR.back() : kind == ITEM_Y ? PY.back() : PZ.back(); }
};

// separate half-cost tasks for the diagonal blocks pay off while a task's rows can cover most of the matrix
int use_xo_diag(int n) { return n > 160 ? 1 : 0; }

// which d of [dB, dE) role Y takes (whole d-blocks) and which role Z (kernels/count_rows.cuh): Z gets the d of the blocks the range
// cuts; a range with fewer than 8 whole blocks goes to Z altogether — its role-Y tasks would hold a handful of items per matrix
// row and be bound by staging the rows, not by comparing (the narrow shards of many GPUs: 8 shards of n = 500 count in 47-50 ms each
// with Z against 52-56 ms with Y, profiles/r02_w2_shards_n500B.txt); over a wide range Y is the faster one (n = 500, one shard: 389 ms
// against 396 ms with Z for every d, profiles/r02_x_z_everywhere.txt) — its 8 consecutive a are neighbours in the table, Z's flush scatters
void z_split(int dB, int dE, int* z_first, int* z_last) {
    int zf = std::min(dE, (dB + 7) & ~7), zl = std::max(zf, dE & ~7);
    // (tuning hooks, read once: shard_bounds calls this ~60 n times)
    static const int min_blocks = [] { const char* env = getenv("QS_Z_MIN_BLOCKS"); return env ? atoi(env) : 8; }();
    static const int max_span = [] { const char* env = getenv("QS_Z_MAX_SPAN"); return env ? atoi(env) : 80; }();
    if (((zl - zf) >> 3) < min_blocks && dE - dB <= max_span) zf = zl = dE;
    *z_first = zf; *z_last = zl;
}

void build_enum_tables(int n, int dB, int dE, HostEnum& H) {
    H.xo_diag = use_xo_diag(n);
    z_split(dB, dE, &H.z_first, &H.z_last);
    H.PXO.assign(n + 1, 0); H.PXD.assign(n + 1, 0); H.PY.assign(n + 1, 0); H.CD.assign(n + 1, 0); H.PXR.assign(n + 1, 0);
    for (int c = 0; c < n; ++c) {
        const int64_t nd = (c >= 2) ? std::max(0, dE - cr_dlo(c, dB)) : 0;
        H.PXO[c + 1] = H.PXO[c] + nd * cr_nxo(c, H.xo_diag);
        H.PXD[c + 1] = H.PXD[c] + nd * cr_nxd(c, H.xo_diag);
        H.PXR[c + 1] = H.PXR[c] + nd * cr_nxr(c, H.xo_diag);
        H.CD[c + 1] = H.CD[c] + ((c >= 2) ? cr_ndb(c, H.z_first, H.z_last) : 0);
    }
    for (int b = 0; b < n; ++b) {
        int64_t items = 0;
        if (b >= 1 && b + 1 < n) items = (int64_t)((b + 7) / 8) * (H.CD[n] - H.CD[b + 1]);
        H.PY[b + 1] = H.PY[b] + items;
    }
    const int nz = dB < dE ? cr_nzd(dB, dE, H.z_first, H.z_last) : 0;
    H.PZ.assign((size_t)nz * n + 1, 0);
    for (int i = 0; i < nz; ++i) {
        const int d = cr_zd(i, dB, H.z_first, H.z_last);
        for (int a = 0; a < n; ++a) H.PZ[(size_t)i * n + a + 1] = H.PZ[(size_t)i * n + a] + cr_nz(a, d);
    }
}

// rows touched by the items [e0, e0+ne) of one kind, as sorted disjoint intervals [lo, hi]
void task_row_intervals(const HostEnum& H, int kind, int64_t e0, int ne, int n, int dB, int dE, std::vector<std::pair<int, int>>& iv) {
    iv.clear();
    int p0, q0, p1, q1, t0, t1;
    if (kind == ITEM_Z) {
        cr_decode_z(H.view(), e0, n, dB, dE, p0, q0, t0, t1);
        cr_decode_z(H.view(), e0 + ne - 1, n, dB, dE, p1, q1, t0, t1);
    } else if (kind == ITEM_Y) {
        cr_decode_y(H.view(), e0, n, p0, q0, t0, t1);
        cr_decode_y(H.view(), e0 + ne - 1, n, p1, q1, t0, t1);
    } else {
        cr_decode_x(H.prefix(kind), kind, H.xo_diag, e0, n, dB, p0, q0, t0);
        cr_decode_x(H.prefix(kind), kind, H.xo_diag, e0 + ne - 1, n, dB, p1, q1, t0);
    }
    // outer (fixed) rows p0..p1; inner (variable) rows: tail of p0, everything of the rows in between, head of p1
    //   X kinds: outer c, inner d in [dlo(c), dE);   Y: outer b, inner c in (b, z_last - 2];   Z: outer d (one of Z's two runs of d), inner a in [0, d - 3]
    auto qlo = [&](int p) { return kind == ITEM_Z ? 0 : kind == ITEM_Y ? p + 1 : cr_dlo(p, dB); };
    auto qhi = [&](int p) { return kind == ITEM_Z ? p - 3 : kind == ITEM_Y ? H.z_last - 2 : dE - 1; };
    iv.emplace_back(p0, p1);
    if (p0 == p1) iv.emplace_back(q0, q1);
    else {
        iv.emplace_back(q0, qhi(p0));
        if (p1 - p0 >= 2) iv.emplace_back(qlo(p0 + 1), std::max(qhi(p0 + 1), qhi(p1 - 1)));
        iv.emplace_back(qlo(p1), q1);
    }
    std::sort(iv.begin(), iv.end());
    std::vector<std::pair<int, int>> m;
    for (auto& x : iv) {
        if (x.second < x.first) continue;
        if (!m.empty() && x.first <= m.back().second + 1) m.back().second = std::max(m.back().second, x.second);
        else m.push_back(x);
    }
    while (m.size() > 3) {                                               // close the smallest gap
        size_t best = 0; int gap = 1 << 30;
        for (size_t i = 0; i + 1 < m.size(); ++i) if (m[i + 1].first - m[i].second < gap) { gap = m[i + 1].first - m[i].second; best = i; }
        m[best].second = m[best + 1].second;
        m.erase(m.begin() + best + 1);
    }
    iv.swap(m);
}

void build_row_tasks(const HostEnum& H, int n, int dB, int dE, int max_rows, int threads, std::vector<RowTask>& xt, std::vector<RowTask>& yt, bool with_y = true) {
    xt.clear(); yt.clear();
    std::vector<std::pair<int, int>> iv;
    auto rows_of = [&](const std::vector<std::pair<int, int>>& v) { int r = 0; for (auto& x : v) r += x.second - x.first + 1; return r; };
    auto emit = [&](std::vector<RowTask>& out, int kind, int64_t total, int cap, int64_t first = 0) {
        int64_t e = first;
        while (e < total) {
            int ne = (int)std::min<int64_t>(cap, total - e);
            task_row_intervals(H, kind, e, ne, n, dB, dE, iv);
            while (rows_of(iv) > max_rows && ne > 1) {                   // shrink until the rows fit the shared-memory slot
                ne = std::max(1, ne / 2);
                task_row_intervals(H, kind, e, ne, n, dB, dE, iv);
            }
            RowTask t{};
            t.kind = kind; t.ne = ne; t.e0 = e;
            for (size_t k = 0; k < 3; ++k) { t.rstart[k] = k < iv.size() ? iv[k].first : 0; t.rcount[k] = k < iv.size() ? iv[k].second - iv[k].first + 1 : 0; }
            out.push_back(t);
            e += ne;
        }
    };
    // longest tasks first (the kernel hands tasks out in this order): full role-X blocks, diagonal blocks (two half items per
    // thread), then the ragged b-blocks, whose tree loop is shorter
    emit(xt, ITEM_XO, H.PXO[n], threads);
#ifdef CR_XR_BEFORE_XD
    emit(xt, ITEM_XR, H.PXR[n], threads);
    emit(xt, ITEM_XD, H.PXD[n], 2 * threads);
#else
    emit(xt, ITEM_XD, H.PXD[n], 2 * threads);
    emit(xt, ITEM_XR, H.PXR[n], threads);
#endif
    if (with_y) {
        emit(yt, ITEM_Y, H.PY[n], 2 * threads);
        const int nz = dB < dE ? cr_nzd(dB, dE, H.z_first, H.z_last) : 0, nz1 = dB < dE ? H.z_first - dB : 0;
        emit(yt, ITEM_Z, H.PZ[(size_t)nz1 * n], 2 * threads);                                   // the d below Y's blocks and the d above them: a task's
        emit(yt, ITEM_Z, H.PZ[(size_t)nz * n], 2 * threads, H.PZ[(size_t)nz1 * n]);              //   d are consecutive matrix rows
    }
}

// host half of a plan: everything that needs no CUDA call, so that a helper thread can prepare the next slab's plan
// while the GPU counts the current one (run_table_free)
// CTA shape of the counting kernel for n taxa (kernels/count_rows.cuh); QS_CR_THREADS = 512 | 256 overrides (tuning)
int cr_threads_for(int n) {
    if (const char* env = getenv("QS_CR_THREADS")) { const int t = atoi(env); if (t == CR_THREADS_BIG || t == CR_THREADS_SMALL) return t; }
    return n <= 112 ? CR_THREADS_BIG : CR_THREADS_SMALL;        // measured crossover between n = 100 and n = 128 (profiles/r01_v_shape_threshold.txt)
}

struct HostPlan {
    int dB = -1, dE = -1, with_y = 0, max_rows = 1, threads = CR_THREADS_BIG;
    HostEnum H;
    std::vector<RowTask> xt, yt;
    std::vector<float> xcost, ycost;     // per-thread work of a task per tree, in units of a full role-X item (chunk-count choice)
};

void build_host_plan(int n, int n_pad, int dB, int dE, bool with_y, int threads, HostPlan& P) {
    const size_t row_bytes = (size_t)n_pad * 2;
    // shared-memory budget per staged tree: ~24 KB (the whole matrix when n <= ~110), at least 4 rows
    const int max_rows = (int)std::max<size_t>(4, std::min<size_t>((size_t)n, (24 * 1024) / row_bytes));
    P.dB = dB; P.dE = dE; P.with_y = with_y ? 1 : 0; P.threads = threads;
    build_enum_tables(n, dB, dE, P.H);
    build_row_tasks(P.H, n, dB, dE, max_rows, threads, P.xt, P.yt, with_y);
    // per-thread cost of a task per tree: a thread runs one XO item, two half-cost XD / Y items, or one XR item over the c & 7 valid
    // taxa of its ragged block (14 % of an item is operand fetch and loop overhead whatever the block holds)
    P.xcost.clear(); P.ycost.assign(P.yt.size(), 1.f);
    for (auto& t : P.xt) {
        float w = 1.f;
        if (t.kind == ITEM_XR) {
            int c0, d0, j0, c1, d1, j1;
            cr_decode_x(P.H.PXR.data(), ITEM_XR, P.H.xo_diag, t.e0, n, dB, c0, d0, j0);
            cr_decode_x(P.H.PXR.data(), ITEM_XR, P.H.xo_diag, t.e0 + t.ne - 1, n, dB, c1, d1, j1);
            int jm = 1;
            for (int cc = c0; cc <= c1; ++cc) jm = std::max(jm, cc & 7);
            w = 0.14f + 0.86f * (float)jm / 8.f;
        }
        P.xcost.push_back(w);
    }
    int mx = 1;
    for (auto* v : {&P.xt, &P.yt}) for (auto& t : *v) mx = std::max(mx, t.rcount[0] + t.rcount[1] + t.rcount[2]);
    P.max_rows = mx;
}

// device half: upload a host plan (the stream is drained first: a running kernel may still read the previous tables)
int upload_plan(qs_ctx* c, const HostPlan& P) {
    const size_t row_bytes = (size_t)c->n_pad * 2;
    if ((size_t)P.max_rows * row_bytes * 2 + CR_SMEM_HEADER + 256 + 1024 > (size_t)c->smem_optin / cr_ctas_per_sm(P.threads)) QS_FAIL(c, QS_E_UNSUPPORTED, "%d taxa: matrix rows of %zu bytes are too long for the counting kernel's shared-memory pipeline", c->n, row_bytes);
    const std::vector<RowTask>&xt = P.xt, &yt = P.yt;
    if (xt.size() + yt.size() > 0x3fffffffull) QS_FAIL(c, QS_E_UNSUPPORTED, "task count overflow");
    const size_t total = xt.size() + yt.size();
    int r;
    if (total > c->tasks_cap) {
        if ((r = dev_alloc(c, &c->d_tasks, total + total / 4))) { c->tasks_cap = 0; return r; }
        c->tasks_cap = total + total / 4;
    }
    const size_t enum_need = (size_t)5 * (c->n + 1) + P.H.PZ.size();
    if (enum_need > c->enum_cap) {
        if ((r = dev_alloc(c, &c->d_enum, enum_need + enum_need / 4))) { c->enum_cap = 0; return r; }
        c->enum_cap = enum_need + enum_need / 4;
    }
    QS_CUDA(c, cudaStreamSynchronize(c->stream));
    if (!xt.empty()) QS_CUDA(c, cudaMemcpy(c->d_tasks, xt.data(), xt.size() * sizeof(RowTask), cudaMemcpyHostToDevice));
    if (!yt.empty()) QS_CUDA(c, cudaMemcpy(c->d_tasks + xt.size(), yt.data(), yt.size() * sizeof(RowTask), cudaMemcpyHostToDevice));
    const size_t np1 = (size_t)c->n + 1;
    QS_CUDA(c, cudaMemcpy(c->d_enum, P.H.PXO.data(), np1 * 8, cudaMemcpyHostToDevice));
    QS_CUDA(c, cudaMemcpy(c->d_enum + np1, P.H.PXD.data(), np1 * 8, cudaMemcpyHostToDevice));
    QS_CUDA(c, cudaMemcpy(c->d_enum + 2 * np1, P.H.PY.data(), np1 * 8, cudaMemcpyHostToDevice));
    QS_CUDA(c, cudaMemcpy(c->d_enum + 3 * np1, P.H.CD.data(), np1 * 8, cudaMemcpyHostToDevice));
    QS_CUDA(c, cudaMemcpy(c->d_enum + 4 * np1, P.H.PXR.data(), np1 * 8, cudaMemcpyHostToDevice));
    QS_CUDA(c, cudaMemcpy(c->d_enum + 5 * np1, P.H.PZ.data(), P.H.PZ.size() * 8, cudaMemcpyHostToDevice));
    c->plan_zF = P.H.z_first; c->plan_zL = P.H.z_last;
    c->plan_dB = P.dB; c->plan_dE = P.dE; c->plan_with_y = P.with_y; c->plan_threads = P.threads; c->plan_nx = (int)xt.size(); c->plan_ny = (int)yt.size(); c->plan_max_rows = P.max_rows;
    c->plan_xcost = P.xcost; c->plan_ycost = P.ycost; ++c->plan_generation;
    return QS_OK;
}

// Role-Y tasks are only needed when some gene tree is class B.  A table-free run plans a task table per slab and per step
// (up to ~1e6 tasks at n = 2000), so it reads the class split back first and leaves role Y out when every tree is class A;
// a table context plans once, always with role Y (the kernel skips the class-B task space when |B| = 0), and never waits
// for the class split.
bool plan_needs_y(const qs_ctx* c) { return c->mode != QS_MODE_TABLE_FREE || c->n_class_a < c->m; }

int ensure_plan(qs_ctx* c, int dB, int dE, const HostPlan* ready = nullptr) {
    const int with_y = plan_needs_y(c) ? 1 : 0, threads = cr_threads_for(c->n);
    if (c->plan_dB == dB && c->plan_dE == dE && c->plan_with_y >= with_y && c->plan_threads == threads) return QS_OK;
    if (ready && ready->dB == dB && ready->dE == dE && ready->with_y >= with_y && ready->threads == threads) return upload_plan(c, *ready);
    HostPlan P;
    build_host_plan(c->n, c->n_pad, dB, dE, with_y != 0, threads, P);
    return upload_plan(c, P);
}

template <typename CINT>
void launch_init(qs_ctx* c, void* table, uint64_t n_entries) {
    int blocks = (int)std::min<uint64_t>((n_entries * 3 + 255) / 256, (uint64_t)c->num_sms * 16);
    if (blocks < 1) blocks = 1;
    qs_table_init_kernel<CINT><<<blocks, 256, 0, c->stream>>>((CINT*)table, n_entries, c->d_nA, (int)c->m);
    c->launches++;
}

template <typename CINT>
void launch_finalize(qs_ctx* c, void* table, uint64_t n_entries) {
    int blocks = (int)std::min<uint64_t>((n_entries + 255) / 256, (uint64_t)c->num_sms * 16);
    if (blocks < 1) blocks = 1;
    qs_table_finalize_kernel<CINT><<<blocks, 256, 0, c->stream>>>((CINT*)table, n_entries, c->d_nA, (int)c->m);
    c->launches++;
}

// count the quartets with d in [dB, dE) into `table` (rank_base = C(dB,4)); the distance matrices must be built
int run_count_rows(qs_ctx* c, int dB, int dE, void* table, const HostPlan* ready = nullptr) {
    const uint64_t rb = binom4((uint64_t)dB), re = binom4((uint64_t)dE);
    const uint64_t nq = re - rb;
    if (nq == 0) return QS_OK;
    int r;
    if ((r = ensure_plan(c, dB, dE, ready))) return r;
    switch (c->cint_bytes) {
        case 1: launch_init<uint8_t>(c, table, nq); break;
        case 2: launch_init<uint16_t>(c, table, nq); break;
        case 4: launch_init<uint32_t>(c, table, nq); break;
        default: launch_init<unsigned long long>(c, table, nq); break;
    }
    QS_CUDA(c, cudaMemsetAsync(c->d_counter, 0, sizeof(int), c->stream));
    CountRowsArgs a;
    a.D = c->d_D; a.order = c->d_order; a.n_class_a = c->d_nA;
    a.tasks = c->d_tasks; a.n_x = c->plan_nx; a.n_y = c->plan_ny;
    const size_t np1 = (size_t)c->n + 1;
    a.E = EnumTables{c->d_enum, c->d_enum + np1, c->d_enum + 2 * np1, c->d_enum + 3 * np1, c->d_enum + 4 * np1, c->d_enum + 5 * np1, c->plan_zF, c->plan_zL, use_xo_diag(c->n)};
    a.task_counter = c->d_counter; a.table = table; a.cint_bytes = c->cint_bytes; a.rank_base = rb;
    a.n = c->n; a.n_pad = c->n_pad; a.m = (int)c->m; a.d_begin = dB; a.d_end = dE;
    a.row_bytes = (uint32_t)c->n_pad * 2u;
    // staging ring: all the shared memory the CTA can get; every task sizes its own stages from the rows it touches
    // (kernels/count_rows.cuh: stream_rows), the largest task must fit twice
    const int threads = c->plan_threads, ctas = cr_ctas_per_sm(threads);
    const size_t budget = ((size_t)c->smem_optin / ctas - CR_SMEM_HEADER - 256 - (ctas > 1 ? 1024 : 0)) & ~(size_t)127;
    const size_t max_slot = (size_t)c->plan_max_rows * a.row_bytes;
    if (2 * max_slot > budget) QS_FAIL(c, QS_E_UNSUPPORTED, "%d taxa: two pipeline stages of %zu-byte row slots do not fit in shared memory", c->n, max_slot);
    a.ring_bytes = (uint32_t)budget;
    a.max_tps = CR_MAX_TPS; a.max_stages = CR_MAX_STAGES;
    if (const char* env = getenv("QS_MAX_TPS")) a.max_tps = std::max(1, std::min(32, atoi(env)));             // tuning hooks
    if (const char* env = getenv("QS_MAX_STAGES")) a.max_stages = std::max(2, std::min((int)CR_MAX_STAGES, atoi(env)));
    // Tree chunks: a task runs over <= QS_MAX_CHUNK_TREES = 4096 trees (the chunk's tree ids live in shared memory).  Tasks differ in
    // length (ragged-block tasks are shorter) and there are only a few per SM, so how well the last ones fill the machine decides
    // several per cent: the chunk count is the one whose greedy schedule — tasks handed out in table order to the first free
    // CTA, exactly what the kernel's atomic counter does — ends earliest, with ~40 trees' worth of flush and pipeline start per task (measured: 13 us per task at cfg2, profiles/r02_j_count_variants.txt).
    // The class split is the previous run's (a hint: it only affects the choice, never the result).
    const int64_t mA_hint = c->counted_once ? std::min<int64_t>(c->n_class_a, c->m) : 0, mB_hint = c->m - mA_hint;
    const int64_t k_min = std::max<int64_t>(1, (c->m + QS_MAX_CHUNK_TREES - 1) / QS_MAX_CHUNK_TREES);
    int64_t best_k = k_min;
    const int64_t ck[4] = {c->plan_generation, c->m, mA_hint, (int64_t)c->num_sms * ctas};
    if (ck[0] == c->chunk_key[0] && ck[1] == c->chunk_key[1] && ck[2] == c->chunk_key[2] && ck[3] == c->chunk_key[3]) best_k = c->chunk_choice;
    else {                                                            // (host work: done once per plan and class split, not per step)
        const int slots = c->num_sms * ctas;
        const size_t n_tasks = c->plan_xcost.size() + (mB_hint > 0 ? c->plan_ycost.size() : 0);
        double best_t = -1;
        std::vector<double> heap;
        for (int64_t k = k_min; k <= std::max<int64_t>(k_min, c->m / 256) && k <= k_min + 24; ++k) {
            if ((double)n_tasks * (double)k > 4e5) break;            // thousands of tasks per CTA: the tail no longer matters
            const int64_t ct = std::max<int64_t>(1, (c->m + k - 1) / k);
            heap.assign(slots, 0.0);                                  // min-heap of the CTAs' finish times
            auto run = [&](double cost) {
                std::pop_heap(heap.begin(), heap.end(), std::greater<double>());
                heap.back() += cost;
                std::push_heap(heap.begin(), heap.end(), std::greater<double>());
            };
            auto run_class = [&](int64_t len, bool with_y_tasks) {    // kernel order: chunk-major, the tasks of a chunk in table order
                if (len <= 0) return;
                const int64_t nch = (len + ct - 1) / ct, per = (len + nch - 1) / nch;
                for (int64_t ch = 0; ch < nch; ++ch) {
                    const double trees = (double)std::min(per, len - ch * per);
                    for (float w : c->plan_xcost) run((double)w * trees + 40.0);
                    if (with_y_tasks) for (float w : c->plan_ycost) run((double)w * trees + 40.0);
                }
            };
            run_class(mA_hint, false);
            run_class(mB_hint, true);
            const double t_end = *std::max_element(heap.begin(), heap.end());
            if (best_t < 0 || t_end < best_t * 0.995) { best_t = t_end; best_k = k; }      // fewer chunks unless clearly earlier
        }
        for (int i = 0; i < 4; ++i) c->chunk_key[i] = ck[i];
        c->chunk_choice = best_k;
    }
    if (const char* env = getenv("QS_CHUNK_COUNT")) {                  // tuning hook: force the number of tree chunks
        const int64_t k = atoll(env);
        if (k >= 1) best_k = std::max<int64_t>(k, (c->m + QS_MAX_CHUNK_TREES - 1) / QS_MAX_CHUNK_TREES);
    }
    a.chunk_trees = (int)std::min<int64_t>(QS_MAX_CHUNK_TREES, std::max<int64_t>(1, (c->m + best_k - 1) / best_k));
    if (getenv("QS_DEBUG_PLAN")) fprintf(stderr, "[qscuda] plan d=[%d,%d) n_x=%d n_y=%d chunks=%lld chunk_trees=%d threads=%d mA_hint=%lld\n", dB, dE, a.n_x, a.n_y, (long long)best_k, a.chunk_trees, threads, (long long)mA_hint);
    const size_t smem = CR_SMEM_HEADER + budget;
    if (threads == CR_THREADS_BIG) QS_CUDA(c, cudaFuncSetAttribute(qs_count_rows_kernel<CR_THREADS_BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else QS_CUDA(c, cudaFuncSetAttribute(qs_count_rows_kernel<CR_THREADS_SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t max_tasks = ((int64_t)a.n_x + a.n_y) * 2 * ((c->m + a.chunk_trees - 1) / a.chunk_trees + 1);
    if (max_tasks > 0x7fffffffLL) QS_FAIL(c, QS_E_UNSUPPORTED, "task count overflow");
    const int grid = (int)std::min<int64_t>(max_tasks, (int64_t)c->num_sms * ctas);
    if (threads == CR_THREADS_BIG) qs_count_rows_kernel<CR_THREADS_BIG><<<grid, CR_THREADS_BIG, smem, c->stream>>>(a);
    else qs_count_rows_kernel<CR_THREADS_SMALL><<<grid, CR_THREADS_SMALL, smem, c->stream>>>(a);
    c->launches++;
    QS_CUDA(c, cudaGetLastError());
    switch (c->cint_bytes) {
        case 1: launch_finalize<uint8_t>(c, table, nq); break;
        case 2: launch_finalize<uint16_t>(c, table, nq); break;
        case 4: launch_finalize<uint32_t>(c, table, nq); break;
        default: launch_finalize<unsigned long long>(c, table, nq); break;
    }
    QS_CUDA(c, cudaGetLastError());
    return QS_OK;
}

int build_reference(qs_ctx* c, int n_nodes, const int32_t* parent, const int32_t* parent_edge, const int32_t* leaf_id,
                    const int32_t* first_child, const int32_t* next_sibling) {
    HostRef& R = c->ref;
    R = HostRef();
    if (n_nodes < 5 || n_nodes > 65535) QS_FAIL(c, QS_E_REFERENCE, "reference tree has %d nodes (need 5..65535)", n_nodes);
    R.n_nodes = n_nodes;
    R.parent.assign(parent, parent + n_nodes); R.parent_edge.assign(parent_edge, parent_edge + n_nodes);
    R.leaf_id.assign(leaf_id, leaf_id + n_nodes); R.first_child.assign(first_child, first_child + n_nodes);
    R.next_sib.assign(next_sibling, next_sibling + n_nodes);
    R.depth.assign(n_nodes, 0);
    if (parent[0] != -1) QS_FAIL(c, QS_E_REFERENCE, "node 0 must be the root");
    std::vector<int> nchild(n_nodes, 0);
    for (int i = 1; i < n_nodes; ++i) {
        if (parent[i] < 0 || parent[i] >= i) QS_FAIL(c, QS_E_REFERENCE, "parent[%d] = %d violates parent < child numbering", i, parent[i]);
        if (parent_edge[i] < 0 || parent_edge[i] >= n_nodes - 1) QS_FAIL(c, QS_E_REFERENCE, "parent_edge[%d] out of range", i);
        R.depth[i] = R.depth[parent[i]] + 1;
        nchild[parent[i]]++;
    }
    // child lists must be consistent with parent[]
    for (int v = 0; v < n_nodes; ++v) {
        int cnt = 0;
        for (int ch = first_child[v]; ch != -1; ch = next_sibling[ch]) {
            if (ch <= v || ch >= n_nodes || parent[ch] != v || ++cnt > nchild[v]) QS_FAIL(c, QS_E_REFERENCE, "child list of node %d inconsistent with parent[]", v);
        }
        if (cnt != nchild[v]) QS_FAIL(c, QS_E_REFERENCE, "child list of node %d inconsistent with parent[]", v);
        if ((nchild[v] == 0) != (leaf_id[v] >= 0)) QS_FAIL(c, QS_E_REFERENCE, "leaf_lookup_id must be >= 0 exactly for leaves (node %d)", v);
    }
    // planar (Euler tour) leaf order must be 0,1,2,...  (QuartetCounterLookup.hpp:249-258)
    const int n = c->n;
    R.leaf_node.assign(n, -1);
    std::vector<int> lo(n_nodes, 0), hi(n_nodes, 0);
    {
        int k = 0;
        std::vector<int> stack{0}, it(n_nodes);
        for (int v = 0; v < n_nodes; ++v) it[v] = first_child[v];
        lo[0] = 0;
        while (!stack.empty()) {
            int v = stack.back();
            int ch = it[v];
            if (ch == -1) { hi[v] = k; stack.pop_back(); continue; }
            it[v] = next_sibling[ch];
            lo[ch] = k;
            if (first_child[ch] == -1) {
                if (k >= n || leaf_id[ch] != k) QS_FAIL(c, QS_E_REFERENCE, "lookup ids must follow the reference tree's Euler-tour leaf order (leaf %d has id %d)", k, leaf_id[ch]);
                R.leaf_node[k] = ch; ++k; hi[ch] = k;
            } else stack.push_back(ch);
        }
        if (k != n) QS_FAIL(c, QS_E_REFERENCE, "reference tree has %d leaves, context was created for %d taxa", k, n);
    }
    // genesis is_bifurcating: max rank (links - 1) == 2 (genesis tree/function/functions.cpp:57-69)
    int max_rank = 0; bool all3 = true;
    R.inner_index.assign(n_nodes, -1);
    // inner index = rank of the node's first gap in the planar leaf order (gap g lies between leaves g and g+1 and belongs
    // to lca(g, g+1); the first gap of v is the one after its first child).  Gaps of different nodes differ, so for leaves
    // x < y the index of lca(x,y) is <= its first gap < y: the scan kernel's CTA of (c,d) only touches accumulator
    // indices below c (kernels/score.cuh).  Nodes with a single child own no gap, are never an LCA and come last.
    {
        std::vector<std::pair<int, int>> order;                       // (first gap, node)
        int unary = 0;
        for (int v = 0; v < n_nodes; ++v) {
            const int links = nchild[v] + (v != 0 ? 1 : 0);
            max_rank = std::max(max_rank, links - 1);
            if (nchild[v] == 0) continue;
            if (links != 3) all3 = false;
            if (nchild[v] >= 2) order.emplace_back(hi[first_child[v]] - 1, v);
            else order.emplace_back(n + (unary++), v);
        }
        std::sort(order.begin(), order.end());
        for (auto& o : order) { R.inner_index[o.second] = (int)R.inner_node.size(); R.inner_node.push_back(o.second); R.inner_gap.push_back(o.first); }
    }
    R.n_inner = (int)R.inner_node.size();
    R.bifurcating = (max_rank == 2);
    if (R.bifurcating && !all3)
        QS_FAIL(c, QS_E_REFERENCE, "reference tree passes is_bifurcating but has an inner node of degree 2 (e.g. a rooted tree); the reference mis-scores such trees (SURVEY.md App. B6) - unroot it");
    // LCA of every leaf pair (inner index) and inner depths
    R.lca.assign((size_t)n * n, 0);
    R.idepth.resize(R.n_inner);
    for (int v = 0; v < n_nodes; ++v) {
        if (nchild[v] == 0) continue;
        const uint16_t iv = (uint16_t)R.inner_index[v];
        R.idepth[iv] = (uint16_t)R.depth[v];
        for (int c1 = first_child[v]; c1 != -1; c1 = next_sibling[c1])
            for (int c2 = next_sibling[c1]; c2 != -1; c2 = next_sibling[c2])
                for (int x = lo[c1]; x < hi[c1]; ++x)
                    for (int y = lo[c2]; y < hi[c2]; ++y) { R.lca[(size_t)x * n + y] = iv; R.lca[(size_t)y * n + x] = iv; }
    }
    R.inner_parent.assign(R.n_inner, -1);
    for (int i = 0; i < R.n_inner; ++i) { const int v = R.inner_node[i]; if (v != 0) R.inner_parent[i] = R.inner_index[parent[v]]; }
    R.leaf_parent.assign(n, -1);
    for (int x = 0; x < n; ++x) R.leaf_parent[x] = R.inner_index[parent[R.leaf_node[x]]];
    return QS_OK;
}

int ensure_pair_arrays(qs_ctx* c) {
    const size_t I = c->ref.n_inner, E = (size_t)c->ref.n_nodes - 1;
    int r;
    if (!c->d_pair_sums && (r = dev_alloc(c, &c->d_pair_sums, I * I * 3))) return r;
    if (!c->d_pair_best && (r = dev_alloc(c, &c->d_pair_best, I * I))) return r;
    if (!c->d_pair_score && (r = dev_alloc(c, &c->d_pair_score, I * I))) return r;
    if (!c->d_edge && (r = dev_alloc(c, &c->d_edge, 4 * E))) return r;
    if (!c->d_edge_out && (r = dev_alloc(c, &c->d_edge_out, 7 * E))) return r;
    if (!c->d_scan_counter && (r = dev_alloc(c, &c->d_scan_counter, 1))) return r;
    return QS_OK;
}

int fill_i64(qs_ctx* c, long long* p, size_t count, long long v) {
    qs_fill_i64_kernel<<<(unsigned)std::max<size_t>(1, std::min<size_t>((count + 255) / 256, 4096)), 256, 0, c->stream>>>(p, count, v);
    c->launches++;
    QS_CUDA(c, cudaGetLastError());
    return QS_OK;
}

int clear_pair_arrays(qs_ctx* c) {
    const size_t I = c->ref.n_inner;
    QS_CUDA(c, cudaMemsetAsync(c->d_pair_sums, 0, I * I * 3 * 8, c->stream));
    int r;
    if ((r = fill_i64(c, c->d_pair_best, I * I, QS_I64_NONE))) return r;
    return fill_i64(c, c->d_pair_score, I * I, QS_I64_NONE);
}

// CTA shape of the scan: a CTA owns one (c,d), lanes take consecutive table entries.  512 threads (two CTAs per SM) while the
// per-CTA accumulators (36 bytes per slot, 2n + 496 slots) fit twice beside the staging rings, 1024 threads above that; for
// very wide references (n > ~2,300) the accumulators move to global memory.
template <typename CINT, int THREADS>
int launch_scan_t(qs_ctx* c, ScoreArgs& a, bool smem_acc, bool carry) {
    const size_t ring = scan_ring_bytes(THREADS, (int)sizeof(CINT)), acc = scan_acc_bytes(c->n, a.q4_levels, carry);
    const size_t smem = ring + (smem_acc ? acc : 0);
    int per_sm = std::max(1, std::min(THREADS <= 512 ? 2 : 1, (int)((size_t)c->smem_optin / (smem + 1024))));
    const int grid = (int)std::max<long long>(1, std::min<long long>(a.n_items, (long long)c->num_sms * per_sm));
    if (!smem_acc) {
        const size_t need = (size_t)grid * acc;
        if (need > c->scan_scratch_bytes) {
            int r = dev_alloc(c, &c->d_scan_scratch, need / 8 + 1);
            if (r) { c->scan_scratch_bytes = 0; return r; }
            c->scan_scratch_bytes = need;
        }
        a.scratch = c->d_scan_scratch;
    }
    QS_CUDA(c, cudaMemsetAsync(c->d_scan_counter, 0, sizeof(int), c->stream));
    auto go = [&](auto kernel) -> int {
        QS_CUDA(c, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kernel<<<grid, THREADS + 32, smem, c->stream>>>(a);                       // + the producer warp
        return QS_OK;
    };
    int rr;
    if (smem_acc) rr = carry ? go(qs_scan_kernel<CINT, THREADS, true, true>) : go(qs_scan_kernel<CINT, THREADS, true, false>);
    else rr = carry ? go(qs_scan_kernel<CINT, THREADS, false, true>) : go(qs_scan_kernel<CINT, THREADS, false, false>);
    if (rr) return rr;
    c->launches++;
    QS_CUDA(c, cudaGetLastError());
    return QS_OK;
}

template <typename CINT>
int launch_scan(qs_ctx* c, ScoreArgs& a) {
    // can a 32-bit sum overflow inside one (c,d)?  largest stored count x the C(c,2) quartets of the widest pair
    const uint64_t max_count = std::min<uint64_t>((uint64_t)c->m * (uint64_t)a.count_scale, a.cint_mask);
    const bool carry = max_count * binom2((uint64_t)std::max(2, a.d_end - 1)) >= (1ull << 32) || getenv("QS_SCAN_CARRY") != nullptr;      // (env: test hook)
    const size_t optin = (size_t)c->smem_optin;
    const bool force_global = getenv("QS_SCAN_GLOBAL_ACC") != nullptr;                    // test hook: the large-n path on a small input
    auto fits = [&](int threads, int copies) { return copies * (scan_ring_bytes(threads, (int)sizeof(CINT)) + scan_acc_bytes(c->n, a.q4_levels, carry) + 1024) <= optin; };
    // accQ shrinks before the accumulators leave shared memory: quartets deeper than its levels take the global-memory path
    // accQ shrinks (down to 40 levels) before the small CTA shape stops fitting twice per SM, and further before the accumulators
    // leave shared memory: quartets deeper than its levels take the global-memory path
    const int full_levels = a.q4_levels;
    while (a.q4_levels > 40 && !fits(352, 2)) a.q4_levels -= 4;
    if (!fits(352, 2)) a.q4_levels = full_levels;
    while (a.q4_levels > 16 && !fits(992, 1)) a.q4_levels = std::max(16, a.q4_levels / 2);
    int threads = fits(352, 2) ? 352 : 1024;          // (352 consumers + the producer warp) x 2 CTAs per SM: 85 registers per thread (512 x 2 caps them at 64 and spills: 30 vs 20 ms at n = 500)
    if (const char* env = getenv("QS_SCAN_THREADS")) { const int t = atoi(env); if (t == 352 || t == 512 || t == 1024) threads = t; }    // tuning / test hook
    if (threads == 352) return launch_scan_t<CINT, 352>(c, a, !force_global && fits(352, 1), carry);
    if (threads == 512) return launch_scan_t<CINT, 512>(c, a, !force_global && fits(512, 1), carry);
    return launch_scan_t<CINT, 992>(c, a, !force_global && fits(992, 1), carry);                  // (992 consumers + the producer warp = 1024 threads)
}

// work items of the scan for the d-range [dB, dE): (c; d0 <= d < d1) with lca(c,d) constant over the run of d, cut so that no
// item exceeds 1/16 of a CTA's share of the work, largest first (the kernel hands them out through an atomic counter)
int build_scan_items(qs_ctx* c, int dB, int dE) {
    if (c->scan_dB == dB && c->scan_dE == dE && c->d_scan_items) return QS_OK;
    const HostRef& R = c->ref;
    const int n = c->n;
    struct Item { int cc, d0, d1; double cost; };
    std::vector<Item> items;
    double total = 0;
    for (int cc = 2; cc < dE - 1 + 1 && cc < n - 1; ++cc) {
        const int dlo = std::max(cc + 1, dB);
        for (int d = dlo; d < dE;) {
            const uint16_t r = R.lca[(size_t)cc * n + d];
            int e = d + 1;
            while (e < dE && R.lca[(size_t)cc * n + e] == r) ++e;
            const double cost = (double)cc * (cc - 1) / 2 * (e - d);
            items.push_back({cc, d, e, cost});
            total += cost;
            d = e;
        }
    }
    const double cap = std::max(4096.0, total / ((double)c->num_sms * 2 * 16));
    std::vector<int4> out;
    std::vector<std::pair<double, size_t>> order;
    for (auto& it : items) {
        const double per_d = it.cost / (it.d1 - it.d0);
        const int step = (int)std::max(1.0, std::floor(cap / per_d));
        for (int d = it.d0; d < it.d1; d += step) {
            const int e = std::min(it.d1, d + step);
            order.emplace_back(-(per_d * (e - d)), out.size());
            out.push_back(make_int4(it.cc, d, e, 0));
        }
    }
    std::sort(order.begin(), order.end());
    std::vector<int4> sorted(out.size());
    for (size_t i = 0; i < order.size(); ++i) sorted[i] = out[order[i].second];
    if (sorted.size() > 0x7fffffffull) QS_FAIL(c, QS_E_UNSUPPORTED, "scan work-item count overflow");
    if (sorted.size() > c->scan_items_cap) {
        int r = dev_alloc(c, &c->d_scan_items, sorted.size() + sorted.size() / 4 + 1);
        if (r) { c->scan_items_cap = 0; return r; }
        c->scan_items_cap = sorted.size() + sorted.size() / 4 + 1;
    }
    QS_CUDA(c, cudaStreamSynchronize(c->stream));                      // a running scan may still read the previous list
    if (!sorted.empty()) QS_CUDA(c, cudaMemcpy(c->d_scan_items, sorted.data(), sorted.size() * sizeof(int4), cudaMemcpyHostToDevice));
    c->scan_n_items = (int64_t)sorted.size();
    c->scan_dB = dB; c->scan_dE = dE;
    return QS_OK;
}

// scan a table holding the quartets with d in [dB, dE) and accumulate into the per-pair partials on the device
int scan_table(qs_ctx* c, const void* table, int dB, int dE, int count_scale) {
    ScoreArgs a{};
    int r0;
    if ((r0 = build_scan_items(c, dB, dE))) return r0;
    a.items = c->d_scan_items; a.n_items = c->scan_n_items;
    if (a.n_items == 0) return QS_OK;
    a.table = table; a.rank_base = binom4((uint64_t)dB); a.lcapd = c->d_lcapd; a.idepth = c->d_idepth;
    // accQ levels: the reference tree's depth (no pair of ancestors of a leaf is further apart), capped by QS_Q4_MAX_LEVELS
    int max_depth = 2;
    for (uint16_t dd : c->ref.idepth) max_depth = std::max<int>(max_depth, dd + 1);
    a.q4_levels = std::min(max_depth, QS_Q4_MAX_LEVELS);
    if (const char* env = getenv("QS_SCAN_LEVELS")) a.q4_levels = std::max(2, std::min(QS_Q4_MAX_LEVELS, atoi(env)));       // test hook: force the deep (global-memory) path
    a.inner_parent = c->d_inner_parent; a.leaf_parent = c->d_leaf_parent; a.inner_gap = c->d_inner_gap;
    a.pair_sums = c->d_pair_sums; a.pair_best = c->d_pair_best; a.pair_score = c->d_pair_score; a.scratch = nullptr; a.work_counter = c->d_scan_counter;
    a.n = c->n; a.I = c->ref.n_inner;
    a.d_begin = dB; a.d_end = dE; a.count_scale = count_scale; a.cint_mask = cint_mask(c->cint_bytes);
    a.bifurcating = c->ref.bifurcating ? 1 : 0;
    switch (c->cint_bytes) {
        case 1: return launch_scan<uint8_t>(c, a);
        case 2: return launch_scan<uint16_t>(c, a);
        case 4: return launch_scan<uint32_t>(c, a);
        default: return launch_scan<unsigned long long>(c, a);
    }
}

// slabs of a table-free run: largest dE (8-aligned when possible, so that role Y's d-blocks are full) whose table fits the budget
std::vector<std::pair<int, int>> plan_slabs(const qs_ctx* c, size_t budget) {
    const size_t eb = 3 * (size_t)c->cint_bytes;
    std::vector<std::pair<int, int>> slabs;
    for (int dB = std::max(3, c->d_begin); dB < c->d_end;) {
        int dE = dB + 1;
        while (dE < c->d_end && (binom4((uint64_t)dE + 1) - binom4((uint64_t)dB)) * eb <= budget) ++dE;
        if (dE < c->d_end && dE - dB > 8) dE = std::max(dB + 1, dE & ~7);
        slabs.emplace_back(dB, dE);
        dB = dE;
    }
    return slabs;
}

size_t slab_budget(const qs_ctx* c) {
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); free_b = (size_t)8 << 30; }
    size_t budget = (free_b + c->table_bytes) / 10 * 6;                 // leave room for the distance matrices of a later, larger run
    if (const char* env = getenv("QS_SLAB_BYTES")) budget = (size_t)strtoull(env, nullptr, 10);   // test hook: force several slabs
    return budget;
}

int ensure_slab_table(qs_ctx* c, size_t need, int dB, int dE) {
    if (need <= c->table_bytes) return QS_OK;
    if (c->d_table) { cudaFree(c->d_table); c->d_table = nullptr; c->table_bytes = 0; }
    cudaError_t e = cudaMalloc(&c->d_table, need + kTableSlack);
    if (e != cudaSuccess) {
        cudaGetLastError();
        QS_FAIL(c, QS_E_MEMORY, "Insufficient memory! a slab of %zu bytes (d in [%d,%d)) does not fit this device", need, dB, dE);
    }
    c->table_bytes = need;
    return QS_OK;
}

// QS_MODE_TABLE_FREE (the -s analogue): the shard's d-range is processed in slabs that fit the device; each slab is
// counted into a temporary table, scanned into the per-pair partials and discarded
int run_table_free(qs_ctx* c) {
    if (!c->has_ref) QS_FAIL(c, QS_E_STATE, "a table-free context scores while it counts: call qs_set_reference before qs_count");
    int r;
    if ((r = ensure_pair_arrays(c))) return r;
    if ((r = clear_pair_arrays(c))) return r;
    const size_t eb = 3 * (size_t)c->cint_bytes;
    const std::vector<std::pair<int, int>> slabs = plan_slabs(c, slab_budget(c));
    // the host plan of slab k+1 (task table: up to ~1e6 tasks at n = 2000) is built by a helper thread while the GPU counts slab k
    const bool with_y = plan_needs_y(c);
    HostPlan plans[2];
    std::thread planner;
    auto start_plan = [&](size_t k) {
        if (k >= slabs.size()) return;
        HostPlan* P = &plans[k & 1];
        const int n = c->n, n_pad = c->n_pad, b0 = slabs[k].first, e0 = slabs[k].second;
        const int threads = cr_threads_for(n);
        planner = std::thread([=]() { build_host_plan(n, n_pad, b0, e0, with_y, threads, *P); });
    };
    start_plan(0);
    int rc = QS_OK;
    for (size_t k = 0; k < slabs.size() && rc == QS_OK; ++k) {
        const int dB = slabs[k].first, dE = slabs[k].second;
        planner.join();
        start_plan(k + 1);
        const size_t need = (size_t)(binom4((uint64_t)dE) - binom4((uint64_t)dB)) * eb;
        if ((rc = ensure_slab_table(c, need, dB, dE))) break;
        if ((rc = run_count_rows(c, dB, dE, c->d_table, &plans[k & 1]))) break;
        rc = scan_table(c, c->d_table, dB, dE, c->fused_scale);
    }
    if (planner.joinable()) planner.join();
    if (rc) return rc;
    c->fused_partials_valid = true;
    return QS_OK;
}

// canonical counts of the ranks [lo, hi) of this shard into host memory.  Table contexts copy them; a table-free context
// keeps no table, so the slabs that cover the range are counted again (the distance matrices are still on the device).
int counts_to_host(qs_ctx* c, uint64_t lo, uint64_t hi, void* out) {
    const size_t eb = 3 * (size_t)c->cint_bytes;
    if (lo >= hi) return QS_OK;
    if (c->mode == QS_MODE_TABLE) {
        QS_CUDA(c, cudaMemcpyAsync(out, (const char*)c->d_table + (lo - c->rank_begin) * eb, (hi - lo) * eb, cudaMemcpyDeviceToHost, c->stream));
        QS_CUDA(c, cudaStreamSynchronize(c->stream));
        return QS_OK;
    }
    if (!c->dist_valid) QS_FAIL(c, QS_E_STATE, "table-free context: the distance matrices are gone, call qs_count again");
    int r;
    for (auto& sl : plan_slabs(c, slab_budget(c))) {
        const uint64_t rb = binom4((uint64_t)sl.first), re = binom4((uint64_t)sl.second);
        const uint64_t a = std::max(lo, rb), b = std::min(hi, re);
        if (a >= b) continue;
        if ((r = ensure_slab_table(c, (size_t)(re - rb) * eb, sl.first, sl.second))) return r;
        if ((r = run_count_rows(c, sl.first, sl.second, c->d_table))) return r;
        QS_CUDA(c, cudaMemcpyAsync((char*)out + (a - lo) * eb, (const char*)c->d_table + (a - rb) * eb, (b - a) * eb, cudaMemcpyDeviceToHost, c->stream));
        QS_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    return QS_OK;
}

// QS_MODE_TABLE: scan the resident table into the device partials; QS_MODE_TABLE_FREE: qs_count has already done it.
// Asynchronous: nothing is copied to the host and the stream is not drained.
int run_score_scan(qs_ctx* c, int count_scale) {
    if (!c->has_ref) QS_FAIL(c, QS_E_STATE, "qs_set_reference has not been called");
    if (!c->counted) QS_FAIL(c, QS_E_STATE, "qs_count has not been called");
    if (count_scale != 1 && count_scale != 2) QS_FAIL(c, QS_E_ARG, "count_scale must be 1 or 2");
    if ((uint64_t)c->m * (uint64_t)count_scale >= (1ull << QS_TRIPLE_BITS) && c->cint_bytes >= 4)
        QS_FAIL(c, QS_E_UNSUPPORTED, "%lld trees x count_scale %d: the LQ-IC selection packs counts into %d bits (fewer than %llu trees)", (long long)c->m, count_scale,
                QS_TRIPLE_BITS, (1ull << QS_TRIPLE_BITS) / (unsigned)count_scale);
    int r;
    c->partials_ready = false;
    if (c->mode == QS_MODE_TABLE_FREE) {
        if (!c->fused_partials_valid || c->fused_scale != count_scale)
            QS_FAIL(c, QS_E_STATE, "table-free context: partials were accumulated by qs_count with count_scale=%d", c->fused_scale);
    } else {
        if ((r = ensure_pair_arrays(c))) return r;
        QS_CUDA(c, cudaEventRecord(c->ev[4], c->stream));
        if ((r = clear_pair_arrays(c))) return r;
        if ((r = scan_table(c, c->d_table, c->d_begin, c->d_end, count_scale))) return r;
        QS_CUDA(c, cudaEventRecord(c->ev[5], c->stream));
        c->score_timed = true;
    }
    c->partials_ready = true;
    return QS_OK;
}

// per-edge reduction of the (possibly all-reduced) device partials + host evaluation of the selected triples / sums with the
// host's libm in the reference's operation order.  Any of the outputs may be null.
int run_edge_reduce(qs_ctx* c, int exact_qp, double* lqic, double* qpic, double* eqpic) {
    if (!c->partials_ready) QS_FAIL(c, QS_E_STATE, "no scanned partials: call qs_score_scan (or qs_score) first");
    const HostRef& R = c->ref;
    const int E = R.n_nodes - 1, I = R.n_inner;
    int r;
    if ((r = fill_i64(c, c->d_edge, (size_t)4 * E, QS_I64_NONE))) return r;
    EdgeArgs a{};
    a.pair_sums = c->d_pair_sums; a.pair_best = c->d_pair_best; a.pair_score = c->d_pair_score;
    a.inner_node = c->d_inner_node; a.node_parent = c->d_node_parent; a.node_depth = c->d_node_depth; a.node_edge = c->d_node_edge; a.node_inner = c->d_node_inner;
    a.edge_lq = c->d_edge; a.edge_eqp = c->d_edge + E; a.edge_lq_arg = c->d_edge + 2 * (size_t)E; a.edge_eqp_arg = c->d_edge + 3 * (size_t)E;
    a.out = c->d_edge_out; a.I = I; a.E = E; a.n_nodes = R.n_nodes; a.bifurcating = R.bifurcating ? 1 : 0; a.exact_qp = exact_qp;
    const unsigned blocks = (unsigned)std::max<size_t>(1, std::min<size_t>(((size_t)I * I + 255) / 256, (size_t)c->num_sms * 8));
    qs_edge_reduce_kernel<1><<<blocks, 256, 0, c->stream>>>(a);
    qs_edge_reduce_kernel<2><<<blocks, 256, 0, c->stream>>>(a);
    qs_edge_gather_kernel<<<(unsigned)((R.n_nodes + 255) / 256), 256, 0, c->stream>>>(a);
    c->launches += 3;
    QS_CUDA(c, cudaGetLastError());
    c->h_edge_out.resize((size_t)E * 7);
    QS_CUDA(c, cudaMemcpyAsync(c->h_edge_out.data(), c->d_edge_out, (size_t)E * 7 * 8, cudaMemcpyDeviceToHost, c->stream));
    QS_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->score_timed) {
        float ms = 0;
        cudaEventElapsedTime(&ms, c->ev[4], c->ev[5]);
        c->score_ms = ms;
        c->score_timed = false;
    }
    const double inf = std::numeric_limits<double>::infinity();
    const uint64_t M = (1ull << QS_TRIPLE_BITS) - 1;
    for (int e = 0; e < E; ++e) {
        const unsigned long long* o = c->h_edge_out.data() + (size_t)e * 7;
        if (lqic) lqic[e] = o[0] == QS_TRIPLE_NONE ? inf : host_log_score(o[0] >> (2 * QS_TRIPLE_BITS), (o[0] >> QS_TRIPLE_BITS) & M, o[0] & M);
        for (int k = 0; k < 2; ++k) {
            double* out = k == 0 ? eqpic : qpic;
            if (!out) continue;
            const unsigned long long* p = o + 1 + 3 * k;
            if (!R.bifurcating || (p[0] == ~0ull && p[1] == ~0ull && p[2] == ~0ull)) { out[e] = inf; continue; }
            uint64_t p1 = p[0], p2 = p[1], p3 = p[2];
            if (!exact_qp) { p1 &= 0xffffffffull; p2 &= 0xffffffffull; p3 &= 0xffffffffull; }   // `unsigned p1,p2,p3` (QuartetScoreComputer.hpp:382)
            out[e] = host_log_score(p1, p2, p3);
        }
    }
    return QS_OK;
}

template <typename F>
void for_path_edges(const HostRef& R, int u, int v, F f) {
    while (u != v) {
        if (R.depth[u] >= R.depth[v]) { f(R.parent_edge[u]); u = R.parent[u]; }
        else { f(R.parent_edge[v]); v = R.parent[v]; }
    }
}

// The two host post-passes below visit every inner-node pair and walk its path (O(I^2 x depth): 5e5 pairs at
// n = 1000).  Rows of the pair matrix are dealt round-robin to host threads; every thread keeps private per-edge
// minima that are merged at the end, so the result does not depend on the thread count.
// A process-wide pool of parked host threads for these passes: at n = 100 a pass is only ~0.3 ms of work, so threads are
// woken (condition variable, ~10 us) rather than created per call.  Leaked on purpose: the workers are detached and
// must not be joined from a static destructor at library unload.
class HostPool {
public:
    static HostPool*& instance() { static HostPool* p = nullptr; return p; }
    static HostPool& get() {
        HostPool*& p = instance();
        if (!p) {
            p = new HostPool();
            // a fork()ed child has none of the parent's threads: start over with an empty pool there
            pthread_atfork(nullptr, nullptr, [] { instance() = new HostPool(); });
        }
        return *p;
    }
    // run job(w) for w in [0, nt): w = 0 on the calling thread, the rest on pool workers; returns when all are done
    void run(int nt, const std::function<void(int)>& job) {
        if (nt <= 1) { job(0); return; }
        std::unique_lock<std::mutex> run_lock(run_mu_);          // one parallel region at a time (contexts may share the pool)
        {
            std::lock_guard<std::mutex> g(mu_);
            while ((int)workers_ < nt - 1) { std::thread(&HostPool::worker, this, (int)workers_ + 1).detach(); ++workers_; }
            job_ = &job; n_active_ = nt; pending_ = nt - 1; ++generation_;
        }
        cv_.notify_all();
        job(0);
        std::unique_lock<std::mutex> g(mu_);
        done_cv_.wait(g, [&] { return pending_ == 0; });
        job_ = nullptr;
    }
private:
    void worker(int id) {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(int)>* job = nullptr;
            {
                std::unique_lock<std::mutex> g(mu_);
                cv_.wait(g, [&] { return generation_ != seen; });
                seen = generation_;
                if (id < n_active_) job = job_;
            }
            if (job) {
                (*job)(id);
                std::lock_guard<std::mutex> g(mu_);
                if (--pending_ == 0) done_cv_.notify_one();
            }
        }
    }
    std::mutex mu_, run_mu_;
    std::condition_variable cv_, done_cv_;
    const std::function<void(int)>* job_ = nullptr;
    uint64_t generation_ = 0;
    int n_active_ = 0, pending_ = 0;
    size_t workers_ = 0;
};

template <typename F>
void for_pair_rows_parallel(int I, int E, int n_arrays, double* const* out, F body) {
    const double inf = std::numeric_limits<double>::infinity();
    const int hw = (int)std::max(1u, std::thread::hardware_concurrency());
    int nt = (I < 48) ? 1 : std::min(I < 256 ? 8 : 16, hw);
    if (const char* env = getenv("QS_HOST_THREADS")) nt = std::max(1, std::min(64, atoi(env)));      // explicit host thread count (tests; -t)
    nt = std::max(1, std::min(nt, I));
    std::vector<std::vector<double>> local((size_t)nt * n_arrays, std::vector<double>((size_t)E, inf));
    HostPool::get().run(nt, [&](int w) {
        double* mine[2] = {nullptr, nullptr};
        for (int k = 0; k < n_arrays; ++k) mine[k] = local[(size_t)w * n_arrays + k].data();
        for (int iu = w; iu < I; iu += nt) body(iu, mine);
    });
    for (int k = 0; k < n_arrays; ++k) {
        if (!out[k]) continue;
        for (int e = 0; e < E; ++e) {
            double m = inf;
            for (int w = 0; w < nt; ++w) m = std::min(m, local[(size_t)w * n_arrays + k][e]);
            out[k][e] = m;
        }
    }
}

// QP-IC / EQP-IC from (reduced) pair sums: QuartetScoreComputer.hpp:472-489
void qp_from_pairs(const qs_ctx* c, const uint64_t* sums, int exact_qp, double* qpic, double* eqpic) {
    const HostRef& R = c->ref;
    const int I = R.n_inner, E = R.n_nodes - 1;
    const double inf = std::numeric_limits<double>::infinity();
    for (int e = 0; e < E; ++e) { if (qpic) qpic[e] = inf; if (eqpic) eqpic[e] = inf; }
    if (!R.bifurcating) return;
    double* out[2] = {qpic, eqpic};
    for_pair_rows_parallel(I, E, 2, out, [&](int iu, double* const* mine) {
        double *qpl = mine[0], *eql = mine[1];
        for (int iv = iu + 1; iv < I; ++iv) {
            const uint64_t* s = sums + ((size_t)iu * I + iv) * 3;
            uint64_t p1 = s[0], p2 = s[1], p3 = s[2];
            if (!exact_qp) { p1 &= 0xffffffffull; p2 &= 0xffffffffull; p3 &= 0xffffffffull; }   // `unsigned p1,p2,p3` (:382)
            const double qp = host_log_score(p1, p2, p3);
            const int u = R.inner_node[iu], v = R.inner_node[iv];
            // :475-481 adjacency through the primary links (for the root: its first child link); an edge has one adjacent pair,
            // so the private arrays hold at most one finite QP-IC per edge and the final min-merge is an assignment
            const int u_outer = (u == 0) ? R.first_child[0] : R.parent[u];
            const int v_outer = (v == 0) ? R.first_child[0] : R.parent[v];
            if (u_outer == v) qpl[(u == 0) ? R.parent_edge[v] : R.parent_edge[u]] = qp;
            else if (v_outer == u) qpl[(v == 0) ? R.parent_edge[u] : R.parent_edge[v]] = qp;
            for_path_edges(R, u, v, [&](int e) { if (qp < eql[e]) eql[e] = qp; });
        }
    });
}

}  // namespace

// ===================================================================================================
// C ABI
// ===================================================================================================

extern "C" {

int qs_abi_version(void) { return QS_ABI_VERSION; }

const char* qs_strerror(int code) {
    switch (code) {
        case QS_OK: return "ok";
        case QS_E_ARG: return "bad argument";
        case QS_E_CUDA: return "CUDA failure";
        case QS_E_TREE: return "malformed tree";
        case QS_E_REFERENCE: return "unusable reference tree";
        case QS_E_MEMORY: return "Insufficient memory!";
        case QS_E_UNSUPPORTED: return "unsupported input";
        case QS_E_STATE: return "call order";
        default: return "unknown error";
    }
}

const char* qs_last_error(const qs_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int qs_create(qs_ctx** out, int n_taxa, int cint_bytes, int mode, int device, int shard_index, int shard_count) {
    if (!out) return QS_E_ARG;
    *out = nullptr;
    if (n_taxa < 4 || n_taxa > 32768) return QS_E_ARG;
    if (cint_bytes != 1 && cint_bytes != 2 && cint_bytes != 4 && cint_bytes != 8) return QS_E_ARG;
    if (mode != QS_MODE_TABLE && mode != QS_MODE_TABLE_FREE && mode != QS_MODE_AUTO) return QS_E_ARG;
    if (shard_count < 1 || shard_index < 0 || shard_index >= shard_count) return QS_E_ARG;
    if (device == QS_DEVICE_NONE) {
        // host-only context: no CUDA call at all; only qs_set_reference / qs_score_finalize / qs_shard_range work on it
        qs_ctx* c = new qs_ctx();
        c->host_only = true; c->device = device; c->n = n_taxa; c->n_pad = (n_taxa + 7) / 8 * 8; c->cint_bytes = cint_bytes;
        c->auto_mode = mode == QS_MODE_AUTO; c->mode = c->auto_mode ? QS_MODE_TABLE : mode;
        c->shard_index = shard_index; c->shard_count = shard_count;
        shard_bounds(n_taxa, shard_index, shard_count, 1.0, &c->d_begin, &c->d_end);
        c->rank_begin = binom4((uint64_t)c->d_begin); c->rank_end = binom4((uint64_t)c->d_end);
        *out = c;
        return QS_OK;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { cudaGetLastError(); return QS_E_CUDA; }
    if (cudaSetDevice(device) != cudaSuccess) return QS_E_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return QS_E_CUDA;
    if (prop.major < 10) return QS_E_CUDA;   // sm_100a code only
    qs_ctx* c = new qs_ctx();
    c->device = device; c->n = n_taxa; c->n_pad = (n_taxa + 7) / 8 * 8; c->cint_bytes = cint_bytes;
    c->auto_mode = mode == QS_MODE_AUTO; c->mode = c->auto_mode ? QS_MODE_TABLE : mode;
    c->shard_index = shard_index; c->shard_count = shard_count;
    c->num_sms = prop.multiProcessorCount; c->smem_optin = (int)prop.sharedMemPerBlockOptin;
    shard_bounds(n_taxa, shard_index, shard_count, 1.0, &c->d_begin, &c->d_end);
    c->rank_begin = binom4((uint64_t)c->d_begin); c->rank_end = binom4((uint64_t)c->d_end);
    if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return QS_E_CUDA; }
    c->stream = c->own_stream;
    for (auto& e : c->ev) if (cudaEventCreate(&e) != cudaSuccess) { free_all(c); delete c; return QS_E_CUDA; }
    if (cudaMalloc((void**)&c->d_flags, 2 * sizeof(int)) != cudaSuccess || cudaMalloc((void**)&c->d_nA, sizeof(int32_t)) != cudaSuccess ||
        cudaMalloc((void**)&c->d_counter, sizeof(int)) != cudaSuccess || cudaMallocHost((void**)&c->h_flags, 4 * sizeof(int)) != cudaSuccess) { free_all(c); delete c; return QS_E_CUDA; }
    *out = c;
    return QS_OK;
}

int qs_destroy(qs_ctx* ctx) {
    if (!ctx) return QS_E_ARG;
    if (!ctx->host_only) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        free_all(ctx);
    }
    delete ctx;
    return QS_OK;
}

int qs_set_count_scale(qs_ctx* ctx, int count_scale) {
    if (!ctx || (count_scale != 1 && count_scale != 2)) return QS_E_ARG;
    ctx->fused_scale = count_scale;
    return QS_OK;
}

int qs_set_stream(qs_ctx* ctx, void* cuda_stream) {
    if (!ctx) return QS_E_ARG;
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return QS_OK;
}

int qs_set_reference(qs_ctx* ctx, int n_nodes, const int32_t* parent, const int32_t* parent_edge, const int32_t* leaf_lookup_id,
                     const int32_t* first_child, const int32_t* next_sibling) {
    if (!ctx || !parent || !parent_edge || !leaf_lookup_id || !first_child || !next_sibling) return QS_E_ARG;
    if (!ctx->host_only) QS_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->has_ref = false;
    int r = build_reference(ctx, n_nodes, parent, parent_edge, leaf_lookup_id, first_child, next_sibling);
    if (r) return r;
    if (ctx->host_only) { ctx->has_ref = true; return QS_OK; }
    const size_t n = ctx->n;
    if ((r = dev_alloc(ctx, &ctx->d_lcapd, n * n))) return r;
    if ((r = dev_alloc(ctx, &ctx->d_idepth, (size_t)ctx->ref.n_inner))) return r;
    {
        std::vector<uint32_t> pd(n * n);
        for (size_t i = 0; i < n * n; ++i) { const uint16_t pn = ctx->ref.lca[i]; pd[i] = (uint32_t)pn | ((uint32_t)ctx->ref.idepth[pn] << 16); }
        QS_CUDA(ctx, cudaMemcpy(ctx->d_lcapd, pd.data(), n * n * 4, cudaMemcpyHostToDevice));
    }
    QS_CUDA(ctx, cudaMemcpy(ctx->d_idepth, ctx->ref.idepth.data(), (size_t)ctx->ref.n_inner * 2, cudaMemcpyHostToDevice));
    {   // run-length encoded LCA rows and the parent arrays of the scan / per-edge reduction kernels (kernels/score.cuh)
        const HostRef& R = ctx->ref;
        auto up = [&](auto** dst, const auto& v) -> int {
            int rr = dev_alloc(ctx, dst, std::max<size_t>(1, v.size()));
            if (rr) return rr;
            if (!v.empty() && cudaMemcpy(*dst, v.data(), v.size() * sizeof(v[0]), cudaMemcpyHostToDevice) != cudaSuccess) { ctx->err = "cudaMemcpy of the reference arrays failed"; return QS_E_CUDA; }
            return QS_OK;
        };
        if ((r = up(&ctx->d_inner_parent, R.inner_parent)) || (r = up(&ctx->d_leaf_parent, R.leaf_parent)) || (r = up(&ctx->d_inner_gap, R.inner_gap)) || (r = up(&ctx->d_inner_node, R.inner_node)) ||
            (r = up(&ctx->d_node_parent, R.parent)) || (r = up(&ctx->d_node_depth, R.depth)) || (r = up(&ctx->d_node_edge, R.parent_edge)) ||
            (r = up(&ctx->d_node_inner, R.inner_index)))
            return r;
    }
    for (auto** q : {(void**)&ctx->d_pair_sums, (void**)&ctx->d_pair_best, (void**)&ctx->d_pair_score, (void**)&ctx->d_pair_score_local, (void**)&ctx->d_edge, (void**)&ctx->d_edge_out})
        if (*q) { cudaFree(*q); *q = nullptr; }
    ctx->fused_partials_valid = false; ctx->partials_ready = false;
    ctx->scan_dB = ctx->scan_dE = -1;
    ctx->has_ref = true;
    return QS_OK;
}

int qs_clear_trees(qs_ctx* ctx) {
    if (!ctx) return QS_E_ARG;
    ctx->m = 0; ctx->total_nodes = 0; ctx->max_nodes = 0;
    ctx->h_off.assign(1, 0);
    ctx->counted = false; ctx->dist_valid = false; ctx->fused_partials_valid = false;
    return QS_OK;
}

int qs_num_trees(const qs_ctx* ctx, int64_t* n_trees) {
    if (!ctx || !n_trees) return QS_E_ARG;
    *n_trees = ctx->m;
    return QS_OK;
}

int qs_add_trees(qs_ctx* ctx, int n_trees, const int64_t* node_offsets, const int32_t* parent, const int32_t* leaf_lookup_id) {
    if (!ctx || n_trees < 0 || !node_offsets || !parent || !leaf_lookup_id) return QS_E_ARG;
    if (ctx->host_only) QS_FAIL(ctx, QS_E_STATE, "host-only context (QS_DEVICE_NONE): no device work; there is no CPU fallback");
    if (n_trees == 0) return QS_OK;
    QS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (node_offsets[0] != 0) QS_FAIL(ctx, QS_E_TREE, "node_offsets[0] must be 0");
    int mx = ctx->max_nodes;
    for (int t = 0; t < n_trees; ++t) {
        int64_t cnt = node_offsets[t + 1] - node_offsets[t];
        if (cnt < 1 || cnt > 1000000) QS_FAIL(ctx, QS_E_TREE, "tree %d has %lld nodes", t, (long long)cnt);
        if (parent[node_offsets[t]] != -1) QS_FAIL(ctx, QS_E_TREE, "tree %d: first node must be the root (parent -1)", t);
        mx = std::max<int>(mx, (int)cnt);
    }
    const int64_t add_nodes = node_offsets[n_trees];
    const int64_t new_m = ctx->m + n_trees, new_nodes = ctx->total_nodes + add_nodes;
    if (new_m > 0x7fffffffLL) QS_FAIL(ctx, QS_E_UNSUPPORTED, "too many trees");
    // grow device arrays (amortised doubling), preserving earlier trees
    if (new_nodes > ctx->cap_nodes) {
        int64_t cap = std::max<int64_t>(new_nodes, ctx->cap_nodes * 2);
        int32_t *np = nullptr, *nl = nullptr;
        if (cudaMalloc((void**)&np, cap * 4) != cudaSuccess || cudaMalloc((void**)&nl, cap * 4) != cudaSuccess) {
            cudaGetLastError(); cudaFree(np);
            QS_FAIL(ctx, QS_E_MEMORY, "Insufficient memory! (tree arrays)");
        }
        if (ctx->total_nodes) {
            QS_CUDA(ctx, cudaMemcpyAsync(np, ctx->d_parent, ctx->total_nodes * 4, cudaMemcpyDeviceToDevice, ctx->stream));
            QS_CUDA(ctx, cudaMemcpyAsync(nl, ctx->d_leaf, ctx->total_nodes * 4, cudaMemcpyDeviceToDevice, ctx->stream));
            QS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        }
        cudaFree(ctx->d_parent); cudaFree(ctx->d_leaf);
        ctx->d_parent = np; ctx->d_leaf = nl; ctx->cap_nodes = cap;
    }
    if (new_m + 1 > ctx->cap_trees) {
        int64_t cap = std::max<int64_t>(new_m + 1, ctx->cap_trees * 2);
        int64_t* no = nullptr;
        if (cudaMalloc((void**)&no, cap * 8) != cudaSuccess) { cudaGetLastError(); QS_FAIL(ctx, QS_E_MEMORY, "Insufficient memory! (tree offsets)"); }
        cudaFree(ctx->d_off);
        ctx->d_off = no; ctx->cap_trees = cap;
        if (ctx->m) QS_CUDA(ctx, cudaMemcpyAsync(ctx->d_off, ctx->h_off.data(), (ctx->m + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    }
    const int64_t base = ctx->total_nodes;
    ctx->h_off.reserve(new_m + 1);
    for (int t = 1; t <= n_trees; ++t) ctx->h_off.push_back(base + node_offsets[t]);
    QS_CUDA(ctx, cudaMemcpyAsync(ctx->d_parent + base, parent, add_nodes * 4, cudaMemcpyHostToDevice, ctx->stream));
    QS_CUDA(ctx, cudaMemcpyAsync(ctx->d_leaf + base, leaf_lookup_id, add_nodes * 4, cudaMemcpyHostToDevice, ctx->stream));
    QS_CUDA(ctx, cudaMemcpyAsync(ctx->d_off + ctx->m, ctx->h_off.data() + ctx->m, (n_trees + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    QS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // caller may reuse its buffers
    ctx->m = new_m; ctx->total_nodes = new_nodes; ctx->max_nodes = mx;
    ctx->counted = false; ctx->dist_valid = false; ctx->fused_partials_valid = false;
    return QS_OK;
}

int qs_count(qs_ctx* ctx) {
    if (!ctx) return QS_E_ARG;
    if (ctx->host_only) QS_FAIL(ctx, QS_E_STATE, "host-only context (QS_DEVICE_NONE): no device work; there is no CPU fallback");
    QS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ctx->m == 0) QS_FAIL(ctx, QS_E_STATE, "no evaluation trees were added");
    if ((uint64_t)ctx->m > cint_mask(ctx->cint_bytes)) QS_FAIL(ctx, QS_E_ARG, "%lld trees do not fit a %d-byte counter", (long long)ctx->m, ctx->cint_bytes);
    ctx->counted = false; ctx->fused_partials_valid = false;
    int r;
    QS_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    if ((r = run_distances(ctx))) return r;
    QS_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    // error flags and the class split come back asynchronously (pinned buffer); they are looked at when the stream is drained
    QS_CUDA(ctx, cudaMemcpyAsync(ctx->h_flags, ctx->d_flags, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    QS_CUDA(ctx, cudaMemcpyAsync(ctx->h_flags + 2, ctx->d_nA, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    auto check_flags = [&]() -> int {
        ctx->n_class_a = ctx->h_flags[2]; ctx->counted_once = true;
        if (ctx->h_flags[1] != 0) QS_FAIL(ctx, QS_E_TREE, "malformed evaluation tree (code %d): need parent[i] < i, leaf ids in [0,n) exactly on leaves, each taxon at most once per tree", ctx->h_flags[1]);
        if (ctx->h_flags[0] > kMaxHalfExact) QS_FAIL(ctx, QS_E_UNSUPPORTED, "an evaluation tree has a leaf-to-leaf path of %d edges; this build packs distances in fp16 (exact up to %d)", ctx->h_flags[0], kMaxHalfExact);
        return QS_OK;
    };
    const uint64_t nq = ctx->rank_end - ctx->rank_begin;
    const size_t need = (size_t)nq * 3 * ctx->cint_bytes;
    // QS_MODE_AUTO: keep the shard's table resident if it fits beside the matrices.  Decided once per (table size, tree count): the
    // query is an ioctl into the kernel driver, which sleeps on the driver's global lock — on a host whose other GPUs are busy that
    // took up to 100 ms of a 4.8 ms step (profiles/r02_x_jitter.txt, tools/ctx_switches.py)
    if (ctx->auto_mode && !(ctx->auto_need == need && ctx->auto_m == ctx->m)) {
        ctx->auto_need = need; ctx->auto_m = ctx->m;
        size_t free_b = 0, total_b = 0;
        QS_CUDA(ctx, cudaMemGetInfo(&free_b, &total_b));
        size_t limit = free_b + ctx->table_bytes > ((size_t)1 << 30) ? free_b + ctx->table_bytes - ((size_t)1 << 30) : 0;
        if (const char* env = getenv("QS_TABLE_BYTES_LIMIT")) limit = (size_t)strtoull(env, nullptr, 10);      // test hook
        const int want = need + kTableSlack <= limit ? QS_MODE_TABLE : QS_MODE_TABLE_FREE;
        if (want != ctx->mode && ctx->d_table) { cudaFree(ctx->d_table); ctx->d_table = nullptr; ctx->table_bytes = 0; }
        ctx->mode = want;
    }
    if (ctx->mode == QS_MODE_TABLE_FREE) {
        // (the class split decides whether the slabs' task tables carry role Y, and a malformed tree should fail before minutes of
        // counting: one wait for the distance kernels, < 1 % of such a step)
        QS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if ((r = check_flags())) return r;
    }
    QS_CUDA(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
    if (ctx->mode == QS_MODE_TABLE) {
        if (need > ctx->table_bytes) {
            if (ctx->d_table) { cudaFree(ctx->d_table); ctx->d_table = nullptr; ctx->table_bytes = 0; }
            cudaError_t e = cudaMalloc(&ctx->d_table, need + kTableSlack);
            if (e != cudaSuccess) { cudaGetLastError(); QS_FAIL(ctx, QS_E_MEMORY, "Insufficient memory! count table of %zu bytes does not fit this device (use QS_MODE_AUTO / QS_MODE_TABLE_FREE or more shards)", need); }
            ctx->table_bytes = need;
        }
        if ((r = run_count_rows(ctx, ctx->d_begin, ctx->d_end, ctx->d_table))) return r;
    } else {
        if ((r = run_table_free(ctx))) return r;
    }
    QS_CUDA(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
    QS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]); ctx->dist_ms = ms;
    cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]); ctx->count_ms = ms;
    if ((r = check_flags())) return r;
    ctx->counted = true;
    return QS_OK;
}

int qs_rebalance_shards(qs_ctx* ctx, int* changed) {
    if (!ctx) return QS_E_ARG;
    if (changed) *changed = 0;
    if (ctx->host_only) QS_FAIL(ctx, QS_E_STATE, "host-only context (QS_DEVICE_NONE): no device work; there is no CPU fallback");
    if (ctx->m == 0) QS_FAIL(ctx, QS_E_STATE, "no evaluation trees were added");
    if (ctx->shard_count == 1) return QS_OK;
    QS_CUDA(ctx, cudaSetDevice(ctx->device));
    int r;
    if (!ctx->dist_valid) {                                    // classify the trees (the distance kernels do that; they are < 1 % of a step)
        if ((r = run_distances(ctx))) return r;
        QS_CUDA(ctx, cudaMemcpyAsync(ctx->h_flags + 2, ctx->d_nA, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        QS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->n_class_a = ctx->h_flags[2]; ctx->counted_once = true;
    }
    int b = 0, e = 0;
    shard_bounds(ctx->n, ctx->shard_index, ctx->shard_count, (double)(ctx->m - ctx->n_class_a) / (double)ctx->m, &b, &e);
    if (b == ctx->d_begin && e == ctx->d_end) return QS_OK;
    QS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->d_table) { cudaFree(ctx->d_table); ctx->d_table = nullptr; ctx->table_bytes = 0; }
    ctx->d_begin = b; ctx->d_end = e;
    ctx->rank_begin = binom4((uint64_t)b); ctx->rank_end = binom4((uint64_t)e);
    ctx->plan_dB = ctx->plan_dE = -1; ctx->scan_dB = ctx->scan_dE = -1;
    ctx->counted = false; ctx->fused_partials_valid = false; ctx->partials_ready = false;
    if (changed) *changed = 1;
    return QS_OK;
}

int qs_table_resident(const qs_ctx* ctx, int* resident) {
    if (!ctx || !resident) return QS_E_ARG;
    *resident = (ctx->mode == QS_MODE_TABLE) ? 1 : 0;
    return QS_OK;
}

int qs_score_num_pairs(const qs_ctx* ctx, int64_t* n_pairs) {
    if (!ctx || !n_pairs) return QS_E_ARG;
    *n_pairs = (int64_t)ctx->ref.n_inner * ctx->ref.n_inner;
    return QS_OK;
}

int qs_score_inner_nodes(const qs_ctx* ctx, int32_t* node_of_inner, int64_t capacity, int64_t* n_inner) {
    if (!ctx || !n_inner) return QS_E_ARG;
    if (!ctx->has_ref) return QS_E_STATE;
    *n_inner = ctx->ref.n_inner;
    if (node_of_inner) {
        if (capacity < ctx->ref.n_inner) return QS_E_ARG;
        memcpy(node_of_inner, ctx->ref.inner_node.data(), (size_t)ctx->ref.n_inner * sizeof(int32_t));
    }
    return QS_OK;
}

int qs_score_partials(qs_ctx* ctx, int count_scale, double* lqic_partial, uint64_t* pair_sums) {
    if (!ctx || !lqic_partial || !pair_sums) return QS_E_ARG;
    if (ctx->host_only) QS_FAIL(ctx, QS_E_STATE, "host-only context (QS_DEVICE_NONE): no device work; there is no CPU fallback");
    QS_CUDA(ctx, cudaSetDevice(ctx->device));
    int r = run_score_scan(ctx, count_scale);
    if (r) return r;
    const size_t I = ctx->ref.n_inner;
    QS_CUDA(ctx, cudaMemcpyAsync(pair_sums, ctx->d_pair_sums, I * I * 3 * 8, cudaMemcpyDeviceToHost, ctx->stream));
    return run_edge_reduce(ctx, 0, lqic_partial, nullptr, nullptr);      // (drains the stream: the copy above has landed)
}

int qs_score_scan(qs_ctx* ctx, int count_scale) {
    if (!ctx) return QS_E_ARG;
    if (ctx->host_only) QS_FAIL(ctx, QS_E_STATE, "host-only context (QS_DEVICE_NONE): no device work; there is no CPU fallback");
    QS_CUDA(ctx, cudaSetDevice(ctx->device));
    int r = run_score_scan(ctx, count_scale);
    if (r) return r;
    // this shard's own scores, kept for qs_score_select_winners after the caller's MIN all-reduce of pair_score
    const size_t I = ctx->ref.n_inner;
    if (!ctx->d_pair_score_local && (r = dev_alloc(ctx, &ctx->d_pair_score_local, I * I))) return r;
    QS_CUDA(ctx, cudaMemcpyAsync(ctx->d_pair_score_local, ctx->d_pair_score, I * I * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    return QS_OK;
}

int qs_score_device_partials(qs_ctx* ctx, void** pair_sums, void** pair_score, void** pair_best, int64_t* n_pairs) {
    if (!ctx || !pair_sums || !pair_score || !pair_best || !n_pairs) return QS_E_ARG;
    if (ctx->host_only) QS_FAIL(ctx, QS_E_STATE, "host-only context (QS_DEVICE_NONE): no device work; there is no CPU fallback");
    if (!ctx->partials_ready) QS_FAIL(ctx, QS_E_STATE, "no scanned partials: call qs_score_scan first");
    *pair_sums = ctx->d_pair_sums; *pair_score = ctx->d_pair_score; *pair_best = ctx->d_pair_best;
    *n_pairs = (int64_t)ctx->ref.n_inner * ctx->ref.n_inner;
    return QS_OK;
}

int qs_score_select_winners(qs_ctx* ctx) {
    if (!ctx) return QS_E_ARG;
    if (ctx->host_only) QS_FAIL(ctx, QS_E_STATE, "host-only context (QS_DEVICE_NONE): no device work; there is no CPU fallback");
    if (!ctx->partials_ready || !ctx->d_pair_score_local) QS_FAIL(ctx, QS_E_STATE, "call qs_score_scan first");
    QS_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t I = ctx->ref.n_inner;
    qs_select_winners_kernel<<<(unsigned)std::max<size_t>(1, std::min<size_t>((I * I + 255) / 256, 4096)), 256, 0, ctx->stream>>>(ctx->d_pair_score_local, ctx->d_pair_score,
                                                                                                                                ctx->d_pair_best, I * I);
    ctx->launches++;
    QS_CUDA(ctx, cudaGetLastError());
    return QS_OK;
}

int qs_score_finish(qs_ctx* ctx, int exact_qp, double* lqic, double* qpic, double* eqpic) {
    if (!ctx || !lqic) return QS_E_ARG;
    if (ctx->host_only) QS_FAIL(ctx, QS_E_STATE, "host-only context (QS_DEVICE_NONE): no device work; there is no CPU fallback");
    QS_CUDA(ctx, cudaSetDevice(ctx->device));
    return run_edge_reduce(ctx, exact_qp, lqic, qpic, eqpic);
}

int qs_score_finalize(qs_ctx* ctx, int exact_qp, const double* lqic_reduced, const uint64_t* pair_sums_reduced,
                      double* lqic, double* qpic, double* eqpic) {
    if (!ctx || !lqic_reduced || !pair_sums_reduced || !lqic) return QS_E_ARG;
    if (!ctx->has_ref) QS_FAIL(ctx, QS_E_STATE, "qs_set_reference has not been called");
    const int E = ctx->ref.n_nodes - 1;
    if (lqic != lqic_reduced) memcpy(lqic, lqic_reduced, (size_t)E * 8);
    qp_from_pairs(ctx, pair_sums_reduced, exact_qp, qpic, eqpic);
    return QS_OK;
}

int qs_score(qs_ctx* ctx, int count_scale, int exact_qp, double* lqic, double* qpic, double* eqpic) {
    if (!ctx || !lqic) return QS_E_ARG;
    if (ctx->host_only) QS_FAIL(ctx, QS_E_STATE, "host-only context (QS_DEVICE_NONE): no device work; there is no CPU fallback");
    QS_CUDA(ctx, cudaSetDevice(ctx->device));
    int r = run_score_scan(ctx, count_scale);
    if (r) return r;
    return run_edge_reduce(ctx, exact_qp, lqic, qpic, eqpic);
}

int qs_shard_bounds(int n_taxa, int shard_index, int shard_count, int* s3_begin, int* s3_end, uint64_t* rank_begin, uint64_t* rank_end) {
    if (n_taxa < 4 || shard_count < 1 || shard_index < 0 || shard_index >= shard_count) return QS_E_ARG;
    int b = 0, e = 0;
    shard_bounds(n_taxa, shard_index, shard_count, 1.0, &b, &e);
    if (s3_begin) *s3_begin = b;
    if (s3_end) *s3_end = e;
    if (rank_begin) *rank_begin = binom4((uint64_t)b);
    if (rank_end) *rank_end = binom4((uint64_t)e);
    return QS_OK;
}

int qs_plan_stats(int n_taxa, int s3_begin, int s3_end, int64_t* stats) {
    if (n_taxa < 4 || n_taxa > 32768 || !stats || s3_begin < 0 || s3_end > n_taxa || s3_begin > s3_end) return QS_E_ARG;
    const int dB = std::max(3, s3_begin), dE = s3_end;
    const size_t row_bytes = (size_t)((n_taxa + 7) / 8 * 8) * 2;
    const int max_rows = (int)std::max<size_t>(4, std::min<size_t>((size_t)n_taxa, (24 * 1024) / row_bytes));
    HostEnum H;
    build_enum_tables(n_taxa, dB, dE, H);
    std::vector<RowTask> xt, yt;
    const int threads = cr_threads_for(n_taxa);
    build_row_tasks(H, n_taxa, dB, dE, max_rows, threads, xt, yt);
    int64_t items[ITEM_KINDS] = {}, slots[ITEM_KINDS] = {}, rows = 0, mx = 0, violations = 0, quartets = 0, slot0_covered = 0, slot12_covered = 0;
    const int cap[ITEM_KINDS] = {threads, 2 * threads, 2 * threads, threads, 2 * threads};
    int64_t next_e[ITEM_KINDS] = {};
    auto in_ranges = [](const RowTask& t, int row) {
        for (int k = 0; k < 3; ++k) if (row >= t.rstart[k] && row < t.rstart[k] + t.rcount[k]) return true;
        return false;
    };
    for (auto* v : {&xt, &yt})
        for (auto& t : *v) {
            if (t.e0 != next_e[t.kind] || t.ne < 1 || t.ne > cap[t.kind]) ++violations;       // tasks tile the enumeration
            next_e[t.kind] = t.e0 + t.ne;
            items[t.kind] += t.ne; slots[t.kind] += cap[t.kind];
            const int r = t.rcount[0] + t.rcount[1] + t.rcount[2];
            rows += r; mx = std::max<int64_t>(mx, r);
            if (r > max_rows && t.ne > 1) ++violations;
            for (int i = 0; i < t.ne; ++i) {                                                   // every item's rows are staged, ids are sane
                int p, q, j, k2;
                if (t.kind == ITEM_Y) {
                    cr_decode_y(H.view(), t.e0 + i, n_taxa, p, q, j, k2);
                    if (!(p >= 1 && p < q && q <= dE - 2 && j < (p + 7) / 8 && k2 * 8 >= H.z_first && k2 * 8 + 8 <= H.z_last && k2 * 8 + 7 > q)) ++violations;   // whole d-blocks inside Y's range, some d above c
                    // the quartets the kernel's flush keeps of this item: a < b in the a-block, d > c in the d-block (count_rows.cuh)
                    slot0_covered += (int64_t)std::max(0, std::min(8, p - j * 8)) * std::max(0, k2 * 8 + 8 - std::max(k2 * 8, q + 1));
                } else if (t.kind == ITEM_Z) {
                    cr_decode_z(H.view(), t.e0 + i, n_taxa, dB, dE, p, q, j, k2);                  // p = d, q = a, blocks j <= k2
                    if (!(p >= dB && p < dE && (p < H.z_first || p >= H.z_last) && q >= 0 && p - q >= 3 && j >= ((q + 1) >> 3) && j <= k2 && k2 <= ((p - 1) >> 3))) ++violations;
                    for (int cc = k2 * 8; cc < k2 * 8 + 8 && cc < p; ++cc)                         // the flush keeps a < b < c < d
                        slot0_covered += std::max(0, std::min(j * 8 + 8, cc) - std::max(j * 8, q + 1));
                } else {
                    cr_decode_x(H.prefix(t.kind), t.kind, H.xo_diag, t.e0 + i, n_taxa, dB, p, q, j);
                    if (!(p >= 2 && p < q && q >= dB && q < dE && j < (t.kind == ITEM_XO ? cr_nxo(p, H.xo_diag) : t.kind == ITEM_XR ? cr_nxr(p, H.xo_diag) : cr_nxd(p, H.xo_diag)))) ++violations;
                    // the pairs a < b < c the kernel's flush keeps of this item (count_rows.cuh): p = c here
                    const int nf = cr_nfull(p), rg = p - nf * 8;
                    auto pairs_in = [](int k) { return k * (k - 1) / 2; };
                    if (t.kind == ITEM_XO) slot12_covered += j < nf * (nf - 1) / 2 ? 64 : pairs_in(8);              // a-block below a full b-block / a full diagonal block
                    else if (t.kind == ITEM_XR) slot12_covered += j < nf ? 8 * rg : pairs_in(rg);                   // ... the ragged b-block / the ragged diagonal block
                    else slot12_covered += pairs_in(std::min(8, p - j * 8));                                         // XD: diagonal block j
                }
                if (!in_ranges(t, p) || !in_ranges(t, q)) ++violations;
            }
        }
    for (int k = 0; k < ITEM_KINDS; ++k) if (next_e[k] != H.total(k)) ++violations;
    if (slot0_covered != (int64_t)(binom4((uint64_t)dE) - binom4((uint64_t)dB))) ++violations;      // roles Y and Z together hold every quartet of the range exactly once
    if (slot12_covered != (int64_t)(binom4((uint64_t)dE) - binom4((uint64_t)dB))) ++violations;     // ... and so do the role-X kinds
    if (dB < dE && !(dB <= H.z_first && H.z_first <= H.z_last && H.z_last <= dE && (H.z_first == dE || H.z_first % 8 == 0) && (H.z_last == dE || H.z_last % 8 == 0))) ++violations;
    quartets = (int64_t)(binom4((uint64_t)dE) - binom4((uint64_t)dB));
    stats[0] = (int64_t)xt.size(); stats[1] = (int64_t)yt.size();
    stats[2] = items[0]; stats[3] = items[1]; stats[4] = items[ITEM_Y] + items[ITEM_Z];
    stats[5] = slots[0]; stats[6] = slots[1]; stats[7] = slots[ITEM_Y] + slots[ITEM_Z];
    stats[8] = rows; stats[9] = mx; stats[10] = violations; stats[11] = quartets;
    stats[12] = items[ITEM_XR]; stats[13] = slots[ITEM_XR];
    // useful compares / issued compares of role X: an XO item issues 128 (8 x 8 quartets x 2 slots), an XD item 64, an XR item 16 per valid b
    int64_t xr_compares = 0;
    for (int cc = 2; cc < n_taxa; ++cc) xr_compares += (int64_t)std::max(0, dE - cr_dlo(cc, dB)) * cr_nxr(cc, H.xo_diag) * 16 * (cc & 7);
    stats[14] = items[ITEM_XO] * 128 + items[ITEM_XD] * 64 + xr_compares;
    stats[15] = items[ITEM_Z];
    stats[16] = (items[ITEM_Y] + items[ITEM_Z]) * 64;
    return QS_OK;
}

int qs_shard_range(const qs_ctx* ctx, uint64_t* rank_begin, uint64_t* rank_end) {
    if (!ctx || !rank_begin || !rank_end) return QS_E_ARG;
    *rank_begin = ctx->rank_begin; *rank_end = ctx->rank_end;
    return QS_OK;
}

int qs_get_counts(qs_ctx* ctx, uint64_t rank_begin, uint64_t rank_end, void* out) {
    if (!ctx || !out || rank_end < rank_begin) return QS_E_ARG;
    if (ctx->host_only) QS_FAIL(ctx, QS_E_STATE, "host-only context (QS_DEVICE_NONE): no device work; there is no CPU fallback");
    if (!ctx->counted) QS_FAIL(ctx, QS_E_STATE, "qs_count has not been called");
    QS_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t eb = 3 * (size_t)ctx->cint_bytes;
    memset(out, 0, (rank_end - rank_begin) * eb);
    const uint64_t lo = std::max(rank_begin, ctx->rank_begin), hi = std::min(rank_end, ctx->rank_end);
    if (lo < hi) return counts_to_host(ctx, lo, hi, (char*)out + (lo - rank_begin) * eb);
    return QS_OK;
}

// ---- host ingest: Newick text -> qs_add_trees (parser in ingest.cpp) -------------------------------------------
int qs_add_newick(qs_ctx* ctx, const char* text, size_t text_len, const char* const* taxon_names, int n_threads, int64_t* n_trees_added) {
    if (!ctx || (!text && text_len) || !taxon_names) return QS_E_ARG;
    if (ctx->host_only) QS_FAIL(ctx, QS_E_STATE, "host-only context (QS_DEVICE_NONE): no device work; there is no CPU fallback");
    qs_flat_trees* f = nullptr;
    char eb[400];
    eb[0] = 0;
    int r = qs_newick_flatten(text, text_len, ctx->n, taxon_names, n_threads, &f, eb, sizeof eb);
    if (r) QS_FAIL(ctx, r, "%s", eb);
    int64_t T = 0, N = 0;
    const int64_t* off; const int32_t *par, *leaf;
    qs_flat_trees_view(f, &T, &N, &off, &par, &leaf);
    if (T > 0x7fffffffLL) { qs_flat_trees_free(f); QS_FAIL(ctx, QS_E_UNSUPPORTED, "too many trees"); }
    r = qs_add_trees(ctx, (int)T, off, par, leaf);
    qs_flat_trees_free(f);
    if (r == QS_OK && n_trees_added) *n_trees_added = T;
    return r;
}

int qs_add_newick_file(qs_ctx* ctx, const char* path, const char* const* taxon_names, int n_threads, int64_t* n_trees_added) {
    if (!ctx || !path || !taxon_names) return QS_E_ARG;
    FILE* fp = fopen(path, "rb");
    if (!fp) QS_FAIL(ctx, QS_E_ARG, "cannot open %s", path);
    std::string buf;
    char chunk[1 << 16];
    size_t got;
    while ((got = fread(chunk, 1, sizeof chunk, fp)) > 0) buf.append(chunk, got);
    fclose(fp);
    return qs_add_newick(ctx, buf.data(), buf.size(), taxon_names, n_threads, n_trees_added);
}

// ---- table persistence ------------------------------------------------------------------------------------------
namespace {
struct TableHeader {
    char magic[8];
    int32_t n_taxa, cint_bytes;
    int64_t n_trees;
    int32_t s3_begin, s3_end;
    uint64_t rank_begin, rank_end;
    uint64_t reserved[2];
};
static_assert(sizeof(TableHeader) == 64, "table file header is 64 bytes");
constexpr size_t kIoChunk = (size_t)64 << 20;
}  // namespace

int qs_save_table(qs_ctx* ctx, const char* path) {
    if (!ctx || !path) return QS_E_ARG;
    if (ctx->host_only) QS_FAIL(ctx, QS_E_STATE, "host-only context (QS_DEVICE_NONE): no device work; there is no CPU fallback");
    if (ctx->mode != QS_MODE_TABLE) QS_FAIL(ctx, QS_E_STATE, "table-free context keeps no table");
    if (!ctx->counted) QS_FAIL(ctx, QS_E_STATE, "qs_count has not been called");
    QS_CUDA(ctx, cudaSetDevice(ctx->device));
    FILE* fp = fopen(path, "wb");
    if (!fp) QS_FAIL(ctx, QS_E_ARG, "cannot create %s", path);
    TableHeader h{};
    memcpy(h.magic, "QSTBL001", 8);
    h.n_taxa = ctx->n; h.cint_bytes = ctx->cint_bytes; h.n_trees = ctx->m; h.s3_begin = ctx->d_begin; h.s3_end = ctx->d_end;
    h.rank_begin = ctx->rank_begin; h.rank_end = ctx->rank_end;
    bool ok = fwrite(&h, sizeof h, 1, fp) == 1;
    const size_t total = (size_t)(ctx->rank_end - ctx->rank_begin) * 3 * ctx->cint_bytes;
    std::vector<char> buf(std::min(total, kIoChunk));
    for (size_t o = 0; ok && o < total; o += kIoChunk) {
        const size_t len = std::min(kIoChunk, total - o);
        cudaError_t e = cudaMemcpyAsync(buf.data(), (const char*)ctx->d_table + o, len, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { fclose(fp); QS_CUDA(ctx, e); }
        ok = fwrite(buf.data(), 1, len, fp) == len;
    }
    ok = (fclose(fp) == 0) && ok;
    if (!ok) QS_FAIL(ctx, QS_E_ARG, "short write to %s", path);
    return QS_OK;
}

int qs_load_table(qs_ctx* ctx, const char* path) {
    if (!ctx || !path) return QS_E_ARG;
    if (ctx->host_only) QS_FAIL(ctx, QS_E_STATE, "host-only context (QS_DEVICE_NONE): no device work; there is no CPU fallback");
    if (ctx->mode != QS_MODE_TABLE) QS_FAIL(ctx, QS_E_STATE, "table-free context keeps no table");
    QS_CUDA(ctx, cudaSetDevice(ctx->device));
    FILE* fp = fopen(path, "rb");
    if (!fp) QS_FAIL(ctx, QS_E_ARG, "cannot open %s", path);
    TableHeader h{};
    if (fread(&h, sizeof h, 1, fp) != 1 || memcmp(h.magic, "QSTBL001", 8) != 0) { fclose(fp); QS_FAIL(ctx, QS_E_ARG, "%s is not a qscuda table file", path); }
    if (h.n_taxa != ctx->n || h.cint_bytes != ctx->cint_bytes || h.s3_begin != ctx->d_begin || h.s3_end != ctx->d_end || h.rank_begin != ctx->rank_begin ||
        h.rank_end != ctx->rank_end) {
        fclose(fp);
        QS_FAIL(ctx, QS_E_ARG, "%s holds %d taxa, %d-byte counters, s3 in [%d,%d); this context has %d taxa, %d-byte counters, s3 in [%d,%d)", path, h.n_taxa,
                h.cint_bytes, h.s3_begin, h.s3_end, ctx->n, ctx->cint_bytes, ctx->d_begin, ctx->d_end);
    }
    const size_t total = (size_t)(ctx->rank_end - ctx->rank_begin) * 3 * ctx->cint_bytes;
    if (total > ctx->table_bytes) {
        if (ctx->d_table) { cudaFree(ctx->d_table); ctx->d_table = nullptr; ctx->table_bytes = 0; }
        cudaError_t e = cudaMalloc(&ctx->d_table, total + kTableSlack);
        if (e != cudaSuccess) { cudaGetLastError(); fclose(fp); QS_FAIL(ctx, QS_E_MEMORY, "Insufficient memory! count table of %zu bytes does not fit this device", total); }
        ctx->table_bytes = total;
    }
    std::vector<char> buf(std::min(total, kIoChunk));
    for (size_t o = 0; o < total; o += kIoChunk) {
        const size_t len = std::min(kIoChunk, total - o);
        if (fread(buf.data(), 1, len, fp) != len) { fclose(fp); QS_FAIL(ctx, QS_E_ARG, "%s is truncated", path); }
        cudaError_t e = cudaMemcpyAsync((char*)ctx->d_table + o, buf.data(), len, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { fclose(fp); QS_CUDA(ctx, e); }
    }
    fclose(fp);
    ctx->counted = true;          // scoring / qs_get_counts may follow without qs_count
    return QS_OK;
}

int qs_get_distances(qs_ctx* ctx, int64_t tree, uint16_t* out) {
    if (!ctx || !out) return QS_E_ARG;
    if (ctx->host_only) QS_FAIL(ctx, QS_E_STATE, "host-only context (QS_DEVICE_NONE): no device work; there is no CPU fallback");
    if (!ctx->dist_valid) QS_FAIL(ctx, QS_E_STATE, "distance matrices are built by qs_count");
    if (tree < 0 || tree >= ctx->m) return QS_E_ARG;
    QS_CUDA(ctx, cudaSetDevice(ctx->device));
    uint16_t* tmp = nullptr;
    const int n = ctx->n;
    QS_CUDA(ctx, cudaMalloc((void**)&tmp, (size_t)n * n * 2));
    qs_dist_to_u16_kernel<<<(n * n + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_D + (size_t)tree * n * ctx->n_pad, n, ctx->n_pad, tmp);
    ctx->launches++;
    cudaError_t e = cudaMemcpyAsync(out, tmp, (size_t)n * n * 2, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(tmp);
    QS_CUDA(ctx, e);
    return QS_OK;
}

int qs_write_raw_qic(qs_ctx* ctx, int count_scale, const char* const* taxon_names, const char* path) {
    if (!ctx) return QS_E_ARG;
    if (ctx->shard_count != 1) QS_FAIL(ctx, QS_E_STATE, "this context holds shard %d of %d: pass all shards to qs_write_raw_qic_shards", ctx->shard_index, ctx->shard_count);
    qs_ctx* one[1] = {ctx};
    return qs_write_raw_qic_shards(one, 1, count_scale, taxon_names, path);
}

int qs_write_raw_qic_shards(qs_ctx* const* ctxs, int n_ctxs, int count_scale, const char* const* taxon_names, const char* path) {
    if (!ctxs || n_ctxs < 1 || !ctxs[0] || !taxon_names || !path) return QS_E_ARG;
    qs_ctx* ctx = ctxs[0];
    if (count_scale != 1 && count_scale != 2) return QS_E_ARG;
    // the shards must tile the whole rank space in order
    uint64_t next = 0;
    for (int g = 0; g < n_ctxs; ++g) {
        qs_ctx* c = ctxs[g];
        if (!c) return QS_E_ARG;
        if (c->host_only) QS_FAIL(ctx, QS_E_STATE, "host-only context (QS_DEVICE_NONE): no device work; there is no CPU fallback");
        if (c->n != ctx->n || c->cint_bytes != ctx->cint_bytes || c->rank_begin != next) QS_FAIL(ctx, QS_E_ARG, "context %d does not continue the rank space at %llu", g, (unsigned long long)next);
        if (!c->counted) QS_FAIL(ctx, QS_E_STATE, "needs qs_count on every shard");
        next = c->rank_end;
    }
    if (next != binom4((uint64_t)ctx->n)) QS_FAIL(ctx, QS_E_ARG, "the contexts do not cover all C(n,4) quartets");
    if (!ctx->has_ref) QS_FAIL(ctx, QS_E_STATE, "needs qs_set_reference");
    // the file is ordered by (a,b,c,d) with a outermost, the table by rank with d outermost: the whole table is gathered on
    // the host first (a table-free shard counts its slabs again for this, see counts_to_host)
    const uint64_t nq = next;
    const size_t eb = (size_t)ctx->cint_bytes;
    std::vector<unsigned char> tab;
    try { tab.resize((size_t)nq * 3 * eb); } catch (const std::exception&) {
        QS_FAIL(ctx, QS_E_MEMORY, "Insufficient memory! the raw QIC dump gathers the whole table (%llu bytes) in host memory", (unsigned long long)(nq * 3 * eb));
    }
    for (int g = 0; g < n_ctxs; ++g) {
        qs_ctx* c = ctxs[g];
        QS_CUDA(ctx, cudaSetDevice(c->device));
        int r = counts_to_host(c, c->rank_begin, c->rank_end, tab.data() + (size_t)c->rank_begin * 3 * eb);
        if (r) { if (c != ctx) ctx->err = c->err; return r; }
    }
    FILE* f = fopen(path, "w");
    if (!f) QS_FAIL(ctx, QS_E_ARG, "cannot open %s for writing", path);
    const HostRef& R = ctx->ref;
    const int n = ctx->n;
    const uint64_t mask = cint_mask(ctx->cint_bytes);
    auto get = [&](uint64_t idx) -> uint64_t {
        uint64_t v = 0;
        memcpy(&v, tab.data() + idx * eb, eb);
        return (v * (uint64_t)count_scale) & mask;
    };
    // printRawQICScores order: a outermost .. d innermost (QuartetScoreComputer.hpp:626-629).  The reference formats
    // serially; here the lines of `nt` consecutive values of a are formatted by host threads into private buffers and
    // written in order, so the file is byte-identical for any thread count.
    auto format_a = [&](int a, std::string& out) {
        char line[1200];
        for (int b = a + 1; b < n; ++b) for (int cc = b + 1; cc < n; ++cc) {
            const int p = R.lca[(size_t)a * n + b], q = R.lca[(size_t)b * n + cc];
            const int dp = R.idepth[p], dq = R.idepth[q];
            for (int d = cc + 1; d < n; ++d) {
                const int dr = R.idepth[R.lca[(size_t)cc * n + d]];
                const int S0 = dp + dr, S2 = std::min(dp, std::min(dq, dr)) + dq;
                if (S0 == S2) continue;                                   // unresolved in the reference tree (:559-562)
                const uint64_t rk = quartet_rank(a, b, cc, d) * 3;
                const uint64_t c0 = get(rk), c1 = get(rk + 1), c2 = get(rk + 2);
                int len;
                if (S0 > S2) len = snprintf(line, sizeof line, "(%s,%s|%s,%s): %g\n", taxon_names[a], taxon_names[b], taxon_names[cc], taxon_names[d], host_log_score(c0, c1, c2));
                else len = snprintf(line, sizeof line, "(%s,%s|%s,%s): %g\n", taxon_names[a], taxon_names[d], taxon_names[b], taxon_names[cc],
                                    host_log_score(c2, c0, c1));              // (u,z|v,w): ab|cd=uz|vw, ac|bd=uv|zw, ad|bc=uw|zv
                if (len >= (int)sizeof line) {                            // very long taxon labels: format again into a large enough buffer
                    std::string big((size_t)len + 1, '\0');
                    if (S0 > S2) snprintf(&big[0], big.size(), "(%s,%s|%s,%s): %g\n", taxon_names[a], taxon_names[b], taxon_names[cc], taxon_names[d], host_log_score(c0, c1, c2));
                    else snprintf(&big[0], big.size(), "(%s,%s|%s,%s): %g\n", taxon_names[a], taxon_names[d], taxon_names[b], taxon_names[cc], host_log_score(c2, c0, c1));
                    out.append(big.data(), (size_t)len);
                } else out.append(line, (size_t)len);
            }
        }
    };
    int nt = (int)std::min<unsigned>(16u, std::max(1u, std::thread::hardware_concurrency()));
    if (const char* env = getenv("QS_HOST_THREADS")) nt = std::max(1, std::min(64, atoi(env)));
    if (n < 24) nt = 1;
    std::vector<std::string> bufs((size_t)nt);
    bool ok = true;
    for (int a0 = 0; a0 < n && ok; a0 += nt) {
        const int cnt = std::min(nt, n - a0);
        HostPool::get().run(cnt, [&](int w) { bufs[(size_t)w].clear(); format_a(a0 + w, bufs[(size_t)w]); });
        for (int w = 0; w < cnt && ok; ++w) ok = bufs[(size_t)w].empty() || fwrite(bufs[(size_t)w].data(), 1, bufs[(size_t)w].size(), f) == bufs[(size_t)w].size();
    }
    ok = (fclose(f) == 0) && ok;
    if (!ok) QS_FAIL(ctx, QS_E_ARG, "short write to %s", path);
    return QS_OK;
}

int qs_last_timing(const qs_ctx* ctx, double* dist_ms, double* count_ms, double* score_ms) {
    if (!ctx) return QS_E_ARG;
    if (dist_ms) *dist_ms = ctx->dist_ms;
    if (count_ms) *count_ms = ctx->count_ms;
    if (score_ms) *score_ms = ctx->score_ms;
    return QS_OK;
}

int qs_launch_count(const qs_ctx* ctx, int64_t* n_launches) {
    if (!ctx || !n_launches) return QS_E_ARG;
    *n_launches = ctx->launches;
    return QS_OK;
}

int qs_tree_classes(const qs_ctx* ctx, int64_t* n_class_a, int64_t* n_class_b) {
    if (!ctx || !n_class_a || !n_class_b) return QS_E_ARG;
    *n_class_a = ctx->n_class_a; *n_class_b = ctx->m - ctx->n_class_a;
    return QS_OK;
}

int qs_measure_alu_peak(qs_ctx* ctx, double* half2_pair_laneops_per_s, double* int32_laneops_per_s) {
    if (!ctx) return QS_E_ARG;
    if (ctx->host_only) QS_FAIL(ctx, QS_E_STATE, "host-only context (QS_DEVICE_NONE): no device work; there is no CPU fallback");
    QS_CUDA(ctx, cudaSetDevice(ctx->device));
    double h = 0, i = 0;
    cudaError_t e = ubench_alu_peak(ctx->num_sms, ctx->stream, &h, &i);
    ctx->launches += 6;
    QS_CUDA(ctx, e);
    if (half2_pair_laneops_per_s) *half2_pair_laneops_per_s = h;
    if (int32_laneops_per_s) *int32_laneops_per_s = i;
    return QS_OK;
}

}  // extern "C"
