// count_rows.cuh — the quartet counting kernel (any n).
//
// Replaces the reference hot loop QuartetCounterLookup::updateQuartetsThreeClades
// (src/QuartetCounterLookup.hpp:66-106) + countQuartets (:197-238).  Instead of enumerating clades per
// tree and scattering increments into an n^4 table, every thread OWNS a fixed 8 x 8 block of quartets,
// keeps their topology counters in registers across all gene trees of its chunk, and decides each
// (quartet, tree) with the four-point condition on the tree's distance matrix D.  Only the matrix ROWS a
// task needs are staged in shared memory, by TMA bulk copies (cp.async.bulk, SASS UBLKCP) through an
// mbarrier pipeline.
//
// Four-point test in "fixed pair" form.  For taxa p,q define G_pq(t) = D[q][t] - D[p][t].  For a tree
// metric, G_pq(u) > G_pq(v)  <=>  D_up + D_vq < D_uq + D_vp  <=>  the tree displays up|vq (the two
// larger pair sums of a tree metric are equal, so one strict inequality decides the topology;
// ties = unresolved or the third topology).  With sorted ids a<b<c<d and the table slots
// 0 = ab|cd, 1 = ac|bd, 2 = ad|bc (src/quartet_lookup_table.hpp:87-111):
//   role X, pair (c,d) fixed:  G_cd(a) > G_cd(b) -> slot 1,   G_cd(a) < G_cd(b) -> slot 2
//   role Y, pair (b,c) fixed:  G_bc(a) > G_bc(d) -> slot 0
//   role Z, pair (a,d) fixed:  G_ad(b) > G_ad(c) -> slot 0   (the same compare with the other two taxa held fixed; Y and Z share slot 0)
// A missing taxon makes D = NaN in its row/column, G = NaN, and every ordered compare false, so such
// (quartet, tree) pairs count nothing — exactly the reference, where absent taxa are never enumerated.
//
// Fully resolved trees need no role Y / Z.  The distance kernel classifies every gene tree: class A = all n
// taxa present and no node of degree > 3, i.e. every quartet is resolved in it, so
// slot0 + slot1 + slot2 = 1 per tree and slot 0 = |A| - slot1_A - slot2_A.  Class-A trees (first in the
// class-sorted order[]) only run role X: 2 compares per quartet x tree instead of 3.
//
// Work items (8 x 8 register blocks = 64 quartets), each kind enumerated in one global order:
//   XO  (c; d; a-block ia < b-block ib): G(a)>G(b) and G(a)<G(b)                               1 item / thread
//   XD  (c; d; diagonal block i, a and b in the same block): all ordered pairs, G(x)>G(y) only;
//       x<y gives slot 1 of (x,y), x>y gives slot 2 of (y,x) — half the cost of a full block   2 items / thread
//   XR  (c; d; a-block x the ragged last b-block of c): as XO over the c mod 8 valid b only           1 item / thread
//   Y   (b; c; whole d-block inside the range; a-block): G(a)>G(d)                                    2 items / thread
//   Z   (d; a; b-block <= c-block between a and d): G(b)>G(c), for the d whose block of 8 the shard's range cuts
//       (and every d of a narrow range): one d per item, so no padding in d                           2 items / thread
// so a thread always carries 64 packed counter registers and issues 64 HSET2 + 64 IMAD.IADD per tree.  A task
// is a run of up to THREADS (512 or 256, see below) thread-items of one kind; consecutive items share matrix rows, and the host records
// the rows a task touches as at most three contiguous row ranges.  Only those rows are staged per tree
// (n = 100: ~3 KB instead of the 21 KB matrix; n = 1000: 4 KB for 32,768 quartets = 0.12 B per evaluation).
// Tasks x tree classes x tree chunks are handed to persistent CTAs through an atomic counter.  Counters are
// flushed once per task straight into the CINT table with red.global.add on 32-bit words (two uint16 / four
// uint8 fields per word; all arithmetic is mod 2^32 and every field ends non-negative and in range, so
// transient carries between fields cancel).
#pragma once
#include <type_traits>
#include "common.cuh"
#include "count_roles.cuh"

namespace qs {

enum { ITEM_XO = 0, ITEM_XD = 1, ITEM_Y = 2, ITEM_XR = 3, ITEM_Z = 4, ITEM_KINDS = 5 };

struct RowTask {
    int32_t kind;           // ITEM_*
    int32_t ne;             // number of items (<= threads * items-per-thread)
    int64_t e0;             // first item in the kind's global enumeration
    int32_t rstart[3];      // up to three contiguous ranges of matrix rows the items touch ...
    int32_t rcount[3];      // ... staged back to back in shared memory
};

// prefix tables of the three enumerations (device pointers in the kernel arguments, host vectors in qscuda.cu)
struct EnumTables {
    const int64_t* PXO;     // [n+1] XO items with c' < c
    const int64_t* PXD;     // [n+1] XD items with c' < c
    const int64_t* PY;      // [n+1] Y items with b' < b
    const int64_t* CD;      // [n+1] d-blocks of role Y over c' < c
    const int64_t* PXR;     // [n+1] XR items with c' < c
    const int64_t* PZ;      // [nz * n + 1] Z items before (d, a), d = cr_zd(index / n), a = index % n
    int z_first, z_last;    // role Y owns the d-blocks inside [z_first, z_last) (multiples of 8, or both = d_end), role Z the d of [d_begin, z_first) and [z_last, d_end)
    int xo_diag;            // 1: diagonal blocks are ordinary XO items (ia <= ib) and there are no XD items (large n, where
                            //    XD tasks would be starved of items by the row budget and diagonal blocks are a few % of the work)
};

struct CountRowsArgs {
    const __half* D;            // [m][n][n_pad] fp16, NaN = missing, indexed by ORIGINAL tree index
    const int32_t* order;       // [m] class-sorted tree order: class A first
    const int32_t* n_class_a;   // device scalar |A|
    const RowTask* tasks;       // X tasks [0, n_x), then Y tasks [n_x, n_x + n_y)
    EnumTables E;
    int n_x, n_y;
    int chunk_trees;            // target trees per chunk (<= QS_MAX_CHUNK_TREES)
    int* task_counter;          // zeroed by the caller
    void* table;                // CINT [(rank - rank_base)][3], initialised by qs_table_init_kernel
    int cint_bytes;
    uint64_t rank_base;
    int n, n_pad, m;
    int d_begin, d_end;         // shard: quartets with d in [d_begin, d_end)
    uint32_t row_bytes;         // n_pad * 2
    uint32_t ring_bytes;        // shared memory available to the staging ring (every task sizes its own stages from it)
    int max_tps, max_stages;    // caps of the ring geometry (<= 32 trees per stage: one lane issues the copies of one tree; <= CR_MAX_STAGES)
};

#ifndef CR_UNROLL_PF
#define CR_UNROLL_PF 1          // trees per iteration of the tree loop: must stay 1 with integer counters (count_roles.cuh); 2 was 2 % faster with fp16 counters
#endif
#ifndef CR_UNROLL_HALVES
#define CR_UNROLL_HALVES 1
#endif

constexpr int kUnrollPf = CR_UNROLL_PF, kUnrollHalves = CR_UNROLL_HALVES;
// CTA shapes of the counting kernel (template parameters THREADS, CTAS per SM), always 16 warps per SM:
//   512 x 1 : one CTA owns the SM's whole staging ring — best while a task's rows are a large part of the matrix (n <= 112)
//   256 x 2 : two independent CTAs per SM desynchronise the stage hand-overs — 5-9 % faster from n = 128 up
// (profiles/r01_u_variants.txt).  A task holds THREADS thread-items, so the host plan is built for the shape in use.
constexpr int CR_THREADS_BIG = 512, CR_THREADS_SMALL = 256;
__host__ __device__ constexpr int cr_ctas_per_sm(int threads) { return CR_THREADS_BIG / threads; }
constexpr int CR_MAX_STAGES = 8;
constexpr int CR_MAX_TPS = 32;        // one lane issues the copies of one tree
constexpr int CR_SMEM_HEADER = 128 + 4 * 4096;     // barriers + done counters + the chunk's tree ids (QS_MAX_CHUNK_TREES)

// ---- the enumerations (shared by the host task builder and the kernel) ---------------------------------------
__host__ __device__ __forceinline__ int cr_nxd(int c, int xo_diag) { return (!xo_diag && c >= 2) ? ((c - 2) >> 3) + 1 : 0; }   // blocks with two taxa below c
// b-blocks below c: nfull of them hold 8 taxa below c; when c is not a multiple of 8 one more, ragged, holds c & 7.  XO items pair
// an a-block with a FULL b-block (ia < ib; ia <= ib when the diagonal blocks are XO items too); the ragged b-block's items are a kind
// of their own, XR, whose tree loop only runs over the c & 7 valid b — with them inside XO every (c,d) paid a whole block row
// for a partly empty one: 15 % of all compares at n = 100 (enumeration efficiency 0.85 -> 0.97).
#ifndef CR_NO_XR
#define CR_NO_XR 0              // tuning build: 1 = no XR kind, the ragged b-block is an ordinary XO block again
#endif
__host__ __device__ __forceinline__ int cr_nfull(int c) { return CR_NO_XR ? (c + 7) >> 3 : c >> 3; }
__host__ __device__ __forceinline__ int cr_nxo(int c, int xo_diag) {
    const int nf = cr_nfull(c);
    return nf * (nf - 1) / 2 + ((xo_diag && c >= 2) ? nf : 0);
}
__host__ __device__ __forceinline__ int cr_nxr(int c, int xo_diag) {
    if (CR_NO_XR || (c & 7) == 0 || c < 2) return 0;
    return cr_nfull(c) + ((xo_diag && (c & 7) >= 2) ? 1 : 0);          // ia = 0 .. nfull-1 (+ the ragged diagonal block, if it holds two taxa)
}
__host__ __device__ __forceinline__ int cr_dlo(int c, int d_begin) { return c + 1 > d_begin ? c + 1 : d_begin; }
// Role Y works on whole d-blocks of 8; the d of a block that the shard's range [d_begin, d_end) cuts belong to role Z instead (one d
// per item, 8 x 8 blocks of (b,c)): as Y blocks both neighbours of a shard boundary paid the whole block (13 % of the counting time of
// 8 shards at n = 500, profiles/r02_v_shards_*.txt), and the last block of an n that is no multiple of 8 ran half empty.
__host__ __device__ __forceinline__ int cr_ndb(int c, int z_first, int z_last) {                                   // d-blocks of role Y above c
    const int k0 = max((c + 1) >> 3, z_first >> 3);
    return (z_last >> 3) > k0 ? (z_last >> 3) - k0 : 0;
}
__host__ __device__ __forceinline__ int cr_nzd(int d_begin, int d_end, int z_first, int z_last) { return (z_first - d_begin) + (d_end - z_last); }
__host__ __device__ __forceinline__ int cr_zd(int i, int d_begin, int z_first, int z_last) { return i < z_first - d_begin ? d_begin + i : z_last + (i - (z_first - d_begin)); }
// Z items of the pair (a,d): the blocks (ib <= ic) of the taxa strictly between a and d, j = ic'(ic'+1)/2 + ib' counted from block (a+1) >> 3
__host__ __device__ __forceinline__ int cr_nz(int a, int d) {
    if (d - a < 3) return 0;
    const int k = ((d - 1) >> 3) - ((a + 1) >> 3) + 1;
    return k * (k + 1) / 2;
}
// largest x in [lo,hi) with P[x] <= key  (P non-decreasing, P[lo] <= key)
__host__ __device__ __forceinline__ int cr_ub(const int64_t* P, int lo, int hi, int64_t key) {
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (P[mid] <= key) lo = mid; else hi = mid;
    }
    return lo;
}
// item e of kind XO / XD -> (c, d, block index j inside the (c,d) pair)
__host__ __device__ __forceinline__ void cr_decode_x(const int64_t* P, int kind, int xo_diag, int64_t e, int n, int d_begin, int& c, int& d, int& j) {
    c = cr_ub(P, 0, n, e);
    const int64_t r = e - P[c];
    const int cnt = kind == ITEM_XO ? cr_nxo(c, xo_diag) : kind == ITEM_XR ? cr_nxr(c, xo_diag) : cr_nxd(c, xo_diag);
    d = cr_dlo(c, d_begin) + (int)(r / cnt);
    j = (int)(r % cnt);
}
// item e of kind Y -> (b, c, a-block ia, d-block id)
__host__ __device__ __forceinline__ void cr_decode_y(const EnumTables& E, int64_t e, int n, int& b, int& c, int& ia, int& id) {
    b = cr_ub(E.PY, 0, n, e);
    const int64_t r = e - E.PY[b];
    const int na = (b + 7) >> 3;
    ia = (int)(r % na);
    const int64_t target = E.CD[b + 1] + r / na;
    c = cr_ub(E.CD, b + 1, n, target);
    id = max((c + 1) >> 3, E.z_first >> 3) + (int)(target - E.CD[c]);
}
// item e of kind Z -> (d, a, block pair ib <= ic)
__host__ __device__ __forceinline__ void cr_decode_z(const EnumTables& E, int64_t e, int n, int d_begin, int d_end, int& d, int& a, int& ib, int& ic) {
    const int z = cr_ub(E.PZ, 0, cr_nzd(d_begin, d_end, E.z_first, E.z_last) * n, e);
    d = cr_zd(z / n, d_begin, E.z_first, E.z_last);
    a = z % n;
    const int j = (int)(e - E.PZ[z]);
    int t = (int)((sqrtf(8.f * (float)j + 1.f) - 1.f) * 0.5f);
    while (t * (t + 1) / 2 > j) --t;
    while ((t + 1) * (t + 2) / 2 <= j) ++t;
    const int k0 = (a + 1) >> 3;
    ic = k0 + t;
    ib = k0 + (j - t * (t + 1) / 2);
}

// add v (mod 2^32, may stand for a negative number) to element `elem` of the CINT table
__device__ __forceinline__ void table_red(void* table, int cint_bytes, uint64_t elem, uint32_t v) {
    if (v == 0u) return;
    switch (cint_bytes) {
        case 1: atomicAdd(reinterpret_cast<uint32_t*>(table) + (elem >> 2), v << (8u * (uint32_t)(elem & 3))); break;
        case 2: atomicAdd(reinterpret_cast<uint32_t*>(table) + (elem >> 1), v << (16u * (uint32_t)(elem & 1))); break;
        case 4: atomicAdd(reinterpret_cast<uint32_t*>(table) + elem, v); break;
        default: atomicAdd(reinterpret_cast<unsigned long long*>(table) + elem, (unsigned long long)(long long)(int32_t)v); break;
    }
}

struct RowPipe {
    uint64_t* full;             // [CR_MAX_STAGES] tx barriers
    int* done;                  // [CR_MAX_STAGES] warps finished with the stage
    int32_t* trees;             // [QS_MAX_CHUNK_TREES] tree ids of the current chunk (order[t0..t1)), staged once per task
    unsigned char* bufs;
    uint32_t phase;             // bit s = parity to wait for on full[s]
};

// Stream the trees [t0,t1) of the class-sorted order through the pipeline, up to CR_MAX_TPS trees per stage; stage(base, nt,
// slot) runs the caller's tree loop over the nt trees staged at base, base + slot, ...  Warps run
// independently: the last warp to finish a stage refills it (no CTA-wide barrier inside the loop).  The chunk's tree
// ids are copied to shared memory once per task, so no warp waits on global memory inside the loop.
// per-stage tree loops.  stage_pf: one operand set per tree, the next tree's operands are fetched before the math of the
// current one (the LDS latency otherwise shows up as a short-scoreboard stall at the top of every tree: profiles/r01_e_*).
// stage_halves: two half-cost items per thread (role Y); the operands of item B are in flight during the math of item A
// and those of the next tree's item A during the math of item B — same register footprint as stage_pf.
template <class LD, class MT>
__device__ __forceinline__ void stage_pf(const unsigned char* base, int nt, uint32_t slot, LD&& ld, MT&& mt) {
    auto cur = ld(base);
#pragma unroll kUnrollPf
    for (int tt = 0; tt < nt; ++tt) {
        if (tt + 1 < nt) base += slot;               // (the last tree of a stage re-reads itself: no branch around the loads)
        auto nx = ld(base);
        mt(cur);
        cur = nx;
    }
}
template <class LDA, class LDB, class MTA, class MTB>
__device__ __forceinline__ void stage_halves(const unsigned char* base, int nt, uint32_t slot, LDA&& lda, LDB&& ldb, MTA&& mta, MTB&& mtb) {
    auto ca = lda(base);
#pragma unroll kUnrollHalves
    for (int tt = 0; tt < nt; ++tt) {
        auto cb = ldb(base);
        mta(ca);
        if (tt + 1 < nt) base += slot;
        ca = lda(base);
        mtb(cb);
    }
}

template <int THREADS, class STAGE>
__device__ __forceinline__ void stream_rows(const CountRowsArgs& a, RowPipe& P, const RowTask& T, int t0, int t1, STAGE&& stage) {
    const int tid = threadIdx.x, lane = tid & 31;
    const int ntrees = t1 - t0;
    const size_t tree_elems = (size_t)a.n * a.n_pad;
    // pipeline geometry of THIS task: a staged tree takes only the rows the task touches (a few KB for most tasks, the
    // whole matrix only for the few tasks with tiny c), so a stage holds up to CR_MAX_TPS trees.  Stages sized for the
    // worst task (3 trees each, round 1e) cost 12 % at cfg2 (profiles/r01_f_*, r01_g_*).
    const uint32_t slot = (uint32_t)(T.rcount[0] + T.rcount[1] + T.rcount[2]) * a.row_bytes;            // bytes staged per tree
    // the hand-over between stages (mbarrier wait, shared-memory atomic, refill) costs ~600 clk per stage and the ring
    // depth does not matter beyond 2 (profiles/r01_p_sweep_*): as many trees per stage as two stages allow
    int tps = a.max_tps;
    while (tps > 1 && (uint32_t)(2 * tps) * slot > a.ring_bytes) --tps;
    const int n_stages = min(a.max_stages, (int)(a.ring_bytes / ((uint32_t)tps * slot)));
    const uint32_t stage_bytes = (uint32_t)tps * slot;
    const int nst = (ntrees + tps - 1) / tps;
    // lanes 0..tps-1 of the calling warp copy the rows of the trees of stage st
    auto issue = [&](int st, int buf) {
        const int nt = min(tps, ntrees - st * tps);
        if (lane == 0) mbar_expect_tx(&P.full[buf], (uint32_t)nt * slot);
        __syncwarp();
        if (lane < nt) {
            unsigned char* dst = P.bufs + (size_t)buf * stage_bytes + (size_t)lane * slot;
            const __half* src = a.D + (size_t)P.trees[st * tps + lane] * tree_elems;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if (T.rcount[k] > 0) bulk_g2s(dst, src + (size_t)T.rstart[k] * a.n_pad, (uint32_t)T.rcount[k] * a.row_bytes, &P.full[buf]);
                dst += (size_t)T.rcount[k] * a.row_bytes;
            }
        }
    };
    __syncthreads();                 // the previous task's readers are done with every stage and with P.trees
    for (int i = tid; i < ntrees; i += THREADS) P.trees[i] = a.order[t0 + i];
    if (tid < CR_MAX_STAGES) P.done[tid] = 0;
    __syncthreads();
    if (tid < 32)
        for (int s = 0; s < n_stages && s < nst; ++s) issue(s, s);
    int buf = 0;
    for (int st = 0; st < nst; ++st) {
        mbar_wait(&P.full[buf], (P.phase >> buf) & 1u);
        P.phase ^= (1u << buf);
        const int nt = min(tps, ntrees - st * tps);
        stage(P.bufs + (size_t)buf * stage_bytes, nt, slot);
        __syncwarp();
        int last = 0;
        if (lane == 0) last = (atomicAdd(&P.done[buf], 1) == THREADS / 32 - 1);
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {                                  // last warp out refills the stage
            if (lane == 0) P.done[buf] = 0;
            if (st + n_stages < nst) issue(st + n_stages, buf);
        }
        buf = (buf + 1 == n_stages) ? 0 : buf + 1;
    }
}

// shared-memory byte offset of matrix row `row` inside a staged tree of task T
__device__ __forceinline__ uint32_t cr_row_off(const RowTask& T, int row, uint32_t row_bytes) {
    int slot = row - T.rstart[0];
    if (slot < 0 || slot >= T.rcount[0]) {
        slot = row - T.rstart[1];
        if (slot >= 0 && slot < T.rcount[1]) slot += T.rcount[0];
        else slot = row - T.rstart[2] + T.rcount[0] + T.rcount[1];
    }
    return (uint32_t)slot * row_bytes;
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, cr_ctas_per_sm(THREADS)) qs_count_rows_kernel(const CountRowsArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ int s_task;
    RowPipe P;
    P.full = reinterpret_cast<uint64_t*>(smem);
    P.done = reinterpret_cast<int*>(smem + 64);
    P.trees = reinterpret_cast<int32_t*>(smem + 128);
    P.bufs = smem + CR_SMEM_HEADER;
    P.phase = 0;
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < CR_MAX_STAGES; ++s) mbar_init(&P.full[s], 1);
        fence_mbar_init();
    }
    // task space: class A runs the X tasks only, class B runs X and Y; each class in chunks of ~chunk_trees trees
    const int mA = *a.n_class_a, mB = a.m - mA;
    const int chA = (mA + a.chunk_trees - 1) / a.chunk_trees, chB = (mB + a.chunk_trees - 1) / a.chunk_trees;
    const long long nA_tasks = (long long)chA * a.n_x, nB_tasks = (long long)chB * (a.n_x + a.n_y);
    const bool sub0 = mB > 0;       // mixed input: class-A role-X hits are subtracted from slot 0 here; otherwise qs_table_finalize derives slot 0

    while (true) {
        __syncthreads();
        if (tid == 0) s_task = atomicAdd(a.task_counter, 1);
        __syncthreads();
        const long long id = s_task;
        if (id >= nA_tasks + nB_tasks) break;
        int cls, chunk, base_task, nch, lo, len;
        // chunk-major: CTAs that run at the same time work on different tasks, i.e. flush into different parts of the table (task-major
        // order — the chunks of one task side by side — made every flush contend for the same L2 lines: +16 % at cfg2, +38 % at n = 200,
        // profiles/r02_g_sweep_*.txt); inside a chunk the tasks come in table order, longest kinds first (qscuda.cu build_row_tasks)
        if (id < nA_tasks) { cls = 0; chunk = (int)(id / a.n_x); base_task = (int)(id % a.n_x); nch = chA; lo = 0; len = mA; }
        else { const long long r = id - nA_tasks; cls = 1; chunk = (int)(r / (a.n_x + a.n_y)); base_task = (int)(r % (a.n_x + a.n_y)); nch = chB; lo = mA; len = mB; }
        const int per = (len + nch - 1) / nch;                     // balanced chunks, each <= chunk_trees
        const int t0 = lo + chunk * per, t1 = min(lo + len, t0 + per);
        if (t0 >= t1) continue;
        const RowTask T = a.tasks[base_task];
        const bool cls_a = (cls == 0);
        const uint32_t rb = a.row_bytes;

        if (T.kind == ITEM_XO || T.kind == ITEM_XR) {
            const bool ragged = T.kind == ITEM_XR;
            const bool valid = tid < T.ne;
            int c = 2, d = 3, j = 0;
            if (valid) cr_decode_x(ragged ? a.E.PXR : a.E.PXO, T.kind, a.E.xo_diag, T.e0 + tid, a.n, a.d_begin, c, d, j);
            int ib = 1, ia = 0;
            if (valid) {
                const int nf = cr_nfull(c);
                if (ragged) { ib = nf; ia = j; }                                  // ragged b-block nf: a-blocks 0 .. nf-1, then (xo_diag) the diagonal block
                else {
                    const int noff = nf * (nf - 1) / 2;                           // off-diagonal blocks first: j = ib(ib-1)/2 + ia, ia < ib
                    if (j < noff) {
                        ib = (int)((1.f + sqrtf(1.f + 8.f * (float)j)) * 0.5f);
                        while (ib * (ib - 1) / 2 > j) --ib;
                        while ((ib + 1) * ib / 2 <= j) ++ib;
                        ia = j - ib * (ib - 1) / 2;
                    } else ia = ib = j - noff;                                    // xo_diag: then the diagonal blocks
                }
            }
            // valid b of the b-block: 8, or c & 7 in the ragged one; the tree loop runs over the largest count in the warp
            // (a task's items nearly always share c, so it is the same for all lanes)
            int jmax = 8;
            if (ragged) jmax = __reduce_max_sync(0xffffffffu, valid ? (c & 7) : 1);
            // pair (p,q) = (c,d)
            const uint32_t rp = valid ? cr_row_off(T, c, rb) : 0u, rq = valid ? cr_row_off(T, d, rb) : 0u;
            const uint32_t oPu = rp + ia * 16u, oQu = rq + ia * 16u, oPv = rp + ib * 16u, oQv = rq + ib * 16u;
            XCounters x; zero(x);
            auto run = [&](auto jm) {
                constexpr int JM = decltype(jm)::value;
                stream_rows<THREADS>(a, P, T, t0, t1, [&](const unsigned char* base, int nt, uint32_t slot) {
                    // the prefetched operand set of a tree is G itself (8 registers), not the four raw row segments (16): -2.5 % at
                    // cfg2, -5 % at n = 200 (profiles/r02_e_count_variants.txt)
                    stage_pf(base, nt, slot,
                             [&](const unsigned char* s) { BlockG g; sub4(g.u, lds128(s, oQu), lds128(s, oPu)); sub4(g.v, lds128(s, oQv), lds128(s, oPv)); return g; },
                             [&](const BlockG& g) { step_gt_lt_g<JM>(x, g); });
                });
            };
            switch (jmax) {
                case 1: run(std::integral_constant<int, 1>()); break;
                case 2: run(std::integral_constant<int, 2>()); break;
                case 3: run(std::integral_constant<int, 3>()); break;
                case 4: run(std::integral_constant<int, 4>()); break;
                case 5: run(std::integral_constant<int, 5>()); break;
                case 6: run(std::integral_constant<int, 6>()); break;
                case 7: run(std::integral_constant<int, 7>()); break;
                default: run(std::integral_constant<int, 8>()); break;
            }
            if (valid) {
                const uint64_t rcd = binom4((uint64_t)d) + binom3((uint64_t)c) - a.rank_base;
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    const int b = ib * 8 + jj;
                    if (b >= c) continue;
                    const uint64_t eb = (rcd + (uint64_t)b * (b - 1) / 2 + ia * 8) * 3;
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        uint32_t g0, g1, l0, l1;
                        decode(x.gt[jj][p], g0, g1);
                        decode(x.lt[jj][p], l0, l1);
                        if (ia * 8 + 2 * p >= b) { g0 = 0; l0 = 0; }              // diagonal block (xo_diag): only a < b
                        if (ia * 8 + 2 * p + 1 >= b) { g1 = 0; l1 = 0; }
                        const uint64_t e = eb + (uint64_t)(2 * p) * 3;
                        table_red(a.table, a.cint_bytes, e + 1, g0); table_red(a.table, a.cint_bytes, e + 2, l0);
                        table_red(a.table, a.cint_bytes, e + 4, g1); table_red(a.table, a.cint_bytes, e + 5, l1);
                        if (cls_a && sub0) { table_red(a.table, a.cint_bytes, e, 0u - (g0 + l0)); table_red(a.table, a.cint_bytes, e + 3, 0u - (g1 + l1)); }
                    }
                }
            }
        } else if (T.kind == ITEM_XD) {
            const bool vA = 2 * tid < T.ne, vB = 2 * tid + 1 < T.ne;
            int cA = 2, dA = 3, jA = 0, cB = 2, dB = 3, jB = 0;
            if (vA) cr_decode_x(a.E.PXD, ITEM_XD, a.E.xo_diag, T.e0 + 2 * tid, a.n, a.d_begin, cA, dA, jA);
            if (vB) cr_decode_x(a.E.PXD, ITEM_XD, a.E.xo_diag, T.e0 + 2 * tid + 1, a.n, a.d_begin, cB, dB, jB);
            const uint32_t oAp = (vA ? cr_row_off(T, cA, rb) : 0u) + jA * 16u, oAq = (vA ? cr_row_off(T, dA, rb) : 0u) + jA * 16u;
            const uint32_t oBp = (vB ? cr_row_off(T, cB, rb) : 0u) + jB * 16u, oBq = (vB ? cr_row_off(T, dB, rb) : 0u) + jB * 16u;
            GCounters ga, gb; zero(ga); zero(gb);
            stream_rows<THREADS>(a, P, T, t0, t1, [&](const unsigned char* base, int nt, uint32_t slot) {
                stage_pf(base, nt, slot,
                         [&](const unsigned char* s) { return BlockRows{lds128(s, oAp), lds128(s, oAq), lds128(s, oBp), lds128(s, oBq)}; },
                         [&](const BlockRows& r) { step_gt_diag(ga, r.pu, r.qu); step_gt_diag(gb, r.pv, r.qv); });
            });
            auto flush = [&](int c, int d, int blk, const GCounters& g) {
                const int x0 = blk * 8;
                const uint64_t rcd = binom4((uint64_t)d) + binom3((uint64_t)c) - a.rank_base;
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    const int y = x0 + jj;                          // the "v" taxon
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        uint32_t h[2];
                        decode(g.gt[jj][p], h[0], h[1]);
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int x = x0 + 2 * p + e;          // the "u" taxon: counted G(x) > G(y)
                            if (x == y || max(x, y) >= c) continue;
                            const int lo2 = min(x, y), hi2 = max(x, y);
                            const uint64_t el = (rcd + (uint64_t)hi2 * (hi2 - 1) / 2 + lo2) * 3;
                            table_red(a.table, a.cint_bytes, el + ((x < y) ? 1 : 2), h[e]);   // x<y: G(a)>G(b) slot 1;  x>y: G(b)>G(a) slot 2
                            if (cls_a && sub0) table_red(a.table, a.cint_bytes, el, 0u - h[e]);
                        }
                    }
                }
            };
            if (vA) flush(cA, dA, jA, ga);
            if (vB) flush(cB, dB, jB, gb);
        } else if (T.kind == ITEM_Z) {   // pair (p,q) = (a,d), u = b-block, v = c-block: G(b) > G(c) -> ab|cd (slot 0)
            const bool vA = 2 * tid < T.ne, vB = 2 * tid + 1 < T.ne;
            int dA = 3, aA = 0, ibA = 0, icA = 0, dB = 3, aB = 0, ibB = 0, icB = 0;
            if (vA) cr_decode_z(a.E, T.e0 + 2 * tid, a.n, a.d_begin, a.d_end, dA, aA, ibA, icA);
            if (vB) cr_decode_z(a.E, T.e0 + 2 * tid + 1, a.n, a.d_begin, a.d_end, dB, aB, ibB, icB);
            const uint32_t pA = vA ? cr_row_off(T, aA, rb) : 0u, qA = vA ? cr_row_off(T, dA, rb) : 0u;
            const uint32_t pB = vB ? cr_row_off(T, aB, rb) : 0u, qB = vB ? cr_row_off(T, dB, rb) : 0u;
            const uint32_t oApu = pA + ibA * 16u, oAqu = qA + ibA * 16u, oApv = pA + icA * 16u, oAqv = qA + icA * 16u;
            const uint32_t oBpu = pB + ibB * 16u, oBqu = qB + ibB * 16u, oBpv = pB + icB * 16u, oBqv = qB + icB * 16u;
            GCounters ga, gb; zero(ga); zero(gb);
            stream_rows<THREADS>(a, P, T, t0, t1, [&](const unsigned char* base, int nt, uint32_t slot) {
                stage_halves(base, nt, slot,
                             [&](const unsigned char* s) { return BlockRows{lds128(s, oApu), lds128(s, oAqu), lds128(s, oApv), lds128(s, oAqv)}; },
                             [&](const unsigned char* s) { return BlockRows{lds128(s, oBpu), lds128(s, oBqu), lds128(s, oBpv), lds128(s, oBqv)}; },
                             [&](const BlockRows& r) { step_gt(ga, r); }, [&](const BlockRows& r) { step_gt(gb, r); });
            });
            auto flush = [&](int d, int a0, int ib, int ic, const GCounters& g) {
                const uint64_t rd = binom4((uint64_t)d) - a.rank_base + (uint64_t)a0;
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    const int c = ic * 8 + jj;
                    if (c >= d) continue;
                    const uint64_t rc = rd + binom3((uint64_t)c);
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        uint32_t h[2];
                        decode(g.gt[jj][p], h[0], h[1]);
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int b = ib * 8 + 2 * p + e;
                            if (b > a0 && b < c) table_red(a.table, a.cint_bytes, (rc + (uint64_t)b * (b - 1) / 2) * 3, h[e]);
                        }
                    }
                }
            };
            if (vA) flush(dA, aA, ibA, icA, ga);
            if (vB) flush(dB, aB, ibB, icB, gb);
        } else {   // ITEM_Y: pair (p,q) = (b,c), u = a-block, v = d-block
            const bool vA = 2 * tid < T.ne, vB = 2 * tid + 1 < T.ne;
            int bA = 1, cA = 2, iaA = 0, idA = 0, bB = 1, cB = 2, iaB = 0, idB = 0;
            if (vA) cr_decode_y(a.E, T.e0 + 2 * tid, a.n, bA, cA, iaA, idA);
            if (vB) cr_decode_y(a.E, T.e0 + 2 * tid + 1, a.n, bB, cB, iaB, idB);
            const uint32_t pA = vA ? cr_row_off(T, bA, rb) : 0u, qA = vA ? cr_row_off(T, cA, rb) : 0u;
            const uint32_t pB = vB ? cr_row_off(T, bB, rb) : 0u, qB = vB ? cr_row_off(T, cB, rb) : 0u;
            const uint32_t oApu = pA + iaA * 16u, oAqu = qA + iaA * 16u, oApv = pA + idA * 16u, oAqv = qA + idA * 16u;
            const uint32_t oBpu = pB + iaB * 16u, oBqu = qB + iaB * 16u, oBpv = pB + idB * 16u, oBqv = qB + idB * 16u;
            GCounters ga, gb; zero(ga); zero(gb);
            stream_rows<THREADS>(a, P, T, t0, t1, [&](const unsigned char* base, int nt, uint32_t slot) {
                stage_halves(base, nt, slot,
                             [&](const unsigned char* s) { return BlockRows{lds128(s, oApu), lds128(s, oAqu), lds128(s, oApv), lds128(s, oAqv)}; },
                             [&](const unsigned char* s) { return BlockRows{lds128(s, oBpu), lds128(s, oBqu), lds128(s, oBpv), lds128(s, oBqv)}; },
                             [&](const BlockRows& r) { step_gt(ga, r); }, [&](const BlockRows& r) { step_gt(gb, r); });
            });
            auto flush = [&](int b, int c, int ia, int id, const GCounters& g) {
                const int dlo = max(c + 1, a.d_begin);
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    const int d = id * 8 + jj;
                    if (d < dlo || d >= a.d_end) continue;
                    const uint64_t eb = (binom4((uint64_t)d) + binom3((uint64_t)c) + (uint64_t)b * (b - 1) / 2 - a.rank_base + ia * 8) * 3;
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        uint32_t h0, h1;
                        decode(g.gt[jj][p], h0, h1);
                        const int a0 = ia * 8 + 2 * p;
                        if (a0 < b) table_red(a.table, a.cint_bytes, eb + (uint64_t)(2 * p) * 3, h0);
                        if (a0 + 1 < b) table_red(a.table, a.cint_bytes, eb + (uint64_t)(2 * p + 1) * 3, h1);
                    }
                }
            };
            if (vA) flush(bA, cA, iaA, idA, ga);
            if (vB) flush(bB, cB, iaB, idB, gb);
        }
    }
}

// before counting: zero the table; with both tree classes present slot 0 starts at |A| (the class-A role-X hits
// are subtracted from it by the counting kernel, so every field stays >= 0 at the end of the word arithmetic)
template <typename CINT>
__global__ void qs_table_init_kernel(CINT* __restrict__ table, uint64_t n_entries, const int32_t* __restrict__ n_class_a, int m) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const int mA = *n_class_a;
    const CINT s0 = (mA > 0 && mA < m) ? (CINT)mA : (CINT)0;
    for (; i < n_entries * 3; i += stride) table[i] = (i % 3 == 0) ? s0 : (CINT)0;
}

// after counting, when EVERY tree is class A: slot 0 = |A| - slot 1 - slot 2; thread = one table entry
template <typename CINT>
__global__ void qs_table_finalize_kernel(CINT* __restrict__ table, uint64_t n_entries, const int32_t* __restrict__ n_class_a, int m) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const int mA = *n_class_a;
    if (mA != m) return;
    for (; i < n_entries; i += stride) {
        CINT* t = table + i * 3;
        t[0] = (CINT)((CINT)mA - t[1] - t[2]);
    }
}

}  // namespace qs
