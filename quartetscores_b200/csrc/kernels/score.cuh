// score.cuh — scan of the count table: QIC per quartet, aggregated per reference inner-node pair, then per edge.
//
// Replaces QuartetScoreComputer::processNodePair / computeQuartetScoresBifurcating /
// computeQuartetScoresMultifurcating (src/QuartetScoreComputer.hpp:379-593).  The reference walks, for every quartet,
// the path between the two inner nodes u,v that the quartet's central edge path connects and takes a critical section
// per edge.  Here the scan is quartet-centric:
//
//   * taxon ids follow the reference tree's planar leaf order, so for sorted a<b<c<d only the non-crossing pairings can
//     be the reference topology.  With p = lca(a,b), q = lca(b,c), r = lca(c,d) and their depths:
//     S0 = dp+dr, S2 = min(dp,dq,dr)+dq;  S0 > S2 -> ab|cd (slot 0), ends u = deeper(p,q), v = deeper(q,r);
//     S2 > S0 -> ad|bc (slot 2), ends u = q, v = deeper(p,r);  equal -> unresolved in a multifurcating reference, skipped
//     (:559-562).
//   * per unordered pair {u,v}: three integer sums (q_ref, q_slot1, q_other) for QP-IC/EQP-IC (:429-431,472) and the quartet
//     with minimal QIC — its count triple, so that the host can re-evaluate log_score with the reference's libm and operation
//     order — for LQ-IC (:432-454).
//   * the per-edge minima over paths are taken on the device too (qs_edge_*_kernel below); the host only evaluates
//     log_score for the ~3 x edges selected triples / sums.
//
// qs_scan_rows_kernel (round 2; replaces the per-thread run scan of round 1, which read 48 bytes at a time from rows
// ~270 KB apart and issued 0.4 global atomics per quartet — 5 % of the HBM roofline).  A CTA owns one (c,d): the
// C(c,2) entries (b,a) of that pair are ONE contiguous piece of the table (rank = C(d,4)+C(c,3)+C(b,2)+a,
// src/quartet_lookup_table.hpp:141-168).  A thread owns a row b (a = 0..b-1, contiguous) paired with the row c-1-b, so
// all threads carry the same load, and the rows a warp reads at the same time are neighbours in memory.  The row is
// streamed through a private shared-memory ring by 16-byte cp.async copies (four 48-byte groups = 8 entries each in
// flight per thread, no register staging) and consumed 12 bytes (two entries) at a time.
//   Along a row the pair key only changes where lca(a,b) changes: the host stores every row of the reference LCA matrix
// run-length encoded (run_off/run_end/run_pd), so the key logic runs once per RUN, and the per-entry work is three
// integer adds and an fp32 QIC estimate.  A run is flushed into CTA-level accumulators in shared memory, indexed by ONE
// inner node because the other end of the key is implied by (c,d):
//     key (x, r)           r = lca(c,d)                 -> accR[x]        x = p or q
//     key (p, q)  p below q, q = lca(b,c) = the ancestor of p that joins c -> accP[p]  (q recorded beside it)
//     key (q, p)  p above q, both ancestors of c deeper than r   -> accQ[depth(q)-dr-1][depth(p)-dr-1] (32 levels; deeper ones go
//                                                                   straight to global memory)
// Inner nodes are numbered by their first gap in the planar leaf order (qscuda.cu build_reference), so every index
// touched by the CTA of (c,d) is < c and only c entries are zeroed and flushed.  One global atomic per touched key and CTA
// instead of one per run: n^3 instead of ~0.4 C(n,4) atomics.
//   LQ-IC selection without an fp64 log per quartet: a run keeps the minimum fp32 estimate (error < 5e-7); only if that
// could beat the pair's current exact minimum (pair_score, with a conservative fp32 copy in shared memory) the run is
// walked again by a cold function that evaluates the quartets within the error margin in fp64.
#pragma once
#include "common.cuh"
#include <type_traits>

namespace qs {

constexpr long long QS_I64_NONE = 0x7fffffffffffffffLL;       // "no triple" / "no score": largest int64, survives a MIN all-reduce
constexpr unsigned long long QS_TRIPLE_NONE = (unsigned long long)QS_I64_NONE;
constexpr int QS_TRIPLE_BITS = 21;                             // three counts per 63-bit triple: qs_count refuses m * count_scale >= 2^21

struct ScoreArgs {
    const void* table;               // CINT [(rank - rank_base)][3]
    uint64_t rank_base;
    const uint16_t* lca;             // [n][n] inner index of lca(leaf x, leaf y)
    const uint16_t* idepth;          // [I] depth of an inner node (by inner index)
    const uint32_t* run_off;         // [n+1] row b of lca[][] restricted to a < b, run-length encoded: runs run_off[b] .. run_off[b+1]
    const uint32_t* run_end;         //   exclusive end (in a) of the run
    const uint32_t* run_pd;          //   inner index of lca(a,b) | its depth << 16
    const int32_t* inner_parent;     // [I] inner index of the parent inner node, -1 for the root
    const int32_t* leaf_parent;      // [n] inner index of the parent of leaf x
    unsigned long long* pair_sums;   // [I*I][3]  key = min(u,v) * I + max(u,v)
    long long* pair_best;            // [I*I] packed triple of the min-QIC quartet, QS_I64_NONE = none
    long long* pair_score;           // [I*I] order-preserving int64 image of the EXACT (device fp64) score of pair_best
    unsigned long long* scratch;     // accumulators in global memory (one region per CTA) when they do not fit shared memory
    int* work_counter;               // dynamic (c,d) scheduling
    long long n_items;
    int n, I;
    int d_begin, d_end;
    int count_scale;                 // 1 or 2
    unsigned long long cint_mask;
    int bifurcating;                 // argument order of the stored triple (see ordered_triple)
};

__device__ __forceinline__ unsigned long long pack_triple(unsigned long long q1, unsigned long long q2, unsigned long long q3) {
    return (q1 << (2 * QS_TRIPLE_BITS)) | (q2 << QS_TRIPLE_BITS) | q3;
}

// order-preserving double <-> int64 (signed compare / MIN all-reduce order doubles, negatives included)
__host__ __device__ __forceinline__ long long double_to_ordered(double d) {
    long long b;
#ifdef __CUDA_ARCH__
    b = __double_as_longlong(d);
#else
    memcpy(&b, &d, 8);
#endif
    return b ^ ((b >> 63) & 0x7fffffffffffffffLL);
}
__host__ __device__ __forceinline__ double ordered_to_double(long long o) {
    const long long b = o ^ ((o >> 63) & 0x7fffffffffffffffLL);
#ifdef __CUDA_ARCH__
    return __longlong_as_double(b);
#else
    double d; memcpy(&d, &b, 8); return d;
#endif
}
__device__ __forceinline__ int float_to_ordered(float f) { const int b = __float_as_int(f); return b ^ ((b >> 31) & 0x7fffffff); }
__device__ __forceinline__ float ordered_to_float(int o) { return __int_as_float(o ^ ((o >> 31) & 0x7fffffff)); }
constexpr int QS_BOUND_NONE = 0x7f800000;        // +inf

// QIC used ON THE DEVICE ONLY TO SELECT the minimum (QuartetScoreComputer.hpp:135-159 restated);
// reported values are recomputed on the host from the winning triple.
__device__ __noinline__ double dev_log_score(unsigned long long q1, unsigned long long q2, unsigned long long q3) {
    const unsigned long long s = q1 + q2 + q3;
    if (s == 0) return 0.0;
    const bool neg = (q1 < q2) || (q1 < q3);
    double qic;
    if (q1 == s || q2 == s || q3 == s) qic = 1.0;
    else {
        const double inv = 1.0 / (double)s, il3 = 0.91023922662683739361;   // 1/ln 3
        qic = 1.0;
        if (q1) { double p = (double)q1 * inv; qic += p * log(p) * il3; }
        if (q2) { double p = (double)q2 * inv; qic += p * log(p) * il3; }
        if (q3) { double p = (double)q3 * inv; qic += p * log(p) * il3; }
    }
    return neg ? -qic : qic;
}

// fp32 estimate of the same score: |estimate - exact| < 5e-7 (__log2f = MUFU.LG2: absolute error <= 2^-22 on [0.5, 2],
// <= 2 ulp elsewhere, i.e. <= 2.4e-7 per p*log2(p) term; with the fp32 rounding of the rest 2e-7 more: tests/test_score_filter.py).
// q1 is the reference topology's count; symmetric in (q2, q3).  The estimate is exactly 1.0f only for (k,0,0), whose
// exact score is exactly 1.
constexpr float QS_EST_EPS = 1e-6f;            // twice the estimate's error bound
constexpr float QS_FILTER_MARGIN = 2e-6f;      // estimates this close to a run's minimum are evaluated exactly
__device__ __forceinline__ float dev_log_score_f32(unsigned q1, unsigned q2, unsigned q3) {
    const unsigned s = q1 + q2 + q3;
    const float inv = __frcp_rn(fmaxf((float)s, 1.f)), il3 = 0.63092975357145743710f;      // 1/log2(3)
    const float p1 = (float)q1 * inv, p2 = (float)q2 * inv, p3 = (float)q3 * inv;
    float acc = p1 * __log2f(fmaxf(p1, 1e-30f));             // p = 0: 0 * log2(1e-30) = 0
    acc = fmaf(p2, __log2f(fmaxf(p2, 1e-30f)), acc);
    acc = fmaf(p3, __log2f(fmaxf(p3, 1e-30f)), acc);
    float qic = fmaf(acc, il3, 1.f);
    if (q1 < max(q2, q3)) qic = -qic;
    return s == 0 ? 0.f : qic;
}

// (q1,q2,q3) in the order the reference passes them to log_score, from the table slots (c0,c1,c2) of a sorted quartet whose
// reference topology is slot `rslot` (0 or 2): processNodePair (ref, S1S3|S2S4, S1S4|S2S3) for a bifurcating reference
// (:428-432), (u,z|v,w): (ad|bc, ab|cd, ac|bd) in the multifurcating loop (:563-570).  Only the order of the fp64 additions
// depends on it.
__device__ __forceinline__ void ordered_triple(int rslot, int bifurcating, unsigned long long c0, unsigned long long c1, unsigned long long c2,
                                               unsigned long long& q1, unsigned long long& q2, unsigned long long& q3) {
    if (rslot == 0) { q1 = c0; q2 = c1; q3 = c2; }
    else if (bifurcating) { q1 = c2; q2 = c1; q3 = c0; }
    else { q1 = c2; q2 = c0; q3 = c1; }
}

// ---- cold path: a run that may hold a new minimum of its pair ------------------------------------------------------
// Walk the run again, evaluate in fp64 every quartet whose estimate is within the margin of the run's smallest one, and
// publish the best if it beats the pair's exact minimum.  Protocol on (pair_score, pair_best): lower pair_score first
// (atomicMin, strict), then install the triple with a CAS loop that gives up as soon as pair_score shows a better one —
// at the end pair_score[key] is the minimum and pair_best[key] a triple that attains it.
template <typename CINT>
__device__ __noinline__ void scan_candidate_run(const CINT* __restrict__ row, int a0, int a1, int rslot, int bifurcating, int shift, unsigned long long mask,
                                                float run_min, long long key, long long* pair_best, long long* pair_score, int* bound) {
    const long long cur = *reinterpret_cast<volatile long long*>(pair_score + key);
    const double H = cur == QS_I64_NONE ? (double)INFINITY : ordered_to_double(cur);
    if (bound && cur != QS_I64_NONE) atomicMin(bound, float_to_ordered(__double2float_ru(H)));
    const float thr = run_min == 1.f ? run_min : run_min - QS_EST_EPS;
    if (!((double)thr < H)) return;
    double best_q = INFINITY;
    unsigned long long best_t = QS_TRIPLE_NONE, last_t = QS_TRIPLE_NONE;
    for (int x = a0; x < a1; ++x) {
        const unsigned long long c0 = ((unsigned long long)row[(size_t)x * 3 + 0] << shift) & mask, c1 = ((unsigned long long)row[(size_t)x * 3 + 1] << shift) & mask,
                                 c2 = ((unsigned long long)row[(size_t)x * 3 + 2] << shift) & mask;
        unsigned long long q1, q2, q3;
        ordered_triple(rslot, bifurcating, c0, c1, c2, q1, q2, q3);
        const unsigned long long t = pack_triple(q1, q2, q3);
        if (t == last_t) continue;
        last_t = t;
        if ((q1 | q2 | q3) < (1ull << 22)) {                     // (wider counts have no fp32 estimate: always exact)
            const float est = dev_log_score_f32((unsigned)q1, (unsigned)q2, (unsigned)q3);
            if (est > run_min + QS_FILTER_MARGIN) continue;
        }
        const double q = dev_log_score(q1, q2, q3);
        if (q < best_q) { best_q = q; best_t = t; }
    }
    if (!(best_q < H)) return;
    const long long mine = double_to_ordered(best_q);
    if (atomicMin(pair_score + key, mine) <= mine) return;       // somebody holds an equal or better score
    if (bound) atomicMin(bound, float_to_ordered(__double2float_ru(best_q)));
    unsigned long long* pb = reinterpret_cast<unsigned long long*>(pair_best + key);
    unsigned long long old = *reinterpret_cast<volatile unsigned long long*>(pb);
    while (true) {
        if (*reinterpret_cast<volatile long long*>(pair_score + key) < mine) return;     // a better one came in: its owner installs its triple
        const unsigned long long prev = atomicCAS(pb, old, best_t);
        if (prev == old) return;
        old = prev;
    }
}

// ---- accumulators of one CTA ------------------------------------------------------------------------------------------
constexpr int QS_Q4_LEVELS = 32;                                 // accQ covers ancestors of c up to 32 levels below r
constexpr int QS_Q4_SLOTS = QS_Q4_LEVELS * (QS_Q4_LEVELS - 1) / 2;
constexpr int QS_SCAN_STAGES = 2;                                // 48-byte groups in flight per thread (96 bytes: ~70 KB per SM at 768 threads)
constexpr int QS_SCAN_SLOT_BYTES = QS_SCAN_STAGES * 48 + 16;     // (+16: consecutive threads start 28 words apart -> 4-way instead of 16-way bank conflicts)
// layout: sums [3][n_acc] u64 | bound [n_acc] int | pq [n] u16   with n_acc = 2 n + QS_Q4_SLOTS: accR [0,n), accP [n,2n), accQ [2n, 2n+496)
__host__ __device__ __forceinline__ size_t scan_acc_bytes(int n) { return (size_t)(2 * n + QS_Q4_SLOTS) * 28 + (size_t)((n + 7) / 8 * 8) * 2 + 64; }
__host__ __device__ __forceinline__ size_t scan_ring_bytes(int threads) { return (size_t)threads * QS_SCAN_SLOT_BYTES; }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

enum { RUN_DEAD = 0, RUN_R = 1, RUN_P = 2, RUN_Q4 = 3, RUN_G = 4 };

// work item -> (c, d): items are ordered by c descending (the largest pieces first), d ascending inside a c
__host__ __device__ __forceinline__ long long scan_item_count(int dB, int dE) {
    const long long W = dE - dB;
    return W <= 0 ? 0 : W * (W + 1) / 2 + W * (long long)(dB - 3 > 0 ? dB - 3 : 0);
}
__device__ __forceinline__ void scan_item_decode(long long j, int dB, int dE, int& c, int& d) {
    const long long W = dE - dB, tri = W * (W + 1) / 2;
    if (j < tri) {
        long long t = (long long)((sqrt(8.0 * (double)j + 1.0) - 1.0) * 0.5);
        while (t * (t + 1) / 2 > j) --t;
        while ((t + 1) * (t + 2) / 2 <= j) ++t;
        c = dE - 2 - (int)t;
        d = c + 1 + (int)(j - t * (t + 1) / 2);
    } else {
        const long long jj = j - tri;
        c = dB - 2 - (int)(jj / W);
        d = dB + (int)(jj % W);
    }
}

template <typename CINT, int THREADS, bool SMEM_ACC>
__global__ void __launch_bounds__(THREADS, (THREADS <= 256 ? 3 : 1)) qs_scan_rows_kernel(const ScoreArgs a) {
    extern __shared__ __align__(16) unsigned char sm_scan[];
    __shared__ int s_item;
    __shared__ int s_anc[QS_Q4_LEVELS];
    const int tid = threadIdx.x, n = a.n;
    const int n_acc = 2 * n + QS_Q4_SLOTS;
    // accumulators: shared memory, or (large n) this CTA's region of a.scratch
    unsigned char* acc_base = SMEM_ACC ? sm_scan + scan_ring_bytes(THREADS) : reinterpret_cast<unsigned char*>(a.scratch) + (size_t)blockIdx.x * scan_acc_bytes(n);
    unsigned long long* acc_s = reinterpret_cast<unsigned long long*>(acc_base);                 // [3][n_acc]
    int* acc_b = reinterpret_cast<int*>(acc_base + (size_t)n_acc * 24);                           // [n_acc]
    uint16_t* acc_pq = reinterpret_cast<uint16_t*>(acc_base + (size_t)n_acc * 28);                // [n]
    unsigned char* ring = sm_scan + (size_t)tid * QS_SCAN_SLOT_BYTES;
    const int shift = a.count_scale == 2 ? 1 : 0;
    const CINT* table = reinterpret_cast<const CINT*>(a.table);
    const bool bif = a.bifurcating != 0;
    typedef typename std::conditional<(sizeof(CINT) <= 2), uint32_t, unsigned long long>::type sum_t;

    while (true) {
        __syncthreads();
        if (tid == 0) s_item = atomicAdd(a.work_counter, 1);
        __syncthreads();
        const long long item = s_item;
        if (item >= a.n_items) break;
        int c, d;
        scan_item_decode(item, a.d_begin, a.d_end, c, d);
        const int r = a.lca[(size_t)c * n + d], dr = a.idepth[r];
        // zero what this (c,d) can touch: indices < c of accR / accP, all of accQ
        for (int x = tid; x < c; x += THREADS) {
#pragma unroll
            for (int k = 0; k < 3; ++k) { acc_s[(size_t)k * n_acc + x] = 0; acc_s[(size_t)k * n_acc + n + x] = 0; }
            acc_b[x] = QS_BOUND_NONE; acc_b[n + x] = QS_BOUND_NONE;
        }
        for (int x = tid; x < QS_Q4_SLOTS; x += THREADS) {
#pragma unroll
            for (int k = 0; k < 3; ++k) acc_s[(size_t)k * n_acc + 2 * n + x] = 0;
            acc_b[2 * n + x] = QS_BOUND_NONE;
        }
        if (tid < QS_Q4_LEVELS) s_anc[tid] = -1;
        __syncthreads();
        if (tid == 0) {                                        // ancestors of leaf c at depths dr+1 .. dr+32
            for (int x = a.leaf_parent[c]; x >= 0; x = a.inner_parent[x]) {
                const int dx = a.idepth[x];
                if (dx <= dr) break;
                if (dx - dr - 1 < QS_Q4_LEVELS) s_anc[dx - dr - 1] = x;
            }
        }
        const uint64_t cd_base = binom4((uint64_t)d) + binom3((uint64_t)c) - a.rank_base;     // entry index of (a=0, b=0 .. ) of this pair

        // ---- rows: thread k owns rows 1+k and c-1-k (together c entries) ----
        for (int k = tid; 2 * k + 2 <= c; k += THREADS) {
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                const int b = half == 0 ? 1 + k : c - 1 - k;
                if (half == 1 && b <= 1 + k) break;
                const int q = a.lca[(size_t)b * n + c], dq = a.idepth[q];
                const int mqr = min(dq, dr);
                const uint64_t e_begin = cd_base + (uint64_t)b * (b - 1) / 2;
                const CINT* row = table + e_begin * 3;
                // per-run state
                uint32_t ri = a.run_off[b];
                int run_end_a = 0, run_start = 0, type = RUN_DEAD, idx = 0, rslot = 0, pnode = 0;
                sum_t s0 = 0, s1 = 0, s2 = 0;       // per-run sums: <= 65535 x 32768 entries fit 32 bits for 1- and 2-byte counters
                float mn = INFINITY;
                auto flush_run = [&](int a_now) {
                    if (type == RUN_DEAD) return;
                    const sum_t q1 = rslot ? s2 : s0, q3 = rslot ? s0 : s2;
                    int* bptr = nullptr;
                    long long key;
                    if (type == RUN_G) {
                        key = (long long)min(q, pnode) * a.I + max(q, pnode);
                        if (bif) {
                            unsigned long long* ps = a.pair_sums + (size_t)key * 3;
                            if (q1) atomicAdd(ps, (unsigned long long)q1);
                            if (s1) atomicAdd(ps + 1, (unsigned long long)s1);
                            if (q3) atomicAdd(ps + 2, (unsigned long long)q3);
                        }
                    } else {
                        if (bif) {
                            if (q1) atomicAdd(acc_s + idx, (unsigned long long)q1);
                            if (s1) atomicAdd(acc_s + n_acc + idx, (unsigned long long)s1);
                            if (q3) atomicAdd(acc_s + 2 * (size_t)n_acc + idx, (unsigned long long)q3);
                        }
                        bptr = acc_b + idx;
                        if (type == RUN_P) acc_pq[idx - n] = (uint16_t)q;
                    }
                    const float bound = bptr ? ordered_to_float(*reinterpret_cast<volatile int*>(bptr)) : INFINITY;
                    if ((mn == 1.f ? mn : mn - QS_EST_EPS) < bound) {
                        if (type == RUN_R) key = (long long)min(idx, r) * a.I + max(idx, r);
                        else if (type == RUN_P) key = (long long)min(idx - n, q) * a.I + max(idx - n, q);
                        else key = (long long)min(q, pnode) * a.I + max(q, pnode);
                        scan_candidate_run<CINT>(row, run_start, a_now, rslot, a.bifurcating, shift, a.cint_mask, mn, key, a.pair_best, a.pair_score, bptr);
                    }
                };
                auto next_run = [&](int a_now) {
                    flush_run(a_now);
                    run_start = a_now;
                    run_end_a = (int)a.run_end[ri];
                    const uint32_t pd = a.run_pd[ri];
                    ++ri;
                    const int p = (int)(pd & 0xffffu), dp = (int)(pd >> 16);
                    const int S0 = dp + dr, S2 = min(dp, mqr) + dq;
                    s0 = s1 = s2 = 0; mn = INFINITY; pnode = p;
                    if (S0 > S2) {                       // ab|cd: u = deeper(p,q), v = deeper(q,r)
                        rslot = 0;
                        if (dr > dq) { type = RUN_R; idx = dp > dq ? p : q; }
                        else { type = RUN_P; idx = n + p; }            // (dp > dq here: u = p, v = q)
                    } else if (S2 > S0) {                // ad|bc: u = q, v = deeper(p,r)
                        rslot = 2;
                        if (dp > dr) {
                            const int i = dq - dr - 1, jj = dp - dr - 1;
                            if (i < QS_Q4_LEVELS) { type = RUN_Q4; idx = 2 * n + i * (i - 1) / 2 + jj; }
                            else type = RUN_G;
                        } else { type = RUN_R; idx = q; }
                    } else type = RUN_DEAD;
                };
                auto entry = [&](int x, uint32_t r0, uint32_t r1, uint32_t r2) {
                    if (x == run_end_a) next_run(x);
                    const uint32_t c0 = (uint32_t)(((unsigned long long)r0 << shift) & a.cint_mask), c1 = (uint32_t)(((unsigned long long)r1 << shift) & a.cint_mask),
                                   c2 = (uint32_t)(((unsigned long long)r2 << shift) & a.cint_mask);
                    s0 += c0; s1 += c1; s2 += c2;
                    mn = fminf(mn, dev_log_score_f32(rslot ? c2 : c0, c1, rslot ? c0 : c2));
                };
                if (sizeof(CINT) == 2) {
                    // 48-byte groups (8 entries) of the row, 16-byte aligned because the table is: group g = entries [8g, 8g+8)
                    const unsigned char* tbytes = reinterpret_cast<const unsigned char*>(table);
                    const long long g_first = (long long)(e_begin >> 3), g_last = (long long)((e_begin + b - 1) >> 3);
                    auto issue = [&](long long g, uint32_t ring_off) {       // (an empty group is committed past the row's end: the group count stays uniform)
                        if (g <= g_last) {
                            const unsigned char* src = tbytes + (size_t)g * 48;
                            unsigned char* dst = ring + ring_off;
                            cp_async16(dst, src); cp_async16(dst + 16, src + 16); cp_async16(dst + 32, src + 32);
                        }
                        cp_async_commit();
                    };
#pragma unroll
                    for (int s = 0; s < QS_SCAN_STAGES; ++s) issue(g_first + s, (uint32_t)s * 48u);
                    const long long P_first = g_first * 4, P_begin = (long long)(e_begin >> 1), P_end = (long long)((e_begin + b - 1) >> 1);
                    uint32_t off = 0;                                 // byte offset of the current pair inside the ring
                    for (long long P = P_first; P <= P_end; ++P) {
                        if ((P & 3) == 0) cp_async_wait<QS_SCAN_STAGES - 1>();
                        if (P >= P_begin) {
                            const uint32_t w0 = *reinterpret_cast<const uint32_t*>(ring + off), w1 = *reinterpret_cast<const uint32_t*>(ring + off + 4),
                                           w2 = *reinterpret_cast<const uint32_t*>(ring + off + 8);
                            const int x0 = (int)(2 * P - (long long)e_begin);
                            if (x0 >= 0) entry(x0, w0 & 0xffffu, w0 >> 16, w1 & 0xffffu);
                            if (x0 + 1 < b) entry(x0 + 1, w1 >> 16, w2 & 0xffffu, w2 >> 16);
                        }
                        off += 12;
                        if ((P & 3) == 3) {                          // the group is consumed: refill its slot with the group QS_SCAN_STAGES ahead
                            issue((P >> 2) + QS_SCAN_STAGES, off - 48u);
                            if (off == QS_SCAN_STAGES * 48) off = 0;
                        }
                    }
                    cp_async_wait<0>();
                } else {
                    for (int x = 0; x < b; ++x) entry(x, (uint32_t)row[(size_t)x * 3], (uint32_t)row[(size_t)x * 3 + 1], (uint32_t)row[(size_t)x * 3 + 2]);   // (counts <= m < 2^31)
                }
                flush_run(b);
            }
        }
        __syncthreads();
        // ---- flush the CTA's accumulators: one global atomic per touched key ----
        if (bif) {
            for (int x = tid; x < c; x += THREADS) {
                const unsigned long long r1 = acc_s[x], r2 = acc_s[n_acc + x], r3 = acc_s[2 * (size_t)n_acc + x];
                if (r1 | r2 | r3) {
                    unsigned long long* ps = a.pair_sums + ((size_t)min(x, r) * a.I + max(x, r)) * 3;
                    if (r1) atomicAdd(ps, r1);
                    if (r2) atomicAdd(ps + 1, r2);
                    if (r3) atomicAdd(ps + 2, r3);
                }
                const unsigned long long p1 = acc_s[n + x], p2 = acc_s[n_acc + n + x], p3 = acc_s[2 * (size_t)n_acc + n + x];
                if (p1 | p2 | p3) {
                    const int qq = acc_pq[x];
                    unsigned long long* ps = a.pair_sums + ((size_t)min(x, qq) * a.I + max(x, qq)) * 3;
                    if (p1) atomicAdd(ps, p1);
                    if (p2) atomicAdd(ps + 1, p2);
                    if (p3) atomicAdd(ps + 2, p3);
                }
            }
            for (int x = tid; x < QS_Q4_SLOTS; x += THREADS) {
                const unsigned long long t1 = acc_s[2 * n + x], t2 = acc_s[n_acc + 2 * n + x], t3 = acc_s[2 * (size_t)n_acc + 2 * n + x];
                if (t1 | t2 | t3) {
                    int i = (int)((1.f + sqrtf(1.f + 8.f * (float)x)) * 0.5f);
                    while (i * (i - 1) / 2 > x) --i;
                    while ((i + 1) * i / 2 <= x) ++i;
                    const int jj = x - i * (i - 1) / 2;
                    const int u = s_anc[i], v = s_anc[jj];
                    unsigned long long* ps = a.pair_sums + ((size_t)min(u, v) * a.I + max(u, v)) * 3;
                    if (t1) atomicAdd(ps, t1);
                    if (t2) atomicAdd(ps + 1, t2);
                    if (t3) atomicAdd(ps + 2, t3);
                }
            }
        }
    }
}

__global__ void qs_fill_i64_kernel(long long* __restrict__ p, size_t n, long long v) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

// multi-GPU: after the MIN all-reduce of pair_score, only the shards that hold the winning score keep their triple
__global__ void qs_select_winners_kernel(const long long* __restrict__ score_local, const long long* __restrict__ score_reduced, long long* __restrict__ best, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        if (score_local[i] != score_reduced[i]) best[i] = QS_I64_NONE;
}

// ---- per-edge reduction of the pair aggregates (QuartetScoreComputer.hpp:448-454, 472-489) -------------------------------
// LQ-IC[e] = min over pairs whose path contains e of the pair's min QIC; EQP-IC[e] = min over those pairs of qpic(pair);
// QP-IC[e] = qpic of the pair of e's two end nodes.  Pass 1 takes the minima (device fp64 scores, order-preserving int64),
// pass 2 finds WHO attains them (smallest triple / smallest pair index among equals), and the host evaluates log_score with
// its libm for the selected triples and sums only.
struct EdgeArgs {
    const unsigned long long* pair_sums;
    const long long* pair_best;
    const long long* pair_score;
    const int32_t* inner_node;      // [I] node id
    const int32_t* node_parent;     // [n_nodes]
    const int32_t* node_depth;      // [n_nodes]
    const int32_t* node_edge;       // [n_nodes] edge above the node
    const int32_t* node_inner;      // [n_nodes] inner index or -1
    long long* edge_lq;             // [E] min score           (pass 1)
    long long* edge_eqp;            // [E] min qp score
    long long* edge_lq_arg;         // [E] triple              (pass 2)
    long long* edge_eqp_arg;        // [E] pair key
    unsigned long long* out;        // [E][7]: lq triple, eqp sums[3], qp sums[3] (~0 in [4] = no adjacent pair)
    int I, E, n_nodes, bifurcating, exact_qp;
};

__device__ __forceinline__ double pair_qp_score(const EdgeArgs& a, long long key) {
    unsigned long long p1 = a.pair_sums[key * 3], p2 = a.pair_sums[key * 3 + 1], p3 = a.pair_sums[key * 3 + 2];
    if (!a.exact_qp) { p1 &= 0xffffffffull; p2 &= 0xffffffffull; p3 &= 0xffffffffull; }      // `unsigned p1,p2,p3` (QuartetScoreComputer.hpp:382)
    return dev_log_score(p1, p2, p3);
}

template <int PASS>
__global__ void qs_edge_reduce_kernel(const EdgeArgs a) {
    const long long total = (long long)a.I * a.I;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
        const int iu = (int)(k / a.I), iv = (int)(k % a.I);
        if (iu >= iv) continue;
        const long long sc = a.pair_score[k];
        const bool has_lq = sc != QS_I64_NONE;
        long long qp = 0;
        if (a.bifurcating) qp = double_to_ordered(pair_qp_score(a, k));
        if (!has_lq && !a.bifurcating) continue;
        int u = a.inner_node[iu], v = a.inner_node[iv];
        while (u != v) {
            int e;
            if (a.node_depth[u] >= a.node_depth[v]) { e = a.node_edge[u]; u = a.node_parent[u]; }
            else { e = a.node_edge[v]; v = a.node_parent[v]; }
            if (PASS == 1) {
                if (has_lq) atomicMin(a.edge_lq + e, sc);
                if (a.bifurcating) atomicMin(a.edge_eqp + e, qp);
            } else {
                if (has_lq && a.edge_lq[e] == sc) atomicMin(a.edge_lq_arg + e, a.pair_best[k]);
                if (a.bifurcating && a.edge_eqp[e] == qp) atomicMin(a.edge_eqp_arg + e, k);
            }
        }
    }
}

__global__ void qs_edge_gather_kernel(const EdgeArgs a) {
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < a.n_nodes; v += gridDim.x * blockDim.x) {
        if (v == 0) continue;
        const int e = a.node_edge[v];
        unsigned long long* o = a.out + (size_t)e * 7;
        o[0] = (unsigned long long)a.edge_lq_arg[e];
        for (int k = 1; k < 7; ++k) o[k] = ~0ull;
        if (!a.bifurcating) continue;
        const long long ke = a.edge_eqp_arg[e];
        if (ke != QS_I64_NONE) for (int k = 0; k < 3; ++k) o[1 + k] = a.pair_sums[ke * 3 + k];
        const int iv = a.node_inner[v], iu = a.node_inner[a.node_parent[v]];      // QP-IC: the edge's own pair (:475-481)
        if (iv >= 0 && iu >= 0) {
            const long long kq = (long long)min(iu, iv) * a.I + max(iu, iv);
            for (int k = 0; k < 3; ++k) o[4 + k] = a.pair_sums[kq * 3 + k];
        }
    }
}

}  // namespace qs
