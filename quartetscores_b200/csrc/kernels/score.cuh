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
// qs_scan_kernel (round 2; replaces the per-thread run scan of round 1, which read 48 bytes at a time from rows ~270 KB
// apart, spent ~290 lane-instructions per quartet and issued 0.4 global atomics per quartet — 5 % of the HBM roofline).
//   * A CTA owns one (c,d): the C(c,2) entries (b,a) of that pair are ONE contiguous piece of the table (rank =
//     C(d,4)+C(c,3)+C(b,2)+a, src/quartet_lookup_table.hpp:141-168).  It is streamed through a shared-memory ring by TMA
//     bulk copies (cp.async.bulk + full/empty mbarriers, 12-24 KB per stage), so control flow is uniform and every byte of the
//     table is read exactly once.
//   * The pair key (u,v) of a quartet has one end implied by (c,d), so the CTA accumulates in shared memory, indexed by ONE node:
//       key (x, r)   r = lca(c,d)                                   -> accR[x]   x = lca(a,b) or lca(b,c)
//       key (p, q)   p = lca(a,b) below q = lca(b,c)                -> accP[p]   (q is the ancestor of p that joins c)
//       key (q, p)   p above q, both ancestors of c deeper than r   -> accQ[depth(q)-dr-1][depth(p)-dr-1]  (as many levels as the
//                                                                      reference tree is deep, up to 128; deeper ones go
//                                                                      straight to global memory)
//     Inner nodes are numbered by their first gap in the planar leaf order (qscuda.cu build_reference), so every index the
//     CTA of (c,d) touches is < c and only c entries are zeroed and flushed: one global atomic per touched key and CTA
//     (~n^3 in total) instead of one per run of equal keys (~0.4 C(n,4)).
//   * A thread takes 4 CONSECUTIVE entries of a staged chunk: their LCA / depth lookups are independent and overlap, and the
//     sums go to shared memory as fire-and-forget 32-bit REDs (with a carry word only when the host cannot rule out an
//     overflow inside one (c,d)).  (A MATCH.ANY + REDUX warp aggregation was measured first: 47 % of all instructions and
//     the longest stalls, profiles/r02_c_scan_n500_ncu_full.txt.)
//   * LQ-IC selection without an fp64 log per quartet: every quartet gets an fp32 estimate (error < 5e-7) that is compared
//     with the slot's bound, an fp32 upper bound of the pair's current exact minimum (fetched from pair_score when the CTA
//     starts, lowered as minima are found); only a quartet that may beat it is evaluated in fp64 (scan_candidate).
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
    const uint32_t* lcapd;           // [n][n] inner index of lca(leaf x, leaf y) | its depth << 16 (one load instead of two dependent ones)
    const uint16_t* idepth;          // [I] depth of an inner node (by inner index)
    const int32_t* inner_parent;     // [I] inner index of the parent inner node, -1 for the root
    const int32_t* leaf_parent;      // [n] inner index of the parent of leaf x
    const int32_t* inner_gap;        // [I] first gap of the inner node = a leaf below its first child (n + k for single-child nodes)
    unsigned long long* pair_sums;   // [I*I][3]  key = min(u,v) * I + max(u,v)
    long long* pair_best;            // [I*I] packed triple of the min-QIC quartet, QS_I64_NONE = none
    long long* pair_score;           // [I*I] order-preserving int64 image of the EXACT (device fp64) score of pair_best
    unsigned long long* scratch;     // accumulators in global memory (one region per CTA) when they do not fit shared memory
    int* work_counter;               // dynamic item scheduling
    const int4* items;               // work items (c, d0, d1, -): the pairs (c,d), d0 <= d < d1, share r = lca(c,d); largest first
    long long n_items;
    int n, I;
    int d_begin, d_end;
    int count_scale;                 // 1 or 2
    unsigned long long cint_mask;
    int bifurcating;                 // argument order of the stored triple (see ordered_triple)
    int q4_levels;                   // levels of accQ (ancestors of c below r it covers), chosen by the host from the reference tree's depth
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
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float dev_log_score_f32(unsigned q1, unsigned q2, unsigned q3) {
    const unsigned s = q1 + q2 + q3;
    // MUFU.RCP / MUFU.LG2 without the denormal fix-ups of __frcp_rn / __log2f: the arguments are >= 1e-30 (rcp: 1 ulp, i.e. one
    // more ~1e-7 on the estimate; still inside QS_EST_EPS, see tests/test_score_filter.py)
    const float inv = rcp_approx(fmaxf((float)s, 1.f)), il3 = 0.63092975357145743710f;      // 1/log2(3)
    const float p1 = (float)q1 * inv, p2 = (float)q2 * inv, p3 = (float)q3 * inv;
    float acc = p1 * lg2_approx(fmaxf(p1, 1e-30f));          // p = 0: 0 * log2(1e-30) = 0
    acc = fmaf(p2, lg2_approx(fmaxf(p2, 1e-30f)), acc);
    acc = fmaf(p3, lg2_approx(fmaxf(p3, 1e-30f)), acc);
    float qic = fmaf(acc, il3, 1.f);
    if (q1 < max(q2, q3)) qic = -qic;
    return s == 0 ? 0.f : qic;
}
// Triples with at most one non-zero count score exactly +1 / -1 / 0 (x * rcp(x) is not exactly 1 in fp32, so the estimate
// would say 0.99999994): the estimate is replaced by the exact value and compared without the error margin.  Unanimous
// quartets are the bulk of a well-supported tree; with the margin every one of them would tie with its pair's minimum of 1
// and take the exact path (measured: 11 % of all quartets, profiles/r02_d_scan_n500_ncu_full.txt).
__device__ __forceinline__ bool score_known_exactly(unsigned q1, unsigned q2, unsigned q3, float& est) {
    const bool z1 = q1 == 0, z2 = q2 == 0, z3 = q3 == 0;
    if ((int)z1 + (int)z2 + (int)z3 >= 2) { est = (z2 && z3) ? (z1 ? 0.f : 1.f) : -1.f; return true; }
    return false;
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

// ---- cold path: one quartet that may be the new minimum of its pair ---------------------------------------------------
// Evaluate it in fp64 and publish it if it beats the pair's exact minimum.  Protocol on (pair_score, pair_best): lower
// pair_score first (atomicMin, strict), then install the triple with a CAS loop that gives up as soon as pair_score shows a
// better one — at the end pair_score[key] is the minimum and pair_best[key] a triple that attains it.  `bound` is the CTA's
// shared-memory copy for the slot (a conservative fp32 bound of the minimum): it only filters, races on it are benign.
__device__ __noinline__ void scan_candidate(unsigned long long q1, unsigned long long q2, unsigned long long q3, long long key, long long* pair_best,
                                            long long* pair_score, int* bound) {
    const unsigned long long t = pack_triple(q1, q2, q3);
    const long long cur = *reinterpret_cast<volatile long long*>(pair_score + key);
    if (bound && cur != QS_I64_NONE) atomicMin(bound, float_to_ordered(__double2float_ru(ordered_to_double(cur))));
    if (*reinterpret_cast<volatile unsigned long long*>(pair_best + key) == t) return;           // the very counts that hold the pair's minimum
    const long long mine = double_to_ordered(dev_log_score(q1, q2, q3));
    if (!(mine < cur)) return;
    if (bound) atomicMin(bound, float_to_ordered(__double2float_ru(ordered_to_double(mine))));
    if (atomicMin(pair_score + key, mine) <= mine) return;       // somebody holds an equal or better score
    unsigned long long* pb = reinterpret_cast<unsigned long long*>(pair_best + key);
    unsigned long long old = *reinterpret_cast<volatile unsigned long long*>(pb);
    while (true) {
        if (*reinterpret_cast<volatile long long*>(pair_score + key) < mine) return;         // a better one came in: its owner installs its triple
        const unsigned long long prev = atomicCAS(pb, old, t);
        if (prev == old) return;
        old = prev;
    }
}

// ---- accumulators of one CTA ------------------------------------------------------------------------------------------
constexpr int QS_Q4_MAX_LEVELS = 128;                            // accQ covers ancestors of c up to q4_levels <= 128 levels below r
__host__ __device__ __forceinline__ int q4_slots(int levels) { return levels * (levels - 1) / 2; }
constexpr int QS_SCAN_STEPS = 4;                                 // entries per thread and staged chunk
__host__ __device__ constexpr int scan_stages(int threads) { return threads >= 900 ? 2 : 4; }
__host__ __device__ constexpr uint32_t scan_chunk_bytes(int threads) { return (uint32_t)threads * QS_SCAN_STEPS * 6u; }
// accumulator slots: accR [0,n), accP [n,2n), accQ [2n, 2n + q4_slots).  Per slot: lo[3] u32 | (carry builds: hi[3] u32) | bound int | tau int;
// then pq [n] u16 and anc [levels] int
__host__ __device__ __forceinline__ size_t scan_acc_bytes(int n, int levels, bool carry) {
    return (size_t)(2 * n + q4_slots(levels)) * (carry ? 32 : 20) + (size_t)((n + 7) / 8 * 8) * 2 + (size_t)levels * 4 + 64;
}
// Candidates for a pair's minimum are not evaluated where they are found: a warp that stops for the fp64 evaluation and its global
// atomics (~5 dependent round trips) holds up its whole CTA at the item's barrier — at n = 100, where an item is a few thousand
// entries, that was most of the kernel's time (profiles/r02_d_scan_cfg2_*).  Each warp queues them in shared memory and evaluates a
// batch with all lanes at once; the slot's bound is lowered at once with the estimate's upper end, which is all the filter needs.
struct ScanCand { uint32_t c0, c1, c2, uv; int key; };            // counts, u | v << 16, slot * 2 + (reference topology is slot 2)
__host__ __device__ constexpr int scan_queue_cap(int threads) { return threads >= 900 ? 16 : 32; }
__host__ __device__ constexpr size_t scan_queue_bytes(int threads) { return (size_t)(threads / 32) * ((size_t)scan_queue_cap(threads) * sizeof(ScanCand) + 16); }
__host__ __device__ __forceinline__ size_t scan_ring_bytes(int threads, int cint_bytes) {       // barriers | staging ring | tau table | candidate queues
    return 128 + (cint_bytes == 2 ? (size_t)scan_stages(threads) * scan_chunk_bytes(threads) : 0) + (1024 + 8) * 4 + scan_queue_bytes(threads);
}

__device__ __forceinline__ int bound_from_score(long long sc) {
    return sc == QS_I64_NONE ? QS_BOUND_NONE : float_to_ordered(__double2float_ru(ordered_to_double(sc)));
}

// ---- cheap exclusion before the estimate ----------------------------------------------------------------------------------
// For q1 >= max(q2,q3) the score is smallest when the rest is split evenly: score >= g(p1), p1 = q1/s,
//   g(p) = 1 + (p ln p + (1-p) ln((1-p)/2)) / ln 3,   increasing on [1/3, 1].
// Each slot keeps tau = the smallest p1 whose g(p1) reaches the slot's bound (rounded up): a quartet with q1 >= max(q2,q3) and
// q1 >= tau * s cannot beat the pair's minimum and needs no estimate — most quartets of a pair whose minimum lies below them.
constexpr int QS_TAU_STEPS = 1024;      // tau is tabulated once per CTA at B = j / 1024 and looked up at the next point above the bound (tau grows with B)
__device__ __forceinline__ float qs_tau_lookup(const float* tab, float B) {
    if (!(B > 0.f)) return 0.f;
    if (!(B < 1.f)) return B > 1.f ? INFINITY : 1.f;
    return tab[min(QS_TAU_STEPS, (int)(B * (float)QS_TAU_STEPS) + 1)];
}
__device__ __noinline__ float qs_tau_of_bound(float B) {
    if (!(B > 0.f)) return 0.f;                                  // every such quartet scores >= 0 >= B
    if (!(B < 1.f)) return B > 1.f ? INFINITY : 1.f;             // no minimum yet: never excluded; minimum 1: only unanimous quartets are
    float lo = 1.f / 3.f, hi = 1.f;
    for (int it = 0; it < 22; ++it) {
        const float mid = 0.5f * (lo + hi), rest = 1.f - mid;
        const float gm = 1.f + (mid * lg2_approx(mid) + rest * lg2_approx(fmaxf(rest * 0.5f, 1e-30f))) * 0.63092975357145743710f;
        if (gm >= B + 4e-6f) hi = mid; else lo = mid;            // (+4e-6: fp32 evaluation error of g, on the safe side)
    }
    return fminf(1.f, hi * (1.f + 2e-6f) + 1e-6f);
}

// ---- cold path of the scan kernel: a quartet that may hold the new minimum of its pair, or one whose key has no slot ------
// (u,v) = the key's inner nodes; slot < 0: the key is deeper than accQ reaches and its sums go to global memory directly.
__device__ __noinline__ void scan_cold(unsigned long long* pair_sums, long long* pair_best, long long* pair_score, int I, int bifurcating, int rslot, uint32_t c0,
                                       uint32_t c1, uint32_t c2, int u, int v, int* bound, int* tau, const float* tau_tab, bool add_sums) {
    const long long key = (long long)min(u, v) * I + max(u, v);
    unsigned long long q1, q2, q3;
    if (add_sums) {                                              // (reference topology, crossing, other): QuartetScoreComputer.hpp:429-431
        unsigned long long* ps = pair_sums + (size_t)key * 3;
        const uint32_t a1 = rslot ? c2 : c0, a3 = rslot ? c0 : c2;
        if (a1) atomicAdd(ps, (unsigned long long)a1);
        if (c1) atomicAdd(ps + 1, (unsigned long long)c1);
        if (a3) atomicAdd(ps + 2, (unsigned long long)a3);
    }
    ordered_triple(rslot, bifurcating, c0, c1, c2, q1, q2, q3);
    const int before = bound ? *reinterpret_cast<volatile int*>(bound) : 0;
    scan_candidate(q1, q2, q3, key, pair_best, pair_score, bound);
    if (bound && *reinterpret_cast<volatile int*>(bound) != before)
        atomicMin(tau, float_to_ordered(qs_tau_lookup(tau_tab, ordered_to_float(*reinterpret_cast<volatile int*>(bound)))));
}

// CARRY: the 32-bit shared-memory sums may overflow inside one item (max count x C(c,2) x run length >= 2^32, decided by the
// host): every add then returns the old value and feeds a carry word.  Without it the adds are fire-and-forget REDs.
template <typename CINT, int THREADS, bool SMEM_ACC, bool CARRY>
__global__ void __launch_bounds__(THREADS + 32, (THREADS <= 512 ? 2 : 1)) qs_scan_kernel(const ScoreArgs a) {
    // THREADS consumer threads + one producer warp that only copies chunks in (TMA): the refill of a stage never waits for a
    // consumer to get around to it (with a consumer thread issuing the copies 20 % of all instructions were spins on the full barrier)
    extern __shared__ __align__(128) unsigned char sm_scan[];
    __shared__ int s_item, s_levels;
    constexpr bool RING = sizeof(CINT) == 2;
    constexpr int K = QS_SCAN_STEPS, STAGES = scan_stages(THREADS), CHUNK = THREADS * K, WARPS = THREADS / 32;
    constexpr uint32_t CHUNK_BYTES = scan_chunk_bytes(THREADS);
    static_assert(K == 4, "the entry loop below is written for 4 consecutive entries (24 bytes) per thread");
    const int tid = threadIdx.x, n = a.n, LV = a.q4_levels;
    const int n_acc = 2 * n + q4_slots(LV);
    uint64_t* full = reinterpret_cast<uint64_t*>(sm_scan);            // [STAGES] chunk landed (TMA transaction barrier)
    uint64_t* empty = full + 4;                                       // [STAGES] every warp is done with the chunk
    unsigned char* ring = sm_scan + 128;
    // accumulators: shared memory, or (large n) this CTA's region of a.scratch
    unsigned char* acc_base = SMEM_ACC ? sm_scan + scan_ring_bytes(THREADS, (int)sizeof(CINT)) : reinterpret_cast<unsigned char*>(a.scratch) + (size_t)blockIdx.x * scan_acc_bytes(n, LV, CARRY);
    uint32_t* acc_lo = reinterpret_cast<uint32_t*>(acc_base);                                     // [3][n_acc] low words of the sums
    uint32_t* acc_hi = acc_lo + (CARRY ? 3 * (size_t)n_acc : 0);                                  // [3][n_acc] carries (carry builds only)
    int* acc_b = reinterpret_cast<int*>(acc_hi + 3 * (size_t)n_acc);                              // [n_acc] fp32 bound of the key's minimum (ordered int)
    int* acc_tau = acc_b + n_acc;                                                                 // [n_acc] exclusion threshold of the bound (ordered int of a float >= 0)
    uint16_t* acc_pq = reinterpret_cast<uint16_t*>(acc_tau + n_acc);                              // [n] q of the accP key (p, q)
    int* s_anc = reinterpret_cast<int*>(acc_pq + (n + 7) / 8 * 8);                                // [LV] ancestors of leaf c below r
    float* s_tau = reinterpret_cast<float*>(sm_scan + 128 + (RING ? (size_t)STAGES * CHUNK_BYTES : 0));   // [QS_TAU_STEPS + 1] (always shared memory)
    constexpr int QCAP = scan_queue_cap(THREADS);
    unsigned char* wq_base = reinterpret_cast<unsigned char*>(s_tau) + (1024 + 8) * 4 + (size_t)(threadIdx.x >> 5 < WARPS ? threadIdx.x >> 5 : 0) * (QCAP * sizeof(ScanCand) + 16);
    int* wq_cnt = reinterpret_cast<int*>(wq_base);                                                // this warp's queue of candidates: count,
    ScanCand* wq = reinterpret_cast<ScanCand*>(wq_base + 16);                                     //   entries
    if ((tid & 31) == 0 && tid < THREADS) *wq_cnt = 0;
    const int shift = a.count_scale == 2 ? 1 : 0;
    const uint32_t mask32 = (uint32_t)a.cint_mask;                    // (counts are <= m < 2^31 whatever CINT is; the scaled value is masked to CINT)
    const CINT* table = reinterpret_cast<const CINT*>(a.table);
    const unsigned char* tbytes = reinterpret_cast<const unsigned char*>(a.table);
    const bool bif = a.bifurcating != 0;
    const bool producer = tid >= THREADS;                              // warp THREADS / 32
    uint32_t full_phase = 0, empty_phase = 0, stage_used = 0;
    if (RING && tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], WARPS); }
        fence_mbar_init();
    }
    auto pair_key = [&](int u, int v) -> long long { return (long long)min(u, v) * a.I + max(u, v); };
    auto add_sum = [&](int k, int slot, uint32_t v) {
        if (v == 0) return;
        if (CARRY) { const uint32_t old = atomicAdd(acc_lo + k * n_acc + slot, v); if (old + v < old) atomicAdd(acc_hi + k * n_acc + slot, 1u); }
        else atomicAdd(acc_lo + k * n_acc + slot, v);
    };
    auto tri_row = [](int x) -> int {                                 // row i of the lower-triangular index x = i(i-1)/2 + j, j < i
        int i = (int)(0.5f + sqrtf(2.f * (float)x + 0.25f));
        while (i * (i - 1) / 2 > x) --i;
        while ((i + 1) * i / 2 <= x) ++i;
        return i;
    };
    auto set_bound = [&](int slot, long long score) {
        const int bb = bound_from_score(score);
        acc_b[slot] = bb;
        acc_tau[slot] = float_to_ordered(qs_tau_lookup(s_tau, ordered_to_float(bb)));
    };
    auto drain = [&]() {                                          // consumer warps, all lanes: evaluate the queued candidates, 32 at a time
        __syncwarp();
        const int cnt = min(*reinterpret_cast<volatile int*>(wq_cnt), QCAP);
        for (int base = 0; base < cnt; base += 32) {
            const int e = base + (tid & 31);
            if (e < cnt) {
                const ScanCand cd = wq[e];
                const int sl = cd.key >> 1;
                scan_cold(a.pair_sums, a.pair_best, a.pair_score, a.I, a.bifurcating, (cd.key & 1) ? 2 : 0, cd.c0, cd.c1, cd.c2, (int)(cd.uv & 0xffffu), (int)(cd.uv >> 16), acc_b + sl, acc_tau + sl, s_tau, false);
            }
        }
        __syncwarp();
        if ((tid & 31) == 0) *wq_cnt = 0;
        __syncwarp();
    };
    for (int j = tid; j <= QS_TAU_STEPS; j += THREADS + 32) s_tau[j] = qs_tau_of_bound((float)j / (float)QS_TAU_STEPS);

    while (true) {
        __syncthreads();
        if (tid == 0) s_item = atomicAdd(a.work_counter, 1);
        __syncthreads();
        const long long item = s_item;
        if (item >= a.n_items) break;
        // item = (c; d0 <= d < d1) with the same r = lca(c,d) for every d: the keys of its quartets do not depend on d, so the
        // accumulators are set up and flushed once for the whole run of d (on average n / depth pairs instead of one)
        const int4 it = a.items[item];
        const int c = it.x, d0 = it.y, d1 = it.z;
        const uint32_t rd = a.lcapd[(uint32_t)c * (uint32_t)n + d0];
        const int r = (int)(rd & 0xffffu), dr = (int)(rd >> 16);
        const int L = c * (c - 1) / 2;                           // entries (b,a), a < b < c, of one (c,d): consecutive in the table from E0(d)
        auto geom = [&](int d, uint64_t& E0, uint64_t& G0, int& n_chunks) {       // ... read in 48-byte groups [G0, G1), CHUNK/8 groups per chunk
            E0 = binom4((uint64_t)d) + binom3((uint64_t)c) - a.rank_base;
            G0 = E0 >> 3;
            const uint64_t G1 = (E0 + (uint64_t)L + 7) >> 3;
            n_chunks = (int)((G1 - G0 + CHUNK / 8 - 1) / (CHUNK / 8));
        };
        // producer warp (one lane): copy the item's chunks in, in the order the consumers take them (d ascending, chunks ascending);
        // a stage is refilled as soon as every consumer warp has arrived on its empty barrier
        int pr_d = d0, pr_k = 0, pr_g = 0;                         // producer's cursor: next d, next chunk of it, chunks issued in this item
        auto issue_next = [&]() {
            uint64_t E0, G0; int nch;
            geom(pr_d, E0, G0, nch);
            const uint64_t G1 = (E0 + (uint64_t)L + 7) >> 3, g = G0 + (uint64_t)pr_k * (CHUNK / 8);
            const uint32_t bytes = (uint32_t)min((uint64_t)(CHUNK / 8), G1 - g) * 48u;
            const int st = pr_g % STAGES;
            if ((stage_used >> st) & 1u) { mbar_wait_relaxed(&empty[st], (empty_phase >> st) & 1u); empty_phase ^= 1u << st; }      // its previous chunk is consumed
            stage_used |= 1u << st;
            mbar_expect_tx(&full[st], bytes);
            bulk_g2s(ring + (size_t)st * CHUNK_BYTES, tbytes + g * 48, bytes, &full[st]);
            ++pr_g;
            if (++pr_k == nch) { pr_k = 0; ++pr_d; }
        };
        if (RING && tid == THREADS) for (int k = 0; k < STAGES && pr_d < d1; ++k) issue_next();     // in flight while the accumulators are prepared
        if (tid == 0) {
            int levels = 0;                                                            // ancestors of leaf c at depths dr+1 .. dr+LV
            for (int x = a.leaf_parent[c]; x >= 0; x = a.inner_parent[x]) {
                const int lv = (int)a.idepth[x] - dr - 1;
                if (lv < 0) break;
                if (lv < LV) { s_anc[lv] = x; levels = max(levels, lv + 1); }
            }
            s_levels = levels;                                                         // (unary nodes never are an LCA: their levels are never addressed)
        }
        __syncthreads();
        const int nq4 = q4_slots(s_levels);
        // zero what this (c, .) can touch (indices < c of accR / accP, the present levels of accQ) and fetch the keys' current minima
        for (int x = tid; x < c && !producer; x += THREADS) {
#pragma unroll
            for (int k = 0; k < 3; ++k) { acc_lo[k * n_acc + x] = 0; acc_lo[k * n_acc + n + x] = 0; if (CARRY) { acc_hi[k * n_acc + x] = 0; acc_hi[k * n_acc + n + x] = 0; } }
            set_bound(x, x != r ? a.pair_score[pair_key(x, r)] : QS_I64_NONE);
            long long sp = QS_I64_NONE;
            const int g = a.inner_gap[x];
            if (g < c) { const int qq = (int)(a.lcapd[(uint32_t)g * (uint32_t)n + c] & 0xffffu); if (qq != x) sp = a.pair_score[pair_key(x, qq)]; }
            set_bound(n + x, sp);
        }
        for (int x = tid; x < nq4 && !producer; x += THREADS) {
#pragma unroll
            for (int k = 0; k < 3; ++k) { acc_lo[k * n_acc + 2 * n + x] = 0; if (CARRY) acc_hi[k * n_acc + 2 * n + x] = 0; }
            const int i = tri_row(x);
            set_bound(2 * n + x, a.pair_score[pair_key(s_anc[i], s_anc[x - i * (i - 1) / 2])]);
        }
        __syncthreads();

        if (producer) {
            if (RING && tid == THREADS) while (pr_d < d1) issue_next();
        } else {
        int gk = 0;                                              // chunks consumed so far in this item (ring position)
        for (int d = d0; d < d1; ++d) {
            uint64_t E0, G0; int n_chunks;
            geom(d, E0, G0, n_chunks);
            const int off0 = (int)((long long)(G0 << 3) - (long long)E0) + tid * K;      // my first entry of chunk 0 relative to E0 (> -8)
            for (int k = 0; k < n_chunks; ++k, ++gk) {
                const int stage = gk % STAGES;
                if (RING) { mbar_wait(&full[stage], (full_phase >> stage) & 1u); full_phase ^= 1u << stage; }
                const int xa = off0 + k * CHUNK;                  // entries xa .. xa+3 of this (c,d); valid ones are in [0, L)
                if (xa + K > 0 && xa < L) {
                    // row of the first valid entry: C(b,2) <= x < C(b+1,2)
                    const int xf = max(xa, 0);
                    int b = (int)(0.5f + sqrtf(2.f * (float)xf + 0.25f));
                    while (b * (b - 1) / 2 > xf) --b;
                    while ((b + 1) * b / 2 <= xf) ++b;
                    int aa = xf - b * (b - 1) / 2;
                    const uint32_t* lrow = a.lcapd + (uint32_t)b * (uint32_t)n;
                    uint32_t w[6];
                    if (RING) {                                   // 24 contiguous bytes, 8-byte aligned: three 64-bit loads, conflict-free
                        const uint2* src = reinterpret_cast<const uint2*>(ring + (size_t)stage * CHUNK_BYTES + (size_t)tid * (K * 6));
#pragma unroll
                        for (int i = 0; i < 3; ++i) { const uint2 v = src[i]; w[2 * i] = v.x; w[2 * i + 1] = v.y; }
                    }
                    // per entry: the two lca words (eight independent loads), then the slot; predicated, not branched
                    uint32_t pd[K], qd[K];
                    int slot[K];
#pragma unroll
                    for (int i = 0; i < K; ++i) {
                        const int x = xa + i;
                        pd[i] = lrow[aa]; qd[i] = lrow[c];
                        if (x >= 0 && x + 1 < L) { if (++aa == b) { aa = 0; ++b; lrow += n; } }            // next entry's row
                    }
#pragma unroll
                    for (int i = 0; i < K; ++i) {
                        const int x = xa + i;
                        const int p = (int)(pd[i] & 0xffffu), dp = (int)(pd[i] >> 16), q = (int)(qd[i] & 0xffffu), dq = (int)(qd[i] >> 16);
                        const int S0 = dp + dr, S2 = min(dp, min(dq, dr)) + dq;
                        const int lq = dq - dr - 1, lp = dp - dr - 1;
                        const int s_ab = dr > dq ? (dp > dq ? p : q) : n + p;                               // ab|cd: (deeper(p,q), r) or (p, q)
                        const int s_ad = dp > dr ? (lq < LV ? 2 * n + lq * (lq - 1) / 2 + lp : -2) : q;     // ad|bc: (q, p) or (q, r)
                        slot[i] = (x < 0 || x >= L) ? -1 : S0 > S2 ? s_ab : S2 > S0 ? s_ad : -1;            // -1: outside, or unresolved in the reference tree (:559-562)
                    }
                    // sums: consecutive entries mostly share their slot (the key changes with lca(a,b), every ~8 entries): added up in
                    // registers, one set of shared-memory REDs per run (per-entry REDs from 32 lanes to 3-4 addresses serialise)
                    uint32_t r1s = 0, r2s = 0, r3s = 0;
#pragma unroll
                    for (int i = 0; i < K; ++i) {
                        if (slot[i] == -1) continue;
                        uint32_t r0, r1, r2;
                        if (RING) {                               // halfwords 3i, 3i+1, 3i+2 of w[]
                            r0 = (i & 1) ? w[(3 * i) >> 1] >> 16 : w[(3 * i) >> 1] & 0xffffu;
                            r1 = (i & 1) ? w[(3 * i + 1) >> 1] & 0xffffu : w[(3 * i + 1) >> 1] >> 16;
                            r2 = (i & 1) ? w[(3 * i + 2) >> 1] >> 16 : w[(3 * i + 2) >> 1] & 0xffffu;
                        } else {
                            const CINT* e = table + (E0 + (uint64_t)(xa + i)) * 3;
                            r0 = (uint32_t)e[0]; r1 = (uint32_t)e[1]; r2 = (uint32_t)e[2];
                        }
                        const uint32_t c0 = (r0 << shift) & mask32, c1 = (r1 << shift) & mask32, c2 = (r2 << shift) & mask32;
                        const int dp = (int)(pd[i] >> 16), dq = (int)(qd[i] >> 16);
                        const bool rs2 = dp + dr <= min(dp, min(dq, dr)) + dq;                              // not ab|cd, i.e. ad|bc
                        const uint32_t a1 = rs2 ? c2 : c0, a3 = rs2 ? c0 : c2;                              // (reference topology, crossing, other)
                        if (slot[i] >= 0) {
                            if (bif) {
                                r1s += a1; r2s += c1; r3s += a3;
                                if (i == K - 1 || slot[i + (i < K - 1 ? 1 : 0)] != slot[i]) {
                                    add_sum(0, slot[i], r1s); add_sum(1, slot[i], r2s); add_sum(2, slot[i], r3s);
                                    if (slot[i] >= n && slot[i] < 2 * n) acc_pq[slot[i] - n] = (uint16_t)(qd[i] & 0xffffu);
                                    r1s = r2s = r3s = 0;
                                }
                            }
                            // LQ-IC.  First the cheap exclusion (tau), then the fp32 estimate against the slot's bound; the rare quartet that may
                            // beat it is evaluated exactly (scan_cold)
                            const float tau = ordered_to_float(*reinterpret_cast<volatile int*>(acc_tau + slot[i]));
                            if (!(a1 != 0 && a1 >= max(c1, a3) && (float)a1 >= tau * (float)(a1 + c1 + a3))) {      // (all-zero counts score 0, not >= g)
                                float est = dev_log_score_f32(a1, c1, a3);
                                const bool exact = score_known_exactly(a1, c1, a3, est);
                                const float bound = ordered_to_float(*reinterpret_cast<volatile int*>(acc_b + slot[i]));
                                if ((exact ? est : est - QS_EST_EPS) < bound) {
                                    const int p = (int)(pd[i] & 0xffffu), q = (int)(qd[i] & 0xffffu);
                                    int u, v;
                                    if (slot[i] < n) { u = slot[i]; v = r; } else if (slot[i] < 2 * n) { u = p; v = q; } else { u = q; v = p; }
                                    // its score is <= ub: whoever cannot beat ub needs no look (the candidate itself still passes: est - eps < ub)
                                    const float ub = exact ? est : est + QS_EST_EPS;
                                    atomicMin(acc_b + slot[i], float_to_ordered(ub));
                                    atomicMin(acc_tau + slot[i], float_to_ordered(qs_tau_lookup(s_tau, ub)));
                                    const int qi = atomicAdd(wq_cnt, 1);
                                    if (qi < QCAP) wq[qi] = ScanCand{c0, c1, c2, (uint32_t)u | ((uint32_t)v << 16), slot[i] * 2 + (rs2 ? 1 : 0)};
                                    else scan_cold(a.pair_sums, a.pair_best, a.pair_score, a.I, a.bifurcating, rs2 ? 2 : 0, c0, c1, c2, u, v, acc_b + slot[i], acc_tau + slot[i], s_tau, false);     // queue full
                                }
                            }
                        } else {                                  // slot -2: key (q, p) deeper than accQ reaches: sums and selection in global memory
                            scan_cold(a.pair_sums, a.pair_best, a.pair_score, a.I, a.bifurcating, 2, c0, c1, c2, (int)(qd[i] & 0xffffu), (int)(pd[i] & 0xffffu), nullptr, nullptr, s_tau, bif);
                        }
                    }
                }
                if (RING) {                                       // hand the stage back: one arrival per consumer warp
                    __syncwarp();
                    if ((tid & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[stage])) : "memory");
                }
                __syncwarp();
                if (*reinterpret_cast<volatile int*>(wq_cnt) >= QCAP / 2) drain();       // (after the hand-over: the next chunk lands meanwhile)
            }
        }
        drain();                                                 // the slots' bounds belong to this item
        }
        __syncthreads();
        // ---- flush the CTA's accumulators: one global atomic per touched key ----
        if (bif) {
            auto flush = [&](int slot, long long key) {
                unsigned long long* ps = a.pair_sums + (size_t)key * 3;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const unsigned long long v = (unsigned long long)acc_lo[k * n_acc + slot] | (CARRY ? (unsigned long long)acc_hi[k * n_acc + slot] << 32 : 0ull);
                    if (v) atomicAdd(ps + k, v);
                }
            };
            auto touched = [&](int y) -> bool {
                uint32_t v = acc_lo[y] | acc_lo[n_acc + y] | acc_lo[2 * n_acc + y];
                if (CARRY) v |= acc_hi[y] | acc_hi[n_acc + y] | acc_hi[2 * n_acc + y];
                return v != 0;
            };
            for (int x = tid; x < c && !producer; x += THREADS) {
                if (touched(x)) flush(x, pair_key(x, r));
                if (touched(n + x)) flush(n + x, pair_key(x, (int)acc_pq[x]));
            }
            for (int x = tid; x < nq4 && !producer; x += THREADS) {
                if (touched(2 * n + x)) { const int i = tri_row(x); flush(2 * n + x, pair_key(s_anc[i], s_anc[x - i * (i - 1) / 2])); }
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
