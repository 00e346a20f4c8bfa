// score.cuh — per-quartet scan of the count table: QIC per quartet, aggregated per reference inner-node pair.
//
// Replaces QuartetScoreComputer::processNodePair / computeQuartetScoresBifurcating /
// computeQuartetScoresMultifurcating (src/QuartetScoreComputer.hpp:379-593).  The reference walks, for
// every quartet, the path between the two inner nodes u,v that the quartet's central edge path connects
// and takes a critical section per edge.  Here the scan is quartet-centric:
//
//   * taxon ids follow the reference tree's planar leaf order, so for sorted a<b<c<d only the
//     non-crossing pairings can be the reference topology.  With p = lca(a,b), q = lca(b,c), r = lca(c,d)
//     and their depths:  S0 = dp+dr, S2 = min(dp,dq,dr)+dq;  S0 > S2 -> ab|cd (slot 0), ends
//     u = deeper(p,q), v = deeper(q,r);  S2 > S0 -> ad|bc (slot 2), ends u = q, v = deeper(p,r);
//     equal -> unresolved in a multifurcating reference, skipped (:559-562).
//   * per unordered pair {u,v}: three integer sums (q_ref, q_slot1, q_other) for QP-IC/EQP-IC (:429-431,472)
//     and the quartet with minimal QIC (its count triple, so that the host can re-evaluate log_score with
//     the reference's libm and operation order) for LQ-IC (:432-454).
//   * the per-edge minima over paths are a tiny host post-pass over the pairs.
//
// Thread = one (c,d) pair for a fixed b (uniform per block); it walks a = 0..b-1, i.e. a contiguous run of
// table entries.  lca(a,b) is piecewise constant along the walk, so aggregates are kept in registers and
// flushed with a handful of atomics only when the pair key changes.
#pragma once
#include "common.cuh"

#ifndef QS_SCORE_L2_PREFETCH
#define QS_SCORE_L2_PREFETCH 1      // table reads carry the L2::256B prefetch hint (see ldg_l2_256)
#endif

namespace qs {

struct ScoreArgs {
    const void* table;         // CINT [(rank - rank_base)][3]
    uint64_t rank_base;
    const uint16_t* lca;       // [n][n] inner index of lca(leaf a, leaf b)
    const uint16_t* idepth;    // [I] depth of inner node (by inner index)
    unsigned long long* pair_sums;   // [I*I][3]
    unsigned long long* pair_best;   // [I*I] packed triple of the min-QIC quartet, ~0 = none
    int* pair_hint;                  // [I*I] order-preserving int image of an fp32 UPPER bound of the pair's current best score
    const int64_t* PB;         // [n+1] first block of each b
    int n, I;
    int d_begin, d_end;
    int count_scale;           // 1 or 2
    unsigned long long cint_mask;
    int bifurcating;           // argument order of the stored triple (see pack below)
};

constexpr unsigned long long QS_TRIPLE_NONE = ~0ull;
constexpr int QS_TRIPLE_BITS = 21;

__device__ __forceinline__ unsigned long long pack_triple(unsigned long long q1, unsigned long long q2, unsigned long long q3) {
    return (q1 << (2 * QS_TRIPLE_BITS)) | (q2 << QS_TRIPLE_BITS) | q3;
}

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

__device__ __forceinline__ double triple_score(unsigned long long t) {
    if (t == QS_TRIPLE_NONE) return INFINITY;
    const unsigned long long M = (1ull << QS_TRIPLE_BITS) - 1;
    return dev_log_score(t >> (2 * QS_TRIPLE_BITS), (t >> QS_TRIPLE_BITS) & M, t & M);
}

// Selecting the minimum-QIC quartet of a pair WITHOUT an fp64 log_score per quartet.  Every quartet gets an fp32
// estimate (|estimate - exact| < 1e-5: three p*log2(p) terms of magnitude <= 0.53, each a few fp32 ulps off).  A quartet
// can be the exact minimum only if its estimate is within QS_FILTER_MARGIN (>= 2 x that error) of the smallest
// estimate / exact value seen so far, so only those are kept as CANDIDATES (two register slots) and evaluated in fp64
// when the run ends — at the same loop iteration for all threads of the block (key runs follow lca(a,b), which is
// block-uniform), so the expensive path is executed converged and ~2 times per run instead of once per quartet.
constexpr float QS_FILTER_MARGIN = 2e-6f;      // estimate error is < 5e-7 (see dev_log_score_f32)
// order-preserving float <-> int (so that atomicMin on ints orders floats, negatives included)
__device__ __forceinline__ int float_to_ordered(float f) { const int b = __float_as_int(f); return b ^ ((b >> 31) & 0x7fffffff); }
__device__ __forceinline__ float ordered_to_float(int o) { return __int_as_float(o ^ ((o >> 31) & 0x7fffffff)); }
constexpr int QS_HINT_NONE = 0x7f800000;         // +inf
__device__ __forceinline__ float dev_log_score_f32(unsigned q1, unsigned q2, unsigned q3) {
    const unsigned s = q1 + q2 + q3;
    if (s == 0) return 0.f;
    const float inv = 1.f / (float)s, il3 = 0.63092975357145743710f;      // 1/log2(3)
    float acc = 0.f;
    // __log2f = MUFU.LG2: absolute error <= 2^-22 on [0.5, 2], <= 2 ulp elsewhere, i.e. <= 2.4e-7 per p*log2(p) term and
    // <= 3e-7 on the estimate; with the fp32 rounding of the rest (2e-7 measured, tests/test_score_filter.py) the estimate is
    // within 5e-7 of the exact score — a quarter of QS_FILTER_MARGIN
    if (q1) { const float p = (float)q1 * inv; acc += p * __log2f(p); }
    if (q2) { const float p = (float)q2 * inv; acc += p * __log2f(p); }
    if (q3) { const float p = (float)q3 * inv; acc += p * __log2f(p); }
    const float qic = 1.f + acc * il3;
    return ((q1 < q2) || (q1 < q3)) ? -qic : qic;
}

// 16-byte read-only load that asks L2 to fetch the whole 256-byte neighbourhood from DRAM: a thread walks its run 48 bytes
// at a time, so the next five visits then hit in L2 instead of opening the DRAM page again (the scan is bound by DRAM row
// locality)
__device__ __forceinline__ uint4 ldg_l2_256(const uint4* p) {
    uint4 v;
#if QS_SCORE_L2_PREFETCH == 2
    // ... and marks the lines evict-first: the table is read once, while the few MB of pair sums / best / hint arrays that
    // the REDs hit should stay in L2 (39 % of the RED sectors missed in L2 with the default policy)
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("ld.global.nc.L2::cache_hint.L2::256B.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
#elif QS_SCORE_L2_PREFETCH
    asm volatile("ld.global.nc.L2::256B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
#else
    v = __ldg(p);
#endif
    return v;
}

struct PairAcc {
    unsigned long long s1, s2, s3;
    unsigned long long best;       // exact champion so far (resolved candidates), QS_TRIPLE_NONE = none
    double best_q;                 // its exact score
    unsigned long long cand0, cand1, last_t;
    float bound;                   // smallest estimate / champion score seen in this run
    int hint_bits;                 // pair_hint[key], requested when the run starts so that it has arrived by the time the run ends
    int ncand;
    int key;                       // pair index or -1
};

__device__ __forceinline__ void pair_reset(PairAcc& acc) {
    acc.s1 = acc.s2 = acc.s3 = 0; acc.best = QS_TRIPLE_NONE; acc.best_q = INFINITY; acc.key = -1;
    acc.cand0 = acc.cand1 = acc.last_t = QS_TRIPLE_NONE; acc.bound = INFINITY; acc.ncand = 0; acc.hint_bits = QS_HINT_NONE;
}

// evaluate the pending candidates exactly and fold them into the champion
__device__ __forceinline__ void pair_resolve(PairAcc& acc) {
#pragma unroll 1
    for (int i = 0; i < 2; ++i) {
        if (i < acc.ncand) {
            const unsigned long long t = i ? acc.cand1 : acc.cand0;
            const double q = triple_score(t);
            if (q < acc.best_q) { acc.best_q = q; acc.best = t; }
        }
    }
    acc.ncand = 0;
    if (acc.best != QS_TRIPLE_NONE) acc.bound = fminf(acc.bound, (float)acc.best_q);
}

__device__ __forceinline__ void pair_flush(const ScoreArgs& a, PairAcc& acc) {
    if (acc.key < 0) return;
    unsigned long long* ps = a.pair_sums + (size_t)acc.key * 3;
    if (acc.s1) atomicAdd(ps + 0, acc.s1);
    if (acc.s2) atomicAdd(ps + 1, acc.s2);
    if (acc.s3) atomicAdd(ps + 2, acc.s3);
    // pair_hint never lies below the pair's true current minimum (it is lowered only AFTER a successful update of
    // pair_best, to a value rounded up), so a run whose every quartet is estimated above it cannot improve the pair:
    // no fp64 evaluation, no CAS.  Most runs end here.
    const float hint = ordered_to_float(acc.hint_bits);      // read at the start of the run: stale only towards +inf, i.e. conservative
    if (acc.bound <= hint + QS_FILTER_MARGIN) {
        pair_resolve(acc);
        if (acc.best != QS_TRIPLE_NONE) {
            unsigned long long* pb = a.pair_best + acc.key;
            unsigned long long old = *reinterpret_cast<volatile unsigned long long*>(pb);
            bool won = false;
            while (true) {
                if (!(acc.best_q < triple_score(old))) break;
                unsigned long long prev = atomicCAS(pb, old, acc.best);
                if (prev == old) { won = true; break; }
                old = prev;
            }
            if (won) atomicMin(a.pair_hint + acc.key, float_to_ordered(__double2float_ru(acc.best_q)));
        }
    }
    pair_reset(acc);
}

// one quartet's contribution; (c0,c1,c2) = table slots after scale/mask
__device__ __forceinline__ void pair_add(const ScoreArgs& a, PairAcc& acc, int key, int rslot,
                                         unsigned long long c0, unsigned long long c1, unsigned long long c2) {
    if (key != acc.key) {
        pair_flush(a, acc);
        acc.key = key;
        acc.hint_bits = *reinterpret_cast<volatile int*>(a.pair_hint + key);      // in flight while the run is scanned (the flush stalled ~600 clk on it)
    }
    unsigned long long q1, q2, q3;
    if (rslot == 0) { q1 = c0; q2 = c1; q3 = c2; }
    else if (a.bifurcating) { q1 = c2; q2 = c1; q3 = c0; }      // processNodePair order: (ref, S1S3|S2S4, S1S4|S2S3)
    else { q1 = c2; q2 = c0; q3 = c1; }                          // (u,z|v,w): ab|cd, ac|bd = uv|zw, ad|bc = uw|zv
    acc.s1 += q1; acc.s2 += q2; acc.s3 += q3;
    const unsigned long long t = pack_triple(q1, q2, q3);
    if (t == acc.last_t) return;                                 // same counts as the previous quartet: nothing new
    acc.last_t = t;
    if ((q1 | q2 | q3) >= (1ull << 24)) {                        // wider than fp32 integers (uint32/uint64 CINT): exact path
        const double q = dev_log_score(q1, q2, q3);
        if (q < acc.best_q) { acc.best_q = q; acc.best = t; acc.bound = fminf(acc.bound, (float)q); }
        return;
    }
    const float est = dev_log_score_f32((unsigned)q1, (unsigned)q2, (unsigned)q3);
    if (est > acc.bound + QS_FILTER_MARGIN) return;              // cannot be the minimum of this run
    if (est < acc.bound - QS_FILTER_MARGIN) acc.ncand = 0;       // strictly better than everything pending: they are obsolete
    acc.bound = fminf(acc.bound, est);
    if ((acc.ncand > 0 && t == acc.cand0) || (acc.ncand > 1 && t == acc.cand1)) return;
    if (acc.ncand == 2) pair_resolve(acc);                       // both slots taken: settle them exactly, keep the champion
    if (acc.ncand == 0) acc.cand0 = t; else acc.cand1 = t;
    acc.ncand++;
}

// reference topology + pair key of sorted quartet (a,b,c,d) from the three adjacent LCAs (inner indices)
__device__ __forceinline__ int quartet_pair_key(const ScoreArgs& a, int p, int q, int r, int dp, int dq, int dr, int& rslot) {
    const int S0 = dp + dr, S2 = min(dp, min(dq, dr)) + dq;
    int u, v;
    if (S0 > S2) { rslot = 0; u = (dp > dq) ? p : q; v = (dr > dq) ? r : q; }
    else if (S2 > S0) { rslot = 2; u = q; v = (dp > dr) ? p : r; }
    else { rslot = -1; return -1; }
    return (u < v) ? u * a.I + v : v * a.I + u;
}

__global__ void qs_fill_int_kernel(int* __restrict__ p, size_t n, int v) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

#ifndef QS_SCORE_PREFETCH
#define QS_SCORE_PREFETCH 0
#endif
#ifndef QS_SCORE_MIN_BLOCKS
#define QS_SCORE_MIN_BLOCKS 6      // CTAs of 128 threads per SM the compiler must leave room for (80 registers; 4 -> 104 regs is 6 % slower, 8 spills: profiles/r01_za_score_variants.txt)
#endif
template <typename CINT>
__global__ void __launch_bounds__(128, QS_SCORE_MIN_BLOCKS) qs_score_table_kernel(const ScoreArgs a) {
    // block -> b (uniform), thread -> (c,d) pair
    const long long blk = blockIdx.x;
    int lo = 1, hi = a.n - 2;              // b in [1, n-3]
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (a.PB[mid] <= blk) lo = mid; else hi = mid; }
    const int b = lo;
    // lca(a,b) and its depth for every a < b are the same for the whole block: staged once in shared memory as
    // (inner index | depth << 16).  Read per quartet from global they were a chain of dependent L1 loads and, with the table
    // reads, half of all warp stalls (profiles/r01_ze_*).
    extern __shared__ uint32_t s_pd[];
    {
        const uint16_t* lrow_g = a.lca + (size_t)b * a.n;
        for (int x = threadIdx.x; x < b; x += blockDim.x) { const uint32_t p = lrow_g[x]; s_pd[x] = p | ((uint32_t)a.idepth[p] << 16); }
    }
    __syncthreads();
    const long long j = (blk - a.PB[b]) * blockDim.x + threadIdx.x;   // pair index for this b
    // pairs (c,d): d in [max(b+2,d_begin), d_end), c in (b, d); enumerate d-major
    // count of pairs with d' < d : sum_{d'=dlo}^{d-1} (d'-b-1)
    const int dlo = max(b + 2, a.d_begin);
    if (dlo >= a.d_end) return;
    // solve for d: f(d) = sum_{x=dlo}^{d-1} (x-b-1) = T(d-b-1) - T(dlo-b-1), T(k)=k(k-1)/2 ... use k = x-b-1
    const long long k0 = dlo - b - 1;                        // first k
    const long long base = k0 * (k0 - 1) / 2;
    // find k >= k0 with  k(k-1)/2 - base <= j < (k+1)k/2 - base
    long long k = (long long)((1.0 + sqrt(1.0 + 8.0 * (double)(j + base))) * 0.5);
    while (k * (k - 1) / 2 - base > j) --k;
    while ((k + 1) * k / 2 - base <= j) ++k;
    const int d = (int)(k + b + 1);
    if (d >= a.d_end) return;
    const int c = b + 1 + (int)(j - (k * (k - 1) / 2 - base));

    const int q = a.lca[(size_t)b * a.n + c], r = a.lca[(size_t)c * a.n + d];
    const int dq = a.idepth[q], dr = a.idepth[r];
    const CINT* tab = reinterpret_cast<const CINT*>(a.table) + (quartet_rank(0, b, c, d) - a.rank_base) * 3;

    PairAcc acc;
    pair_reset(acc);
    int last_p = -1, key = -1, rslot = -1;
    auto one = [&](int x, unsigned long long r0, unsigned long long r1, unsigned long long r2) {
        const uint32_t pd = s_pd[x];
        const int p = (int)(pd & 0xffffu);
        if (p != last_p) { last_p = p; key = quartet_pair_key(a, p, q, r, (int)(pd >> 16), dq, dr, rslot); }
        if (key < 0) return;
        pair_add(a, acc, key, rslot, (r0 * a.count_scale) & a.cint_mask, (r1 * a.count_scale) & a.cint_mask, (r2 * a.count_scale) & a.cint_mask);
    };
    int x = 0;
    if (sizeof(CINT) == 2) {
        // the run of b entries is contiguous: read it 8 entries (48 bytes = 3 x LDG.128) at a time once the entry index is a
        // multiple of 8 — a thread's scalar 2-byte loads cost one L1 wavefront each, 24 per 8 entries instead of 3.  The 48
        // bytes sit in 12 registers used as a shift register (6 bytes out per entry), so the loop body — with its fp64
        // log_score and atomics — exists ONCE: unrolling it 8x made the kernel 11k instructions and 30x slower
        // (instruction-cache misses under divergence, profiles/r01_k_*).
        const uint64_t e0 = quartet_rank(0, b, c, d) - a.rank_base;
        for (; x < b && ((e0 + x) & 7); ++x) one(x, tab[x * 3 + 0], tab[x * 3 + 1], tab[x * 3 + 2]);
#if QS_SCORE_PREFETCH
        // the next 48 bytes are requested before the current 8 entries are processed: the scan is latency-bound
        // (issue-active 51 %, long-scoreboard stalls on these loads: profiles/r01_n_score_table_n500_ncu_full.txt)
        uint4 n0 = make_uint4(0, 0, 0, 0), n1 = n0, n2 = n0;
        if (x + 8 <= b) { const uint4* v = reinterpret_cast<const uint4*>(tab + (size_t)x * 3); n0 = ldg_l2_256(v); n1 = ldg_l2_256(v + 1); n2 = ldg_l2_256(v + 2); }
        for (; x + 8 <= b; x += 8) {
            uint32_t w[12] = {n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, n1.z, n1.w, n2.x, n2.y, n2.z, n2.w};
            if (x + 16 <= b) { const uint4* v = reinterpret_cast<const uint4*>(tab + (size_t)(x + 8) * 3); n0 = ldg_l2_256(v); n1 = ldg_l2_256(v + 1); n2 = ldg_l2_256(v + 2); }
#else
        for (; x + 8 <= b; x += 8) {
            const uint4* v = reinterpret_cast<const uint4*>(tab + (size_t)x * 3);
            const uint4 w0 = ldg_l2_256(v), w1 = ldg_l2_256(v + 1), w2 = ldg_l2_256(v + 2);
            uint32_t w[12] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w};
#endif
#pragma unroll 1
            for (int i = 0; i < 8; ++i) {
                one(x + i, w[0] & 0xffffu, w[0] >> 16, w[1] & 0xffffu);
#pragma unroll
                for (int k = 0; k < 10; ++k) w[k] = __funnelshift_r(w[k + 1], w[k + 2], 16);      // >> 48 bits
                w[10] = w[11] >> 16;
            }
        }
    }
    for (; x < b; ++x) one(x, tab[x * 3 + 0], tab[x * 3 + 1], tab[x * 3 + 2]);
    pair_flush(a, acc);
}

}  // namespace qs
