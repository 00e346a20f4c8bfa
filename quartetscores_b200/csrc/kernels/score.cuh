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

namespace qs {

struct ScoreArgs {
    const void* table;         // CINT [(rank - rank_base)][3]
    uint64_t rank_base;
    const uint16_t* lca;       // [n][n] inner index of lca(leaf a, leaf b)
    const uint16_t* idepth;    // [I] depth of inner node (by inner index)
    unsigned long long* pair_sums;   // [I*I][3]
    unsigned long long* pair_best;   // [I*I] packed triple of the min-QIC quartet, ~0 = none
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
__device__ __forceinline__ double dev_log_score(unsigned long long q1, unsigned long long q2, unsigned long long q3) {
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

struct PairAcc {
    unsigned long long s1, s2, s3, best;
    double best_q;
    int key;   // pair index or -1
};

__device__ __forceinline__ void pair_flush(const ScoreArgs& a, PairAcc& acc) {
    if (acc.key < 0) return;
    unsigned long long* ps = a.pair_sums + (size_t)acc.key * 3;
    if (acc.s1) atomicAdd(ps + 0, acc.s1);
    if (acc.s2) atomicAdd(ps + 1, acc.s2);
    if (acc.s3) atomicAdd(ps + 2, acc.s3);
    if (acc.best != QS_TRIPLE_NONE) {
        unsigned long long* pb = a.pair_best + acc.key;
        unsigned long long old = *reinterpret_cast<volatile unsigned long long*>(pb);
        while (true) {
            if (!(acc.best_q < triple_score(old))) break;
            unsigned long long prev = atomicCAS(pb, old, acc.best);
            if (prev == old) break;
            old = prev;
        }
    }
    acc.s1 = acc.s2 = acc.s3 = 0; acc.best = QS_TRIPLE_NONE; acc.best_q = INFINITY; acc.key = -1;
}

// one quartet's contribution; (c0,c1,c2) = table slots after scale/mask
__device__ __forceinline__ void pair_add(const ScoreArgs& a, PairAcc& acc, int key, int rslot,
                                         unsigned long long c0, unsigned long long c1, unsigned long long c2,
                                         unsigned long long& memo_t, double& memo_q) {
    if (key != acc.key) { pair_flush(a, acc); acc.key = key; }
    unsigned long long q1, q2, q3;
    if (rslot == 0) { q1 = c0; q2 = c1; q3 = c2; }
    else if (a.bifurcating) { q1 = c2; q2 = c1; q3 = c0; }      // processNodePair order: (ref, S1S3|S2S4, S1S4|S2S3)
    else { q1 = c2; q2 = c0; q3 = c1; }                          // (u,z|v,w): ab|cd, ac|bd = uv|zw, ad|bc = uw|zv
    acc.s1 += q1; acc.s2 += q2; acc.s3 += q3;
    const unsigned long long t = pack_triple(q1, q2, q3);
    if (t != memo_t) { memo_t = t; memo_q = dev_log_score(q1, q2, q3); }
    if (memo_q < acc.best_q) { acc.best_q = memo_q; acc.best = t; }
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

template <typename CINT>
__global__ void __launch_bounds__(128) qs_score_table_kernel(const ScoreArgs a) {
    // block -> b (uniform), thread -> (c,d) pair
    const long long blk = blockIdx.x;
    int lo = 1, hi = a.n - 2;              // b in [1, n-3]
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (a.PB[mid] <= blk) lo = mid; else hi = mid; }
    const int b = lo;
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
    const uint16_t* lrow = a.lca + (size_t)b * a.n;

    PairAcc acc; acc.s1 = acc.s2 = acc.s3 = 0; acc.best = QS_TRIPLE_NONE; acc.best_q = INFINITY; acc.key = -1;
    unsigned long long memo_t = QS_TRIPLE_NONE; double memo_q = 0.0;
    int last_p = -1, key = -1, rslot = -1;
    for (int x = 0; x < b; ++x) {
        const int p = lrow[x];
        if (p != last_p) { last_p = p; key = quartet_pair_key(a, p, q, r, a.idepth[p], dq, dr, rslot); }
        if (key < 0) continue;
        const unsigned long long c0 = ((unsigned long long)tab[x * 3 + 0] * a.count_scale) & a.cint_mask;
        const unsigned long long c1 = ((unsigned long long)tab[x * 3 + 1] * a.count_scale) & a.cint_mask;
        const unsigned long long c2 = ((unsigned long long)tab[x * 3 + 2] * a.count_scale) & a.cint_mask;
        pair_add(a, acc, key, rslot, c0, c1, c2, memo_t, memo_q);
    }
    pair_flush(a, acc);
}

}  // namespace qs
