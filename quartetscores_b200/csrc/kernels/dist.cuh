// dist.cuh — per-gene-tree topological distance matrices.
//
// Reference semantics: TreeInformation::distanceInEdges (src/TreeInformation.hpp:40-43):
// d(u,v) = depth[u] + depth[v] - 2*depth[lca(u,v)], with the LCA from an Euler tour + RMQ
// (:71-76,95-113).  The reference applies it to the reference tree only; the B200 design applies the
// same formula to every gene tree and feeds the counting kernel with the n x n matrices.
//
// One CTA per tree (grid-stride).  Flat tree encoding -> (a) depths and an Euler-tour leaf sequence
// (one thread, O(N); the tree is tiny), (b) all leaf pairs in parallel: the LCA depth of the leaves at
// tour positions i<j is the minimum of the "turning depths" h[i..j-1] between consecutive leaves, taken
// as a running minimum while all lanes sweep the same j so that stores go to one matrix row.
//
// Output: D[tree][row][col] as IEEE fp16 (exact for integers <= 2048), NaN where either taxon is
// absent from the tree (such a tree's matrix is filled with 0xFFFF = NaN before the pair sweep), row pitch n_pad
// (multiple of 8); the padding columns n..n_pad-1 are never initialised and only ever feed lanes whose results are dropped.
#pragma once
#include "common.cuh"

namespace qs {

struct DistArgs {
    const int64_t* node_off;   // [m+1]
    const int32_t* parent;     // concatenated, local indices, parent[i] < i
    const int32_t* leaf_id;    // lookup id or -1
    int m, n, n_pad, max_nodes;
    __half* D;
    int* max_dist;             // [0] global max over all trees (atomicMax), [1] first tree-encoding error code
    int32_t* tree_class;       // [m] 0 = class A: all n taxa present and no node of degree > 3 (every quartet resolved), 1 = class B
};

__global__ void __launch_bounds__(256) qs_dist_kernel(DistArgs a) {
    extern __shared__ int32_t sm_i[];
    const int NMAX = a.max_nodes;
    int32_t* par = sm_i;              // [NMAX]
    int32_t* fc = par + NMAX;         // first child / DFS iterator
    int32_t* ns = fc + NMAX;          // next sibling
    int32_t* dep = ns + NMAX;         // depth
    int32_t* stk = dep + NMAX;        // DFS stack
    int32_t* ltid = stk + NMAX;       // per leaf (tour order): taxon id
    int32_t* ldep = ltid + NMAX;      // depth
    int32_t* lh = ldep + NMAX;        // lh[i] = depth of lca(leaf i, leaf i+1)
    uint32_t* seen = reinterpret_cast<uint32_t*>(lh + NMAX);   // [(n+31)/32] taxon bitmap (duplicate check)
    __shared__ int s_k;
    int local_max = 0;

    for (int t = blockIdx.x; t < a.m; t += gridDim.x) {
        const int64_t off = a.node_off[t];
        const int N = (int)(a.node_off[t + 1] - off);
        __syncthreads();
        for (int i = threadIdx.x; i < N; i += blockDim.x) { par[i] = a.parent[off + i]; fc[i] = -1; ns[i] = -1; lh[i] = 0; }
        for (int i = threadIdx.x; i < (a.n + 31) / 32; i += blockDim.x) seen[i] = 0u;
        __syncthreads();
        if (threadIdx.x == 0) {
            int bad = 0;
            dep[0] = 0;
            for (int i = 1; i < N; ++i) { if (par[i] < 0 || par[i] >= i) { bad = 1; par[i] = 0; } }
            for (int i = 1; i < N; ++i) dep[i] = dep[par[i]] + 1;
            int maxdeg = 0;                                   // lh[] holds child counts until the tour overwrites it
            for (int i = N - 1; i >= 1; --i) { int p = par[i]; ns[i] = fc[p]; fc[p] = i; maxdeg = max(maxdeg, ++lh[p] + (p > 0 ? 1 : 0)); }
            int k = 0, sp = 0, cur_min = 0x7fffffff;
            if (N > 0 && fc[0] == -1) {                       // single-node tree
                int id = a.leaf_id[off];
                if (id >= 0 && id < a.n) { ltid[0] = id; ldep[0] = 0; k = 1; } else bad = 2;
            } else if (N > 0) {
                stk[sp++] = 0;
                while (sp > 0) {
                    int v = stk[sp - 1];
                    int c = fc[v];
                    if (c == -1) { --sp; continue; }
                    fc[v] = ns[c];                            // advance iterator
                    cur_min = min(cur_min, dep[v]);           // v is a turning point on the way to the next leaf
                    if (fc[c] == -1) {                        // leaf
                        if (k > 0) lh[k - 1] = cur_min;
                        int id = a.leaf_id[off + c];
                        if (id < 0 || id >= a.n) { bad = 2; id = 0; }
                        else if (seen[id >> 5] & (1u << (id & 31))) bad = 3;
                        else seen[id >> 5] |= 1u << (id & 31);
                        ltid[k] = id; ldep[k] = dep[c]; ++k;
                        cur_min = 0x7fffffff;
                    } else {
                        if (a.leaf_id[off + c] >= 0) bad = 4;   // inner node carrying a taxon id
                        stk[sp++] = c;
                    }
                }
            }
            if (bad) { atomicCAS(a.max_dist + 1, 0, bad); k = 0; }
            a.tree_class[t] = (!bad && k == a.n && maxdeg <= 3) ? 0 : 1;
            s_k = k;
        }
        __syncthreads();
        const int k = s_k;
        __half* Dt = a.D + (size_t)t * a.n * a.n_pad;
        if (k < a.n) {                                        // some taxon is absent (or the tree is malformed): NaN everywhere first
            uint4* w = reinterpret_cast<uint4*>(Dt);          // (a tree that has every taxon writes all n x n entries below, so the
            const size_t nw16 = (size_t)a.n * a.n_pad / 8;    //  host no longer memsets the whole buffer: 208 MB per step at cfg2)
            for (size_t x = threadIdx.x; x < nw16; x += blockDim.x) w[x] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
            __syncthreads();
        }
        for (int i0 = 0; i0 < k; i0 += blockDim.x) {
            const int i = i0 + threadIdx.x;
            const bool act = i < k;
            const int ti = act ? ltid[i] : 0, di = act ? ldep[i] : 0;
            if (act) Dt[(size_t)ti * a.n_pad + ti] = __int2half_rn(0);
            // sweep 1: j ascending, lanes with i < j
            int mn = 0x7fffffff;
            const int jlo = i0 + 1;                           // first j any lane of this block-iteration needs
            for (int j = jlo; j < k; ++j) {
                if (act && j > i) {
                    mn = min(mn, lh[j - 1]);
                    int d = di + ldep[j] - 2 * mn;
                    local_max = max(local_max, d);
                    Dt[(size_t)ltid[j] * a.n_pad + ti] = __int2half_rn(d);
                }
            }
            // sweep 2: j descending, lanes with i > j
            mn = 0x7fffffff;
            const int jhi = min(k - 1, i0 + (int)blockDim.x - 1) - 1;
            for (int j = jhi; j >= 0; --j) {
                if (act && j < i) {
                    mn = min(mn, lh[j]);
                    int d = di + ldep[j] - 2 * mn;
                    Dt[(size_t)ltid[j] * a.n_pad + ti] = __int2half_rn(d);
                }
            }
        }
    }
    // block max -> global
    for (int o = 16; o > 0; o >>= 1) local_max = max(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
    if ((threadIdx.x & 31) == 0 && local_max > 0) atomicMax(a.max_dist, local_max);
}

// ---- warp-per-tree variant for small gene trees ------------------------------------------------------------
// The CTA-per-tree kernel above leaves 255 threads idle while one thread walks the tree, and at 100 taxa that
// serial walk is most of its time (0.78 ms for 10,000 trees, profiles/r01_e_launches.csv).  Here every WARP owns a
// tree, up to 64 trees per SM are in flight, and nothing is walked by one thread: with parent[i] < i as the only
// assumption about the node order,
//   * depth[i] and, later, the tour position of the first leaf of subtree i are sums along the root path: pointer
//     jumping on packed (ancestor, partial sum) words, in place — a node reads its ancestor's word in one 32-bit load, so
//     any interleaving of the lanes keeps "sum over the path up to, not including, the ancestor" true; ~log2(depth) rounds;
//   * child counts, leaves per subtree (every leaf adds one to each of its ancestors) and a child's offset inside its
//     parent's leaf interval (fetch-and-add of its leaf count on the parent: children are placed in arrival order, any
//     planar order gives the same matrix) are shared-memory atomics on 16-bit fields;
//   * a tree whose leaves lie deeper than 16 edges on average (caterpillars) counts its subtree leaves in one serial
//     backward pass instead, as round 1 did for every tree.
// Validation, the leaf arrays, the turning depths lh[q] = depth of lca(leaf q, leaf q+1), written by the unique non-first
// child whose subtree starts at q+1, and the pair sweep run on all 32 lanes as before.
constexpr int DW_WARPS = 8;
constexpr int DW_ARRAYS = 10;    // int16 arrays of max_nodes entries per warp
// SMEM_MATRIX: the warp also keeps the tree's matrix in shared memory and writes it out in 16-byte pieces at the end.  The pair sweep
// runs in tour order but the matrix is indexed by taxon id, so written straight to global memory every store instruction puts 32
// two-byte values into scattered columns of one row (5-7 partial sectors each): at cfg2 that, not the arithmetic, was the kernel's
// time.  Only possible while the matrix is small (n <= ~110: 8 warps x 25 KB per SM) — and measured SLOWER there (0.47 against 0.27 ms at
// cfg2, profiles/r02_d_*): with 8 warps per SM instead of 48 the latency chains of the per-tree passes are no longer hidden.  Opt-in
// (QS_DIST_SMEM_MATRIX=1), kept for the record and for smaller matrices.
__host__ __device__ __forceinline__ size_t dist_warp_smem_per_warp(int max_nodes, int n, bool smem_matrix = false) {
    const size_t arrays = ((size_t)DW_ARRAYS * 2 * ((max_nodes + 7) / 8 * 8) + (size_t)((n + 31) / 32) * 4 + 15) & ~(size_t)15;
    return arrays + (smem_matrix ? (size_t)n * ((n + 7) / 8 * 8) * 2 : 0);
}

// 16-bit field of a shared int16 array, += v through the 32-bit word that holds it (fields stay < 32768: no carry); returns the old field
__device__ __forceinline__ int dw_add16(int16_t* arr, int idx, int v) {
    const uint32_t sh = (uint32_t)(idx & 1) * 16u;
    const uint32_t old = atomicAdd(reinterpret_cast<uint32_t*>(arr) + (idx >> 1), (uint32_t)v << sh);
    return (int)((old >> sh) & 0xffffu);
}
// words (ancestor << 16 | sum of the weights from the node up to, not including, that ancestor); the root is node 0 with
// weight 0.  On return every ancestor is 0 and the low halves hold the sums over the whole root paths.
__device__ __forceinline__ void dw_jump(volatile uint32_t* pj, int N, int lane) {
    bool again;
    do {
        again = false;
        for (int i = lane; i < N; i += 32) {
            const uint32_t w = pj[i];
            const uint32_t anc = w >> 16;
            if (anc != 0u) {
                const uint32_t wa = pj[anc];
                pj[i] = (wa & 0xffff0000u) | ((w + wa) & 0xffffu);
                again |= (wa >> 16) != 0u;
            }
        }
        __syncwarp();
        again = __any_sync(0xffffffffu, again);
    } while (again);
}

template <bool SMEM_MATRIX>
__global__ void __launch_bounds__(32 * DW_WARPS) qs_dist_warp_kernel(DistArgs a) {
    extern __shared__ __align__(16) unsigned char sm_w[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NP = (a.max_nodes + 7) / 8 * 8, nw = (a.n + 31) / 32;
    unsigned char* mine = sm_w + (size_t)warp * dist_warp_smem_per_warp(a.max_nodes, a.n, SMEM_MATRIX);
    __half* const Ms = reinterpret_cast<__half*>(mine + dist_warp_smem_per_warp(a.max_nodes, a.n, false));     // SMEM_MATRIX: this warp's n x n_pad matrix
    int16_t* par = reinterpret_cast<int16_t*>(mine);   // parent
    int16_t* dep = par + NP;                           // depth
    int16_t* cnt = dep + NP;                           // number of children
    int16_t* lfc = cnt + NP;                           // leaves in the subtree
    int16_t* off = lfc + NP;                           // tour position of the subtree's first leaf
    int16_t* nf = off + NP;                            // leaves of the children placed so far
    int16_t* lid = nf + NP;                            // taxon id (-1: none)
    int16_t* ltid = lid + NP;                          // per tour position: taxon id,
    int16_t* ldep = ltid + NP;                         //   depth,
    int16_t* lh = ldep + NP;                           //   depth of lca(leaf q, leaf q+1)
    uint32_t* seen = reinterpret_cast<uint32_t*>(lh + NP);
    volatile uint32_t* pj = reinterpret_cast<volatile uint32_t*>(ltid);   // pointer-jumping words, over ltid + ldep (filled after the jumps)
    const unsigned FULL = 0xffffffffu;
    int local_max = 0;
    const int gw = blockIdx.x * DW_WARPS + warp, stride = gridDim.x * DW_WARPS;

    for (int t = gw; t < a.m; t += stride) {
        const int64_t o = a.node_off[t];
        const int N = (int)(a.node_off[t + 1] - o);
        int bad = 0;
        __syncwarp();
        for (int i = lane; i < N; i += 32) {
            int p = a.parent[o + i];
            if (i > 0 && (p < 0 || p >= i)) { bad = 1; p = 0; }
            par[i] = (int16_t)p; cnt[i] = 0; lfc[i] = 0; nf[i] = 0;
            const int id = a.leaf_id[o + i];
            lid[i] = (int16_t)((id >= 0 && id < a.n) ? id : (id < 0 ? -1 : -2));
        }
        for (int i = lane; i < nw; i += 32) seen[i] = 0u;
        __syncwarp();
        // depths; child counts
        for (int i = lane; i < N; i += 32) {
            pj[i] = i > 0 ? (((uint32_t)par[i] << 16) | 1u) : 0u;
            if (i > 0) dw_add16(cnt, par[i], 1);
        }
        __syncwarp();
        dw_jump(pj, N, lane);
        int sumdep = 0;
        for (int i = lane; i < N; i += 32) {
            const int d = (int)(pj[i] & 0xffffu);
            dep[i] = (int16_t)d;
            if (cnt[i] == 0) sumdep += d;
        }
        for (int s = 16; s > 0; s >>= 1) sumdep += __shfl_xor_sync(FULL, sumdep, s);
        __syncwarp();
        // leaves per subtree
        if (sumdep <= 16 * N + 64) {
            for (int i = lane; i < N; i += 32)
                if (cnt[i] == 0) {
                    int p = i;
                    dw_add16(lfc, p, 1);
                    while (p != 0) { p = par[p]; dw_add16(lfc, p, 1); }
                }
        } else if (lane == 0) {
            for (int i = N - 1; i >= 1; --i) { const int l = cnt[i] == 0 ? 1 : lfc[i]; lfc[i] = (int16_t)l; lfc[par[i]] = lfc[par[i]] + l; }
            if (cnt[0] == 0) lfc[0] = 1;
        }
        __syncwarp();
        // tour position of the first leaf of every subtree: the leaves of the earlier siblings, summed along the root path
        for (int i = lane; i < N; i += 32) pj[i] = i > 0 ? (((uint32_t)par[i] << 16) | (uint32_t)dw_add16(nf, par[i], lfc[i])) : 0u;
        __syncwarp();
        dw_jump(pj, N, lane);
        for (int i = lane; i < N; i += 32) off[i] = (int16_t)(pj[i] & 0xffffu);
        __syncwarp();
        // validation, degrees, leaf arrays, turning depths: all lanes
        int maxdeg = 0;
        for (int i = lane; i < N; i += 32) {
            const int c = cnt[i], id = lid[i];
            maxdeg = max(maxdeg, c + (i > 0 ? 1 : 0));
            if (c == 0) {
                if (id < 0) bad = max(bad, 2);
                else {
                    const uint32_t bit = 1u << (id & 31);
                    if (atomicOr(&seen[id >> 5], bit) & bit) bad = max(bad, 3);
                    ltid[off[i]] = (int16_t)id; ldep[off[i]] = dep[i];
                }
            } else if (i > 0 && id != -1) bad = max(bad, 4);            // inner node carrying a taxon id
            if (i > 0 && off[i] > off[par[i]]) lh[off[i] - 1] = dep[par[i]];
        }
        for (int s = 16; s > 0; s >>= 1) { bad = max(bad, __shfl_xor_sync(FULL, bad, s)); maxdeg = max(maxdeg, __shfl_xor_sync(FULL, maxdeg, s)); }
        __syncwarp();
        const int k = bad ? 0 : (int)lfc[0];
        if (lane == 0) {
            if (bad) atomicCAS(a.max_dist + 1, 0, bad);
            a.tree_class[t] = (!bad && k == a.n && maxdeg <= 3) ? 0 : 1;
        }
        __half* const Dg = a.D + (size_t)t * a.n * a.n_pad;
        __half* const Dt = SMEM_MATRIX ? Ms : Dg;             // where the sweep writes
        const size_t nw16 = (size_t)a.n * a.n_pad / 8;
        if (k < a.n) {                                        // absent taxa / malformed tree: NaN everywhere first (see qs_dist_kernel)
            uint4* w = reinterpret_cast<uint4*>(Dt);
            for (size_t x = lane; x < nw16; x += 32) w[x] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
            __syncwarp();
        }
        for (int i0 = 0; i0 < k; i0 += 32) {
            const int i = i0 + lane;
            const bool act = i < k;
            const int ti = act ? ltid[i] : 0, di = act ? ldep[i] : 0;
            if (act) Dt[(size_t)ti * a.n_pad + ti] = __int2half_rn(0);
            int mn = 0x7fff;
            for (int j = i0 + 1; j < k; ++j) {                   // j ascending, lanes with i < j
                if (act && j > i) {
                    mn = min(mn, (int)lh[j - 1]);
                    const int d = di + ldep[j] - 2 * mn;
                    local_max = max(local_max, d);
                    Dt[(size_t)ltid[j] * a.n_pad + ti] = __int2half_rn(d);
                }
            }
            mn = 0x7fff;
            for (int j = min(k - 1, i0 + 31) - 1; j >= 0; --j) {  // j descending, lanes with i > j
                if (act && j < i) {
                    mn = min(mn, (int)lh[j]);
                    const int d = di + ldep[j] - 2 * mn;
                    Dt[(size_t)ltid[j] * a.n_pad + ti] = __int2half_rn(d);
                }
            }
        }
        if (SMEM_MATRIX) {                                    // the finished matrix, 16 bytes per lane and store (the padding columns carry whatever the previous tree left: never read for a result)
            __syncwarp();
            const uint4* src = reinterpret_cast<const uint4*>(Ms);
            uint4* dst = reinterpret_cast<uint4*>(Dg);
            for (size_t x = lane; x < nw16; x += 32) dst[x] = src[x];
        }
    }
    for (int s = 16; s > 0; s >>= 1) local_max = max(local_max, __shfl_xor_sync(FULL, local_max, s));
    if (lane == 0 && local_max > 0) atomicMax(a.max_dist, local_max);
}

// class-sorted tree order (stable partition: class-A trees first) and |A|; one CTA
__global__ void __launch_bounds__(1024) qs_order_kernel(const int32_t* __restrict__ tree_class, int m, int32_t* __restrict__ order, int32_t* __restrict__ n_class_a) {
    __shared__ int cntA[1024];
    __shared__ int s_totalA;
    const int tid = threadIdx.x;
    const int seg = (m + 1023) / 1024;
    const int lo = min(m, tid * seg), hi = min(m, lo + seg);
    int ca = 0;
    for (int t = lo; t < hi; ++t) ca += (tree_class[t] == 0);
    cntA[tid] = ca;
    __syncthreads();
    if (tid == 0) {
        int acc = 0;
        for (int i = 0; i < 1024; ++i) { int v = cntA[i]; cntA[i] = acc; acc += v; }
        s_totalA = acc;
        *n_class_a = acc;
    }
    __syncthreads();
    int pa = cntA[tid], pb = s_totalA + (lo - cntA[tid]);
    for (int t = lo; t < hi; ++t) {
        if (tree_class[t] == 0) order[pa++] = t; else order[pb++] = t;
    }
}

// fp16 matrix of one tree -> uint16 (0xFFFF for NaN): parity hook qs_get_distances
__global__ void qs_dist_to_u16_kernel(const __half* Dt, int n, int n_pad, uint16_t* out) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * n) return;
    int r = idx / n, c = idx % n;
    __half h = Dt[(size_t)r * n_pad + c];
    out[idx] = __hisnan(h) ? (uint16_t)0xFFFF : (uint16_t)__half2int_rn(h);
}

}  // namespace qs
