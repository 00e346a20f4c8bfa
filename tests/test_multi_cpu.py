"""N > 1 host logic on CPU (no GPU): the sharding rule, the host-only context and the partial-score
exchange of quartetscores_b200/multi.py with world_size 2 over gloo.

The per-shard partials the GPU kernels would produce are restated here in pure Python (small inputs) from
the golden count tables, following SURVEY.md App. A3 / kernels/score.cuh; after the all-reduce the library's
own host finaliser (qs_score_finalize on a QS_DEVICE_NONE context) must reproduce the reference's scores."""
import os
import socket
from math import comb

import numpy as np
import pytest

import _oracle as O
from _cases import load_input
from quartetscores_b200 import Context, QSError, QS_DEVICE_NONE
from quartetscores_b200.multi import allreduce_partials, finalize_on_host, shard_bounds

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("n", [4, 5, 9, 50, 100, 500, 1000, 2000])
@pytest.mark.parametrize("G", [1, 2, 3, 4, 8])
def test_shard_bounds_partition_the_rank_space(n, G):
    prev_end, prev_rank = 3, 0
    sizes = []
    for g in range(G):
        b, e, rb, re_ = shard_bounds(n, g, G)
        assert b == prev_end and rb == prev_rank and e >= b          # contiguous, no gap, no overlap
        assert rb == comb(b, 4) and re_ == comb(e, 4)                # QuartetLookupTable rank of the first quartet with s3 = b
        prev_end, prev_rank = e, re_
        sizes.append(re_ - rb)
    assert prev_end == n and prev_rank == comb(n, 4)
    if n >= 500:                                                      # balanced: boundaries ~ n (g/G)^(1/4)
        assert max(sizes) < 1.15 * (comb(n, 4) / G)


def test_host_only_context_has_no_compute(golden):
    g = golden("c1_known_answer")
    _, ref, flat = load_input(g)
    with Context(ref.n_taxa, 1, device=QS_DEVICE_NONE) as ctx:
        ctx.set_reference(ref)
        assert ctx.num_pairs() > 0
        with pytest.raises(QSError):
            ctx.add_trees(flat)                                       # there is no CPU fallback
        with pytest.raises(QSError):
            ctx.count()
        with pytest.raises(QSError):
            ctx.score(1)


def _ref_tables(ref):
    """inner index, depth and leaf-pair LCA tables with the library's conventions (qscuda.cu build_reference)."""
    N = ref.n_nodes
    parent = [int(x) for x in ref.parent]
    depth = [0] * N
    for i in range(1, N):
        depth[i] = depth[parent[i]] + 1
    is_inner = [ref.leaf_lookup_id[i] < 0 for i in range(N)]
    with Context(ref.n_taxa, 2, device=QS_DEVICE_NONE) as host:          # the library's order of inner nodes (include/qscuda.h)
        host.set_reference(ref)
        inner_nodes = [int(v) for v in host.inner_nodes()]
    assert sorted(inner_nodes) == [i for i in range(N) if is_inner[i]]
    inner_index = {v: k for k, v in enumerate(inner_nodes)}
    leaf_node = {int(ref.leaf_lookup_id[i]): i for i in range(N) if not is_inner[i]}

    def lca(x, y):
        while x != y:
            if depth[x] >= depth[y]:
                x = parent[x]
            else:
                y = parent[y]
        return x

    n = ref.n_taxa
    L = [[inner_index[lca(leaf_node[x], leaf_node[y])] if x != y else -1 for y in range(n)] for x in range(n)]
    idepth = [depth[v] for v in inner_nodes]
    nchild = [0] * N
    for i in range(1, N):
        nchild[parent[i]] += 1
    bif = max(nchild[v] + (1 if v else 0) for v in inner_nodes) == 3
    return parent, depth, inner_nodes, L, idepth, bif


def shard_partials(ref, table, g, G):
    """What qs_score_partials returns for shard g of G (count_scale 1): (lqic_partial[E], pair_sums[I*I*3])."""
    parent, depth, inner_nodes, L, idepth, bif = _ref_tables(ref)
    n, I, E = ref.n_taxa, len(inner_nodes), ref.edge_count
    s3b, s3e, _, _ = shard_bounds(n, g, G)
    lq = np.full(E, np.inf)
    sums = np.zeros(I * I * 3, np.uint64)
    best = {}
    for d in range(max(3, s3b), s3e):
        for c in range(2, d):
            for b in range(1, c):
                q, r = L[b][c], L[c][d]
                for a in range(b):
                    p = L[a][b]
                    dp, dq, dr = idepth[p], idepth[q], idepth[r]
                    S0, S2 = dp + dr, min(dp, dq, dr) + dq
                    if S0 == S2:
                        continue
                    c0, c1, c2 = (int(x) for x in table[O.rank(a, b, c, d)])
                    if S0 > S2:
                        u, v = (p if dp > dq else q), (r if dr > dq else q)
                        t = (c0, c1, c2)
                    else:
                        u, v = q, (p if dp > dr else r)
                        t = (c2, c1, c0) if bif else (c2, c0, c1)
                    key = (min(u, v), max(u, v))
                    k = (key[0] * I + key[1]) * 3
                    sums[k] += np.uint64(t[0]); sums[k + 1] += np.uint64(t[1]); sums[k + 2] += np.uint64(t[2])
                    qic = O.log_score(*t)
                    if key not in best or qic < best[key]:
                        best[key] = qic
    for (iu, iv), qic in best.items():
        x, y = inner_nodes[iu], inner_nodes[iv]
        while x != y:
            if depth[x] >= depth[y]:
                e = int(ref.parent_edge[x]); x = parent[x]
            else:
                e = int(ref.parent_edge[y]); y = parent[y]
            lq[e] = min(lq[e], qic)
    return lq, sums


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, case, out_q):
    import torch.distributed as dist

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    z = np.load(os.path.join(GOLDEN, case + ".npz"))
    g = {k: (z[k].item() if z[k].ndim == 0 else z[k]) for k in z.files}
    _, ref, _ = load_input(g)
    lq, sums = shard_partials(ref, g["counts"], rank, world)
    lq_r, sums_r = allreduce_partials(lq, sums)                       # MIN / SUM over gloo
    res = finalize_on_host(ref, ref.n_taxa, lq_r, sums_r)
    out_q.put((rank, [np.asarray(x) for x in res], int(sums.sum())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("case", ["s13x257_poly", "s16x300_missing_poly", "s20x40_multiref_missing"])
def test_two_rank_gloo_exchange_reproduces_reference_scores(case, golden):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, case, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = golden(case)
    local_sums = [r[2] for r in sorted(results)]
    assert all(s > 0 for s in local_sums), "both shards must own quartets"      # a real 2-way split
    for _, (lq, qp, eqp), _ in results:                                            # every rank ends with the full answer
        for got, key in ((lq, "lqic"), (qp, "qpic"), (eqp, "eqpic")):
            want = g[key]
            assert np.array_equal(np.isinf(got), np.isinf(want)), key
            fin = np.isfinite(want)
            assert np.allclose(got[fin], want[fin], rtol=0, atol=1e-9), key


def test_host_finalize_is_thread_count_independent(monkeypatch):
    """qs_score_finalize on a 300-taxon reference (host-only context): the threaded pair post-pass must give the
    single-thread result bit for bit."""
    from quartetscores_b200.newick import flatten_reference, parse_newick
    from quartetscores_b200.synth import SyntheticInput

    s = SyntheticInput(300, 1, 77, k_max=0, want_newick=False)
    ref = flatten_reference(parse_newick(s.ref_newick))
    rng = np.random.default_rng(5)
    out = {}
    for th in ("1", "7"):
        monkeypatch.setenv("QS_HOST_THREADS", th)
        with Context(ref.n_taxa, 2, device=QS_DEVICE_NONE) as ctx:
            ctx.set_reference(ref)
            P = ctx.num_pairs()
            if "sums" not in out:
                out["sums"] = rng.integers(0, 1 << 34, size=P * 3, dtype=np.uint64)
                out["lq"] = rng.standard_normal(ref.edge_count)
            out[th] = ctx.score_finalize(out["lq"], out["sums"])
    for a, b in zip(out["1"], out["7"]):
        assert np.array_equal(a, b)
    assert np.isfinite(out["1"][2]).sum() > 250          # EQP-IC reaches every internal edge
