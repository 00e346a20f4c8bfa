"""GPU parity: the CUDA path (through the C ABI) against the reference-pinned golden fixtures and the
CPU oracle.  Run on the B200 box with `pytest -m gpu`."""
import os

import numpy as np
import pytest

import _oracle as O
from _cases import cint_bits_for, load_input
from conftest import golden_cases
from quartetscores_b200 import Context, QuartetScoreComputer
from quartetscores_b200.newick import write_annotated_newick
from quartetscores_b200.synth import SyntheticInput
from quartetscores_b200.newick import flatten_reference, parse_newick

pytestmark = pytest.mark.gpu


def run_ctx(ref, flat, **kw):
    bits = cint_bits_for(flat.n_trees)
    ctx = Context(ref.n_taxa, bits // 8, **kw)
    ctx.set_reference(ref)
    ctx.add_trees(flat)
    ctx.count()
    return ctx


@pytest.mark.parametrize("name", golden_cases())
def test_golden_counts_bit_exact(name, golden):
    g = golden(name)
    _, ref, flat = load_input(g)
    with run_ctx(ref, flat) as ctx:
        got = ctx.get_counts()
        assert got.dtype.itemsize * 8 == cint_bits_for(flat.n_trees)
        assert np.array_equal(got.astype(np.uint32), g["counts"].astype(np.uint32))


@pytest.mark.parametrize("name", golden_cases())
def test_golden_distances(name, golden):
    g = golden(name)
    _, ref, flat = load_input(g)
    with run_ctx(ref, flat) as ctx:
        for t in range(min(flat.n_trees, 5)):
            o = flat.node_offsets
            want, _ = O.distance_matrix(flat.parent[o[t]:o[t + 1]], flat.leaf_lookup_id[o[t]:o[t + 1]], ref.n_taxa)
            assert np.array_equal(ctx.get_distances(t), want), f"tree {t}"


@pytest.mark.parametrize("name", golden_cases())
def test_golden_scores_and_output_text(name, golden):
    g = golden(name)
    ref_root, ref, flat = load_input(g)
    with run_ctx(ref, flat) as ctx:
        for suffix, scale in (("", 1), ("_s", 2)):
            lq, qp, eqp = ctx.score(scale)
            # tolerance of the task: 1e-9 absolute per internode score; in practice identical bits
            for got, key in ((lq, "lqic"), (qp, "qpic"), (eqp, "eqpic")):
                want = g[key + suffix]
                assert np.array_equal(np.isinf(got), np.isinf(want)), key + suffix
                fin = np.isfinite(want)
                assert np.allclose(got[fin], want[fin], rtol=0, atol=1e-9), key + suffix
            bif = bool(np.isfinite(g["qpic"]).any())
            text = write_annotated_newick(ref_root, ref, lq, qp if bif else None, eqp if bif else None)
            assert text == g["out_newick" + suffix]


@pytest.mark.parametrize("name", golden_cases())
def test_golden_raw_qic_file(name, golden, tmp_path):
    g = golden(name)
    _, ref, flat = load_input(g)
    with run_ctx(ref, flat) as ctx:
        p = str(tmp_path / "raw.txt")
        ctx.write_raw_qic(ref.taxa, p)
        assert open(p).read() == g["rawqic"]


@pytest.mark.parametrize("n,m,seed,kw", [
    (40, 300, 11, dict(k_max=10, p_missing=0.1, p_contract=0.1)),
    (50, 1000, 12, dict(k_max=10)),                       # BASELINE config 1 shape
    (64, 257, 13, dict(k_max=30, nni_fraction=0.1)),
    (9, 2500, 14, dict(k_max=3, p_missing=0.3)),           # more than one tree chunk per task
    (33, 100, 15, dict(k_max=5, p_contract=0.5)),
])
def test_counts_vs_oracle_seeded(n, m, seed, kw):
    s = SyntheticInput(n, m, seed, want_newick=False, **kw)
    ref = flatten_reference(parse_newick(s.ref_newick))
    want = O.count_clades_compact(n, s.flat) // 2           # reference enumeration (doubled table) halved
    with run_ctx(ref, s.flat) as ctx:
        got = ctx.get_counts().astype(np.uint32)
    assert np.array_equal(got, want)


def test_scores_vs_oracle_seeded():
    s = SyntheticInput(36, 400, 21, k_max=12, p_missing=0.05, want_newick=False)
    ref = flatten_reference(parse_newick(s.ref_newick))
    with run_ctx(ref, s.flat) as ctx:
        table = ctx.get_counts().astype(np.uint32)
        lq, qp, eqp = ctx.score(1)
    wl, wq, we, bif = O.score(ref, table, 1, 16)
    assert bif
    for got, want in ((lq, wl), (qp, wq), (eqp, we)):
        assert np.array_equal(np.isinf(got), np.isinf(want))
        fin = np.isfinite(want)
        assert np.allclose(got[fin], want[fin], rtol=0, atol=1e-9)


def test_mirror_class_known_answer(golden):
    g = golden("c1_known_answer")
    qsc = QuartetScoreComputer(g["ref_newick"], g["eval_newick"])
    text = write_annotated_newick(qsc.ref_root, qsc.ref, qsc.getLQICScores(), qsc.getQPICScores(), qsc.getEQPICScores())
    assert text == g["out_newick"]
    # ABCD 2 1 0, and any argument order gives the matching pairing
    assert qsc.countQuartetOccurrences(0, 1, 2, 3) == (2, 1, 0)
    assert qsc.countQuartetOccurrences(0, 2, 1, 3) == (1, 2, 0)
    assert qsc.countQuartetOccurrences(3, 0, 2, 1) == (0, 2, 1)   # 30|21 = slot 2, 32|01 = slot 0, 31|02 = slot 1


def test_unknown_taxon_rejected(golden):
    g = golden("c1_known_answer")
    with pytest.raises(IndexError):
        QuartetScoreComputer(g["ref_newick"], "((A,B),(C,Z));")


def test_bad_tree_encoding_rejected(golden):
    g = golden("c1_known_answer")
    _, ref, flat = load_input(g)
    from quartetscores_b200 import QSError
    bad = flat.leaf_lookup_id.copy()
    leaves = np.nonzero(bad >= 0)[0]
    bad[leaves[1]] = bad[leaves[0]]          # duplicate taxon inside the first tree
    ctx = Context(ref.n_taxa, 1)
    ctx.set_reference(ref)
    ctx.add_trees_raw(flat.node_offsets, flat.parent, bad)
    with pytest.raises(QSError):
        ctx.count()
    ctx.close()


# ---- more shapes for the counting kernel, and table-free mode ---------------------------------------------

@pytest.mark.parametrize("n,m,seed,kw", [
    (70, 120, 32, dict(k_max=15)),                          # several variable rows per task, class A only
    (19, 4300, 33, dict(k_max=4, p_missing=0.2)),          # > 4096 trees: more than one counter chunk (QS_MAX_CHUNK_TREES), class B
    (21, 4500, 34, dict(k_max=4)),                          # > 4096 class-A trees
    (130, 40, 35, dict(k_max=25, p_missing=0.05, p_contract=0.05)),   # wider rows, tasks that split a row's items
])
def test_counts_vs_oracle_more_shapes(n, m, seed, kw):
    s = SyntheticInput(n, m, seed, want_newick=False, **kw)
    ref = flatten_reference(parse_newick(s.ref_newick))
    want = O.count_clades_compact(n, s.flat) // 2
    with run_ctx(ref, s.flat) as ctx:
        assert np.array_equal(ctx.get_counts().astype(np.uint32), want)


@pytest.fixture
def tiny_slabs(monkeypatch):
    monkeypatch.setenv("QS_SLAB_BYTES", "20000")          # table-free mode: force many slabs of the d-range


@pytest.mark.parametrize("name", golden_cases())
@pytest.mark.parametrize("scale,suffix", [(1, ""), (2, "_s")])
def test_table_free_mode_golden(name, golden, scale, suffix, tiny_slabs):
    """-s analogue: no resident table; the d-range is counted and scored slab by slab."""
    from quartetscores_b200 import QS_MODE_TABLE_FREE
    g = golden(name)
    ref_root, ref, flat = load_input(g)
    ctx = Context(ref.n_taxa, cint_bits_for(flat.n_trees) // 8, mode=QS_MODE_TABLE_FREE)
    ctx.set_reference(ref)
    ctx.set_count_scale(scale)
    ctx.add_trees(flat)
    ctx.count()
    lq, qp, eqp = ctx.score(scale)
    ctx.close()
    bif = bool(np.isfinite(g["qpic"]).any())
    assert write_annotated_newick(ref_root, ref, lq, qp if bif else None, eqp if bif else None) == g["out_newick" + suffix]
    for got, key in ((lq, "lqic"), (qp, "qpic"), (eqp, "eqpic")):
        want = g[key + suffix]
        fin = np.isfinite(want)
        assert np.array_equal(np.isinf(got), np.isinf(want)) and np.allclose(got[fin], want[fin], rtol=0, atol=1e-9)


# ---- multi-GPU path: shards emulated sequentially on one GPU (integer sums and exact mins are
# order-independent, so G shards must reproduce the 1-shard result bit for bit) ---------------------------

def _run_shards(ref, flat, G, scale=1, mode=None):
    from quartetscores_b200 import QS_MODE_TABLE
    bits = cint_bits_for(flat.n_trees)
    tables, lq_parts, sum_parts, ranges = [], [], [], []
    fin = None
    for g in range(G):
        ctx = Context(ref.n_taxa, bits // 8, mode=QS_MODE_TABLE if mode is None else mode, shard_index=g, shard_count=G)
        ctx.set_reference(ref)
        if mode is not None:
            ctx.set_count_scale(scale)
        ctx.add_trees(flat)
        ctx.count()
        r0, r1 = ctx.shard_range()
        ranges.append((r0, r1))
        if mode is None:
            tables.append(ctx.get_counts(r0, r1))
        lq, sums = ctx.score_partials(scale)
        lq_parts.append(lq)
        sum_parts.append(sums)
        if g == G - 1:
            fin = ctx.score_finalize(np.minimum.reduce(lq_parts), np.add.reduce(sum_parts))
        ctx.close()
    return tables, ranges, fin


@pytest.mark.parametrize("G", [2, 3, 8])
def test_shards_reproduce_single_shard(G):
    s = SyntheticInput(30, 300, 41, k_max=10, p_missing=0.05, p_contract=0.05, want_newick=False)
    ref = flatten_reference(parse_newick(s.ref_newick))
    with run_ctx(ref, s.flat) as ctx:
        full = ctx.get_counts()
        want = ctx.score(1)
    tables, ranges, got = _run_shards(ref, s.flat, G)
    assert ranges[0][0] == 0 and ranges[-1][1] == len(full) and all(ranges[i][1] == ranges[i + 1][0] for i in range(G - 1))
    assert np.array_equal(np.concatenate(tables), full)
    for a, b in zip(got, want):
        assert np.array_equal(a, b)                       # bit-identical, not just within tolerance


@pytest.mark.parametrize("G", [2, 4])
def test_table_free_shards_reproduce_single_shard(G, tiny_slabs):
    from quartetscores_b200 import QS_MODE_TABLE_FREE
    s = SyntheticInput(28, 200, 42, k_max=8, want_newick=False)
    ref = flatten_reference(parse_newick(s.ref_newick))
    with run_ctx(ref, s.flat) as ctx:
        want = ctx.score(1)
    _, _, got = _run_shards(ref, s.flat, G, mode=QS_MODE_TABLE_FREE)
    for a, b in zip(got, want):
        assert np.array_equal(a, b)


def test_mixed_tree_classes_and_odd_counts():
    """class A (complete, fully resolved) and class B trees interleaved, odd class sizes -> padded tree pairs."""
    a = SyntheticInput(26, 151, 51, k_max=6, want_newick=False)                                   # all class A
    b = SyntheticInput(26, 90, 51, k_max=6, p_missing=0.15, p_contract=0.1, want_newick=False)    # same reference, class B
    assert a.ref_newick == b.ref_newick
    ref = flatten_reference(parse_newick(a.ref_newick))
    ctx = Context(ref.n_taxa, 1)
    ctx.set_reference(ref)
    # interleave: B-chunk, A, B-chunk
    fb = b.flat
    half = 45
    ob = fb.node_offsets
    ctx.add_trees_raw(ob[:half + 1], fb.parent[:ob[half]], fb.leaf_lookup_id[:ob[half]])
    ctx.add_trees(a.flat)
    ctx.add_trees_raw(ob[half:] - ob[half], fb.parent[ob[half]:], fb.leaf_lookup_id[ob[half]:])
    ctx.count()
    nA, nB = ctx.tree_classes()
    assert nA >= 151 and nA + nB == 241 and nB > 0
    got = ctx.get_counts().astype(np.uint32)
    ctx.close()
    want = (O.count_clades_compact(26, a.flat) + O.count_clades_compact(26, b.flat)) // 2
    assert np.array_equal(got, want)


# ---- native ingest and table persistence -------------------------------------------------

@pytest.mark.parametrize("name", golden_cases())
def test_golden_counts_through_native_newick_ingest(name, golden, tmp_path):
    """Newick text -> qs_add_newick(_file) -> counts: the reference's golden table, without the Python parser."""
    g = golden(name)
    _, ref, flat = load_input(g)
    bits = cint_bits_for(flat.n_trees)
    p = tmp_path / "eval.nwk"
    p.write_text(g["eval_newick"])
    for use_file in (False, True):
        with Context(ref.n_taxa, bits // 8) as ctx:
            ctx.set_reference(ref)
            got = ctx.add_newick_file(str(p), ref.taxa, 3) if use_file else ctx.add_newick(g["eval_newick"], ref.taxa, 2)
            assert got == flat.n_trees == ctx.num_trees()
            ctx.count()
            assert np.array_equal(ctx.get_counts().astype(np.uint32), g["counts"].astype(np.uint32))


def test_table_save_load_roundtrip(tmp_path):
    s = SyntheticInput(30, 400, seed=21, k_max=8, p_missing=0.1, p_contract=0.1, want_newick=False)
    ref = flatten_reference(parse_newick(s.ref_newick))
    path = str(tmp_path / "table.qstbl")
    with run_ctx(ref, s.flat) as ctx:
        table = ctx.get_counts()
        scores = ctx.score(1)
        ctx.save_table(path)
    assert os.path.getsize(path) == 64 + table.nbytes
    with Context(ref.n_taxa, 2) as ctx:                    # a fresh context: no trees, no counting
        ctx.set_reference(ref)
        ctx.load_table(path)
        assert np.array_equal(ctx.get_counts(), table)
        for a, b in zip(ctx.score(1), scores):
            assert np.array_equal(a, b)
    with Context(ref.n_taxa, 1) as ctx:                    # wrong counter width is refused
        with pytest.raises(Exception):
            ctx.load_table(path)


@pytest.mark.parametrize("threads", ["1", "5"])
def test_scores_vs_oracle_wide_reference(threads, monkeypatch):
    """120 taxa: rows long enough for the vectorised table scan, host post-pass forced onto several threads."""
    monkeypatch.setenv("QS_HOST_THREADS", threads)
    s = SyntheticInput(120, 300, 61, k_max=15, p_missing=0.05, p_contract=0.03, want_newick=False)
    ref = flatten_reference(parse_newick(s.ref_newick))
    with run_ctx(ref, s.flat) as ctx:
        table = ctx.get_counts().astype(np.uint32)
        got = {scale: ctx.score(scale) for scale in (1, 2)}
    for scale in (1, 2):
        wl, wq, we, bif = O.score(ref, table, scale, 16)
        assert bif
        for g, want in zip(got[scale], (wl, wq, we)):
            assert np.array_equal(np.isinf(g), np.isinf(want))
            fin = np.isfinite(want)
            assert np.allclose(g[fin], want[fin], rtol=0, atol=1e-9)


@pytest.mark.parametrize("threads", ["512", "256"])
@pytest.mark.parametrize("name", ["s32x270_spr", "s16x300_missing_poly", "s24x60_u8"])
def test_golden_counts_both_cta_shapes(name, threads, golden, monkeypatch):
    """the counting kernel has two CTA shapes (512 x 1 for n < 150, 256 x 2 above): both must give the reference's table"""
    monkeypatch.setenv("QS_CR_THREADS", threads)
    g = golden(name)
    _, ref, flat = load_input(g)
    with run_ctx(ref, flat) as ctx:
        assert np.array_equal(ctx.get_counts().astype(np.uint32), g["counts"].astype(np.uint32))


def test_counts_vs_oracle_small_cta_shape_default():
    """160 taxa selects the 256 x 2 shape by itself (n > 112); class-B trees so that role Y runs too.  Few trees: the oracle is O(n^4 m)."""
    s = SyntheticInput(160, 4, 71, k_max=25, p_missing=0.08, p_contract=0.05, want_newick=False)
    ref = flatten_reference(parse_newick(s.ref_newick))
    want = O.count_clades_compact(160, s.flat) // 2
    with run_ctx(ref, s.flat) as ctx:
        got = ctx.get_counts().astype(np.uint32)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("cint_bytes", [4, 8])
def test_wide_counter_types(cint_bytes, golden):
    """uint32 / uint64 CINT (src/QuartetScores.cpp:115-147 picks them for m >= 65,536 / 2^32 trees): same counts, same scores."""
    g = golden("s16x300_missing_poly")
    _, ref, flat = load_input(g)
    with Context(ref.n_taxa, cint_bytes) as ctx:
        ctx.set_reference(ref)
        ctx.add_trees(flat)
        ctx.count()
        got = ctx.get_counts()
        assert got.dtype.itemsize == cint_bytes
        assert np.array_equal(got.astype(np.uint64), g["counts"].astype(np.uint64))
        lq, qp, eqp = ctx.score(1)
    for a, key in ((lq, "lqic"), (qp, "qpic"), (eqp, "eqpic")):
        want = g[key]
        assert np.array_equal(np.isinf(a), np.isinf(want))
        fin = np.isfinite(want)
        assert np.allclose(a[fin], want[fin], rtol=0, atol=1e-9)


@pytest.mark.parametrize("threads", ["1", "6"])
def test_raw_qic_file_thread_count_independent(threads, golden, tmp_path, monkeypatch):
    monkeypatch.setenv("QS_HOST_THREADS", threads)
    g = golden("s32x270_spr")
    _, ref, flat = load_input(g)
    with run_ctx(ref, flat) as ctx:
        p = str(tmp_path / "raw.txt")
        ctx.write_raw_qic(ref.taxa, p)
        assert open(p).read() == g["rawqic"]


# ---- round 2: scan kernel variants, device-resident multi-shard reduction, documented limits ---------------------------

@pytest.mark.parametrize("name", ["s32x270_spr", "s16x300_missing_poly", "s20x40_multiref_missing", "c2_multifurcating_ref"])
def test_scan_with_global_accumulators(name, golden, monkeypatch):
    """the scan keeps its per-CTA accumulators in global memory when they do not fit shared memory (n > ~2,300): forced here"""
    monkeypatch.setenv("QS_SCAN_GLOBAL_ACC", "1")
    g = golden(name)
    _, ref, flat = load_input(g)
    with run_ctx(ref, flat) as ctx:
        for suffix, scale in (("", 1), ("_s", 2)):
            for got, key in zip(ctx.score(scale), ("lqic", "qpic", "eqpic")):
                want = g[key + suffix]
                assert np.array_equal(np.isinf(got), np.isinf(want)), key + suffix
                fin = np.isfinite(want)
                assert np.allclose(got[fin], want[fin], rtol=0, atol=1e-9), key + suffix


@pytest.mark.parametrize("G", [2, 5])
@pytest.mark.parametrize("mode_name", ["table", "table_free"])
def test_device_resident_shard_reduction(G, mode_name, tiny_slabs):
    """The N-GPU score path with the partials left on the device (qs_score_scan / qs_score_device_partials /
    qs_score_select_winners / qs_score_finish): G shard contexts on ONE GPU, the three all-reduces replaced by the same
    element-wise SUM / MIN / MIN over the shards' device buffers -> the 1-shard scores, bit for bit."""
    import torch
    from quartetscores_b200 import QS_MODE_TABLE, QS_MODE_TABLE_FREE
    from quartetscores_b200.multi import _DeviceArray
    s = SyntheticInput(34, 320, 43, k_max=10, p_missing=0.08, p_contract=0.05, want_newick=False)
    ref = flatten_reference(parse_newick(s.ref_newick))
    with run_ctx(ref, s.flat) as ctx:
        want = ctx.score(1)
    mode = QS_MODE_TABLE if mode_name == "table" else QS_MODE_TABLE_FREE
    stream = torch.cuda.current_stream()
    ctxs = []
    for g in range(G):
        c = Context(ref.n_taxa, 2, mode=mode, shard_index=g, shard_count=G)
        c.set_stream(stream.cuda_stream)
        c.set_reference(ref)
        c.add_trees(s.flat)
        c.count()
        c.score_scan(1)
        ctxs.append(c)
    dev = torch.device("cuda", 0)
    views = []
    for c in ctxs:
        ps, sc, be, npairs = c.score_device_partials()
        views.append((torch.as_tensor(_DeviceArray(ps, 3 * npairs), device=dev), torch.as_tensor(_DeviceArray(sc, npairs), device=dev),
                      torch.as_tensor(_DeviceArray(be, npairs), device=dev)))
    sums = torch.stack([v[0] for v in views]).sum(0)
    score = torch.stack([v[1] for v in views]).min(0).values
    for v in views:                                   # "all-reduce" SUM and MIN in place
        v[0].copy_(sums)
        v[1].copy_(score)
    for c in ctxs:
        c.score_select_winners()
    best = torch.stack([v[2] for v in views]).min(0).values
    for v in views:
        v[2].copy_(best)
    for c in ctxs:                                    # every rank finishes on its own and must get the same result
        got = c.score_finish()
        for a, b in zip(got, want):
            assert np.array_equal(a, b)
        c.close()


def test_lq_selection_count_limit_is_refused_not_wrapped():
    """m * count_scale >= 2^21 does not fit the packed triple of the LQ-IC selection: QS_E_UNSUPPORTED instead of a silently wrong
    LQ-IC (the reference's log_score takes size_t, QuartetScoreComputer.hpp:135).  2^20 copies of one 5-taxon tree, uint32 CINT."""
    from quartetscores_b200 import QSError
    ref = flatten_reference(parse_newick("((A,B),(C,D),E);"))
    m = 1 << 20
    par = np.array([-1, 0, 1, 1, 0, 4, 4, 0], np.int32)               # ((A,B),(C,D),E) in pre-order
    leaf = np.array([-1, -1, 0, 1, -1, 2, 3, 4], np.int32)
    off = np.arange(m + 1, dtype=np.int64) * 8
    with Context(5, 4) as ctx:
        ctx.set_reference(ref)
        ctx.add_trees_raw(off, np.tile(par, m), np.tile(leaf, m))
        ctx.count()
        assert ctx.get_counts().tolist() == [[m, 0, 0], [m, 0, 0], [m, 0, 0], [0, 0, m], [0, 0, m]]      # ABCD ABCE ABDE ACDE BCDE
        lq, qp, eqp = ctx.score(1)                                     # m * 1 < 2^21: fine, every quartet agrees with the reference tree
        assert np.all(lq[np.isfinite(lq)] == 1.0) and np.all(qp[np.isfinite(qp)] == 1.0)
        with pytest.raises(QSError) as e:
            ctx.score(2)                                               # m * 2 = 2^21
        assert e.value.code == -6


def test_auto_mode_and_table_free_counts_export(golden, monkeypatch, tmp_path):
    """QS_MODE_AUTO picks the resident table when it fits and table-free slabs when it does not (forced); a table-free context
    still answers qs_get_counts / writes the -q file by counting the covering slabs again."""
    from quartetscores_b200 import QS_MODE_AUTO, QS_MODE_TABLE_FREE
    g = golden("s32x270_spr")
    _, ref, flat = load_input(g)
    want = g["counts"].astype(np.uint32)
    monkeypatch.setenv("QS_SLAB_BYTES", "30000")
    for mode, limit in ((QS_MODE_AUTO, None), (QS_MODE_AUTO, "1000"), (QS_MODE_TABLE_FREE, None)):
        if limit:
            monkeypatch.setenv("QS_TABLE_BYTES_LIMIT", limit)
        with Context(ref.n_taxa, 2, mode=mode) as ctx:
            ctx.set_reference(ref)
            ctx.add_trees(flat)
            ctx.count()
            assert np.array_equal(ctx.get_counts().astype(np.uint32), want)
            assert np.array_equal(ctx.get_counts(1000, 1700).astype(np.uint32), want[1000:1700])
            lq, qp, eqp = ctx.score(1)
            assert np.allclose(lq[np.isfinite(lq)], g["lqic"][np.isfinite(g["lqic"])], rtol=0, atol=1e-9)
            p = str(tmp_path / f"raw_{mode}_{limit}.txt")
            ctx.write_raw_qic(ref.taxa, p)
            assert open(p).read() == g["rawqic"]


def test_raw_qic_from_several_shards(golden, tmp_path):
    g = golden("s16x300_missing_poly")
    _, ref, flat = load_input(g)
    ctxs = []
    for k in range(3):
        c = Context(ref.n_taxa, 2, shard_index=k, shard_count=3)
        c.set_reference(ref)
        c.add_trees(flat)
        c.count()
        ctxs.append(c)
    p = str(tmp_path / "raw.txt")
    Context.write_raw_qic_shards(ctxs, ref.taxa, p)
    assert open(p).read() == g["rawqic"]
    from quartetscores_b200 import QSError
    with pytest.raises(QSError):
        ctxs[1].write_raw_qic(ref.taxa, p)                 # one shard alone cannot write the file
    for c in ctxs:
        c.close()


@pytest.mark.parametrize("kw", [dict(k_max=10), dict(k_max=10, p_missing=0.1, p_contract=0.1)])
def test_rebalanced_shards_tile_the_rank_space_and_reproduce_single_shard(kw):
    """qs_rebalance_shards re-cuts the ranges for the observed class mix (class A trees skip role Y): the ranges must still tile
    the rank space, every shard must arrive at the same cut without talking to the others, and tables / scores must not change"""
    s = SyntheticInput(60, 300, 44, want_newick=False, **kw)
    ref = flatten_reference(parse_newick(s.ref_newick))
    with run_ctx(ref, s.flat) as ctx:
        full, want = ctx.get_counts(), ctx.score(1)
    G, tables, lqs, sums, ranges, fin = 5, [], [], [], [], None
    for g in range(G):
        ctx = Context(ref.n_taxa, 2, shard_index=g, shard_count=G)
        ctx.set_reference(ref)
        ctx.add_trees(s.flat)
        before = ctx.shard_range()
        moved = ctx.rebalance_shards()
        assert moved == (ctx.shard_range() != before)
        ctx.count()
        r0, r1 = ctx.shard_range()
        ranges.append((r0, r1))
        tables.append(ctx.get_counts(r0, r1))
        lq, sm = ctx.score_partials(1)
        lqs.append(lq); sums.append(sm)
        if g == G - 1:
            fin = ctx.score_finalize(np.minimum.reduce(lqs), np.add.reduce(sums))
        ctx.close()
    assert ranges[0][0] == 0 and ranges[-1][1] == len(full) and all(ranges[i][1] == ranges[i + 1][0] for i in range(G - 1))
    assert np.array_equal(np.concatenate(tables), full)
    for a, b in zip(fin, want):
        assert np.array_equal(a, b)
