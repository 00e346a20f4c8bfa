"""Host-side mirror of the reference's operator interface for the hot path.

``QuartetScoreComputer`` keeps the constructor arguments, method names and error behaviour of the
reference class (src/QuartetScoreComputer.hpp:43-51, :698-785): it is built from a reference tree and the
evaluation trees, does all the work in its constructor, and then hands out
``getLQICScores / getQPICScores / getEQPICScores`` (vectors indexed by genesis edge index, +inf where
untouched; QP/EQP empty for a multifurcating reference) and ``printRawQICScores``.
``countQuartetOccurrences`` mirrors QuartetCounterLookup::countQuartetOccurrences
(src/QuartetCounterLookup.hpp:300-318).  All computation goes through the C ABI of libqscuda.so.
"""
from __future__ import annotations

import ctypes as C
from math import comb
from typing import List, Optional, Sequence, Union

import numpy as np

from . import _ffi
from .newick import FlatReference, FlatTrees, Node, flatten_eval_trees, flatten_reference, parse_newick, parse_newick_many


def cint_bytes_for(m: int) -> int:
    """CINT width selection of the reference main (src/QuartetScores.cpp:115-147)."""
    return 1 if m < (1 << 8) else 2 if m < (1 << 16) else 4 if m < (1 << 32) else 8


_NP_CINT = {1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}


def _ptr(a: np.ndarray, t):
    return a.ctypes.data_as(C.POINTER(t))


class Context:
    """Thin RAII wrapper over qs_ctx (one CUDA device, one shard of the quartet rank space)."""

    def __init__(self, n_taxa: int, cint_bytes: int = 2, mode: int = _ffi.QS_MODE_TABLE, device: int = 0,
                 shard_index: int = 0, shard_count: int = 1):
        self._lib = _ffi.load()
        self._h = C.c_void_p()
        rc = self._lib.qs_create(C.byref(self._h), n_taxa, cint_bytes, mode, device, shard_index, shard_count)
        if rc != 0:
            raise _ffi.QSError(rc, "qs_create: " + self._lib.qs_strerror(rc).decode())
        self.n_taxa, self.cint_bytes, self.mode = n_taxa, cint_bytes, mode
        self.shard_index, self.shard_count, self.device = shard_index, shard_count, device
        self.edge_count = 0
        self.bifurcating_hint = None

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise _ffi.QSError(rc, f"{what}: {self._lib.qs_last_error(self._h).decode()}")

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.qs_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def set_stream(self, cuda_stream_handle: int):
        """Run this context's work on the given CUDA stream.  torch's default stream has handle 0, which the C ABI reads as
        "the context's own stream": it is passed as cudaStreamLegacy (1) instead, so that torch ops and NCCL collectives issued
        on torch's current stream are ordered with the library's kernels."""
        self._check(self._lib.qs_set_stream(self._h, C.c_void_p(cuda_stream_handle if cuda_stream_handle else 1)), "qs_set_stream")

    def reset_stream(self):
        self._check(self._lib.qs_set_stream(self._h, None), "qs_set_stream")

    def set_count_scale(self, count_scale: int):
        self._check(self._lib.qs_set_count_scale(self._h, count_scale), "qs_set_count_scale")

    def set_reference(self, ref: FlatReference):
        arrs = [np.ascontiguousarray(x, np.int32) for x in (ref.parent, ref.parent_edge, ref.leaf_lookup_id, ref.first_child, ref.next_sibling)]
        self._check(self._lib.qs_set_reference(self._h, ref.n_nodes, *[_ptr(a, C.c_int32) for a in arrs]), "qs_set_reference")
        self.edge_count = ref.edge_count

    def add_trees(self, flat: FlatTrees):
        self.add_trees_raw(flat.node_offsets, flat.parent, flat.leaf_lookup_id)

    def add_trees_raw(self, node_offsets: np.ndarray, parent: np.ndarray, leaf_lookup_id: np.ndarray):
        off = np.ascontiguousarray(node_offsets, np.int64)
        par = np.ascontiguousarray(parent, np.int32)
        leaf = np.ascontiguousarray(leaf_lookup_id, np.int32)
        self._check(self._lib.qs_add_trees(self._h, len(off) - 1, _ptr(off, C.c_int64), _ptr(par, C.c_int32), _ptr(leaf, C.c_int32)), "qs_add_trees")

    def add_trees_ptr(self, n_trees: int, off_ptr: int, parent_ptr: int, leaf_ptr: int):
        """Same call with raw host addresses (e.g. pinned torch tensors): no numpy copies."""
        self._check(self._lib.qs_add_trees(self._h, n_trees, C.cast(off_ptr, C.POINTER(C.c_int64)), C.cast(parent_ptr, C.POINTER(C.c_int32)),
                                           C.cast(leaf_ptr, C.POINTER(C.c_int32))), "qs_add_trees")

    def add_newick(self, text: Union[str, bytes], taxa: Sequence[str], n_threads: int = 0) -> int:
        """Native single-pass parallel ingest (csrc/ingest.cpp): Newick text -> device, returns the trees added."""
        data = text.encode() if isinstance(text, str) else text
        arr = (C.c_char_p * len(taxa))(*[t.encode() for t in taxa])
        got = C.c_int64()
        self._check(self._lib.qs_add_newick(self._h, data, len(data), arr, n_threads, C.byref(got)), "qs_add_newick")
        return got.value

    def add_newick_file(self, path: str, taxa: Sequence[str], n_threads: int = 0) -> int:
        arr = (C.c_char_p * len(taxa))(*[t.encode() for t in taxa])
        got = C.c_int64()
        self._check(self._lib.qs_add_newick_file(self._h, path.encode(), arr, n_threads, C.byref(got)), "qs_add_newick_file")
        return got.value

    def save_table(self, path: str):
        self._check(self._lib.qs_save_table(self._h, path.encode()), "qs_save_table")

    def load_table(self, path: str):
        self._check(self._lib.qs_load_table(self._h, path.encode()), "qs_load_table")

    def clear_trees(self):
        self._check(self._lib.qs_clear_trees(self._h), "qs_clear_trees")

    def num_trees(self) -> int:
        v = C.c_int64()
        self._check(self._lib.qs_num_trees(self._h, C.byref(v)), "qs_num_trees")
        return v.value

    def count(self):
        self._check(self._lib.qs_count(self._h), "qs_count")

    def score(self, count_scale: int = 1, exact_qp: bool = False):
        E = self.edge_count
        lq, qp, eqp = (np.empty(E, np.float64) for _ in range(3))
        self._check(self._lib.qs_score(self._h, count_scale, int(exact_qp), _ptr(lq, C.c_double), _ptr(qp, C.c_double), _ptr(eqp, C.c_double)), "qs_score")
        return lq, qp, eqp

    def num_pairs(self) -> int:
        v = C.c_int64()
        self._check(self._lib.qs_score_num_pairs(self._h, C.byref(v)), "qs_score_num_pairs")
        return v.value

    def inner_nodes(self) -> np.ndarray:
        """reference-tree node id of every inner index (the order of the per-pair arrays, see include/qscuda.h)"""
        n = C.c_int64()
        self._check(self._lib.qs_score_inner_nodes(self._h, None, 0, C.byref(n)), "qs_score_inner_nodes")
        out = np.empty(n.value, np.int32)
        self._check(self._lib.qs_score_inner_nodes(self._h, _ptr(out, C.c_int32), n.value, C.byref(n)), "qs_score_inner_nodes")
        return out

    def score_partials(self, count_scale: int = 1):
        lq = np.empty(self.edge_count, np.float64)
        sums = np.empty(self.num_pairs() * 3, np.uint64)
        self._check(self._lib.qs_score_partials(self._h, count_scale, _ptr(lq, C.c_double), _ptr(sums, C.c_uint64)), "qs_score_partials")
        return lq, sums

    def score_finalize(self, lqic_reduced: np.ndarray, pair_sums_reduced: np.ndarray, exact_qp: bool = False):
        E = self.edge_count
        lqr = np.ascontiguousarray(lqic_reduced, np.float64)
        ps = np.ascontiguousarray(pair_sums_reduced, np.uint64)
        lq, qp, eqp = (np.empty(E, np.float64) for _ in range(3))
        self._check(self._lib.qs_score_finalize(self._h, int(exact_qp), _ptr(lqr, C.c_double), _ptr(ps, C.c_uint64), _ptr(lq, C.c_double),
                                                _ptr(qp, C.c_double), _ptr(eqp, C.c_double)), "qs_score_finalize")
        return lq, qp, eqp

    # ---- multi-GPU, device-resident partials (include/qscuda.h) ----
    def score_scan(self, count_scale: int = 1):
        self._check(self._lib.qs_score_scan(self._h, count_scale), "qs_score_scan")

    def score_device_partials(self):
        """(pair_sums_ptr, pair_score_ptr, pair_best_ptr, n_pairs): raw device addresses of this context's partials."""
        a, b, c, n = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_int64()
        self._check(self._lib.qs_score_device_partials(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(n)), "qs_score_device_partials")
        return a.value, b.value, c.value, n.value

    def score_select_winners(self):
        self._check(self._lib.qs_score_select_winners(self._h), "qs_score_select_winners")

    def score_finish(self, exact_qp: bool = False):
        E = self.edge_count
        lq, qp, eqp = (np.empty(E, np.float64) for _ in range(3))
        self._check(self._lib.qs_score_finish(self._h, int(exact_qp), _ptr(lq, C.c_double), _ptr(qp, C.c_double), _ptr(eqp, C.c_double)), "qs_score_finish")
        return lq, qp, eqp

    def rebalance_shards(self) -> bool:
        """Re-cut the shard ranges for the class mix of the trees added so far (every shard must call it); True if this
        shard's range moved (count again)."""
        ch = C.c_int()
        self._check(self._lib.qs_rebalance_shards(self._h, C.byref(ch)), "qs_rebalance_shards")
        return bool(ch.value)

    def table_resident(self) -> bool:
        v = C.c_int()
        self._check(self._lib.qs_table_resident(self._h, C.byref(v)), "qs_table_resident")
        return bool(v.value)

    def shard_range(self):
        a, b = C.c_uint64(), C.c_uint64()
        self._check(self._lib.qs_shard_range(self._h, C.byref(a), C.byref(b)), "qs_shard_range")
        return a.value, b.value

    def get_counts(self, rank_begin: int = 0, rank_end: Optional[int] = None) -> np.ndarray:
        if rank_end is None:
            rank_end = comb(self.n_taxa, 4)
        out = np.empty((rank_end - rank_begin, 3), _NP_CINT[self.cint_bytes])
        self._check(self._lib.qs_get_counts(self._h, rank_begin, rank_end, out.ctypes.data_as(C.c_void_p)), "qs_get_counts")
        return out

    def get_distances(self, tree: int) -> np.ndarray:
        out = np.empty((self.n_taxa, self.n_taxa), np.uint16)
        self._check(self._lib.qs_get_distances(self._h, tree, _ptr(out, C.c_uint16)), "qs_get_distances")
        return out

    def write_raw_qic(self, taxa: Sequence[str], path: str, count_scale: int = 1):
        arr = (C.c_char_p * len(taxa))(*[t.encode() for t in taxa])
        self._check(self._lib.qs_write_raw_qic(self._h, count_scale, arr, path.encode()), "qs_write_raw_qic")

    @staticmethod
    def write_raw_qic_shards(ctxs, taxa: Sequence[str], path: str, count_scale: int = 1):
        """-q file from the contexts of ALL shards (same process, shard 0 .. G-1): qs_write_raw_qic_shards."""
        lib = ctxs[0]._lib
        arr = (C.c_char_p * len(taxa))(*[t.encode() for t in taxa])
        hs = (C.c_void_p * len(ctxs))(*[c._h for c in ctxs])
        ctxs[0]._check(lib.qs_write_raw_qic_shards(hs, len(ctxs), count_scale, arr, path.encode()), "qs_write_raw_qic_shards")

    def last_timing(self):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        self._check(self._lib.qs_last_timing(self._h, C.byref(a), C.byref(b), C.byref(c)), "qs_last_timing")
        return {"dist_ms": a.value, "count_ms": b.value, "score_ms": c.value}

    def launch_count(self) -> int:
        v = C.c_int64()
        self._check(self._lib.qs_launch_count(self._h, C.byref(v)), "qs_launch_count")
        return v.value

    def tree_classes(self):
        """(class A, class B) gene-tree counts of the last count(): class A = complete and fully resolved."""
        a, b = C.c_int64(), C.c_int64()
        self._check(self._lib.qs_tree_classes(self._h, C.byref(a), C.byref(b)), "qs_tree_classes")
        return a.value, b.value

    def measure_alu_peak(self):
        a, b = C.c_double(), C.c_double()
        self._check(self._lib.qs_measure_alu_peak(self._h, C.byref(a), C.byref(b)), "qs_measure_alu_peak")
        return a.value, b.value


def flatten_newick_native(text: Union[str, bytes], taxa: Sequence[str], n_threads: int = 0) -> FlatTrees:
    """Newick text -> FlatTrees through the library's parallel parser (qs_newick_flatten; no GPU, no context).
    Same arrays as newick.flatten_eval_trees(parse_newick_many(text), taxa); an unknown taxon raises QSError."""
    lib = _ffi.load()
    data = text.encode() if isinstance(text, str) else text
    arr = (C.c_char_p * len(taxa))(*[t.encode() for t in taxa])
    h = C.c_void_p()
    err = C.create_string_buffer(512)
    rc = lib.qs_newick_flatten(data, len(data), len(taxa), arr, n_threads, C.byref(h), err, len(err))
    if rc != 0:
        raise _ffi.QSError(rc, "qs_newick_flatten: " + err.value.decode(errors="replace"))
    try:
        T, N = C.c_int64(), C.c_int64()
        po, pp, pl = C.POINTER(C.c_int64)(), C.POINTER(C.c_int32)(), C.POINTER(C.c_int32)()
        lib.qs_flat_trees_view(h, C.byref(T), C.byref(N), C.byref(po), C.byref(pp), C.byref(pl))
        off = np.ctypeslib.as_array(po, shape=(T.value + 1,)).copy()
        par = np.ctypeslib.as_array(pp, shape=(N.value,)).copy() if N.value else np.empty(0, np.int32)
        leaf = np.ctypeslib.as_array(pl, shape=(N.value,)).copy() if N.value else np.empty(0, np.int32)
    finally:
        lib.qs_flat_trees_free(h)
    return FlatTrees(off, par, leaf)


class QuartetScoreComputer:
    """Mirror of ``QuartetScoreComputer<CINT>`` (src/QuartetScoreComputer.hpp:43-51).

    ``ref_tree`` is a Newick string / parsed Node; ``eval_trees`` a path to a Newick file (as in the
    reference), a Newick string with several trees, or an already flattened ``FlatTrees``.
    ``savemem`` selects the reference's -s semantics: the stored counts are the doubled, CINT-wrapped
    ones and no table is kept on the device.
    """

    def __init__(self, ref_tree: Union[str, Node], eval_trees: Union[str, FlatTrees], m: Optional[int] = None, verbose: bool = False,
                 savemem: bool = False, device: int = 0):
        self.ref_root = parse_newick(ref_tree) if isinstance(ref_tree, str) else ref_tree
        self.ref = flatten_reference(self.ref_root)
        if isinstance(eval_trees, FlatTrees):
            flat = eval_trees
        else:
            text = eval_trees
            if "(" not in text:
                with open(text) as f:
                    text = f.read()
            try:
                flat = flatten_newick_native(text, self.ref.taxa)
            except _ffi.QSError as e:      # reference: std::out_of_range from unordered_map::at (QuartetCounterLookup.hpp:218)
                if "is not in the reference tree" in str(e):
                    raise IndexError(f"unordered_map::at: {e}") from None
                raise
        self.m = flat.n_trees if m is None else m
        self.savemem = savemem
        self.cint_bytes = cint_bytes_for(self.m)
        self.count_scale = 2 if savemem else 1
        # NOTE: -s keeps the reference's *semantics* (doubled counts); whether a table is kept on the
        # device is a capacity decision made by the library, not by this flag.
        self.ctx = Context(self.ref.n_taxa, self.cint_bytes, _ffi.QS_MODE_TABLE, device)
        self.ctx.set_reference(self.ref)
        self.ctx.add_trees(flat)
        self.ctx.count()
        lq, qp, eqp = self.ctx.score(self.count_scale)
        self._lqic = lq
        bif = bool(np.isfinite(qp).any() or np.isfinite(eqp).any())
        self._qpic = qp if bif else np.empty(0)
        self._eqpic = eqp if bif else np.empty(0)
        if verbose:
            print(f"There are {self.m} evaluation trees.\nThe reference tree has {self.ref.n_taxa} taxa.")

    def getLQICScores(self) -> np.ndarray:
        return self._lqic

    def getQPICScores(self) -> np.ndarray:
        return self._qpic

    def getEQPICScores(self) -> np.ndarray:
        return self._eqpic

    def countQuartetOccurrences(self, a: int, b: int, c: int, d: int):
        """Counts of ab|cd, ac|bd, ad|bc for lookup ids a,b,c,d (any order)."""
        s = sorted((a, b, c, d))
        rank = comb(s[3], 4) + comb(s[2], 3) + comb(s[1], 2) + s[0]
        tup = self.ctx.get_counts(rank, rank + 1)[0].astype(np.uint64) * self.count_scale
        tup &= np.uint64((1 << (8 * self.cint_bytes)) - 1) if self.cint_bytes < 8 else np.uint64(0xFFFFFFFFFFFFFFFF)

        def slot(w, x, y, z):  # quartet_lookup_table.hpp:87-111
            lo1, hi1, lo2, hi2 = min(w, x), max(w, x), min(y, z), max(y, z)
            if hi1 < lo2 or hi2 < lo1:
                return 0
            if (lo1 < lo2 and hi2 < hi1) or (lo2 < lo1 and hi1 < hi2):
                return 2
            return 1

        return int(tup[slot(a, b, c, d)]), int(tup[slot(a, c, b, d)]), int(tup[slot(a, d, b, c)])

    def printRawQICScores(self, path: str):
        self.ctx.write_raw_qic(self.ref.taxa, path, self.count_scale)
