"""ctypes wrapper around oracle/libqs_oracle.so (the CPU restatement).  TEST INFRASTRUCTURE:
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from math import comb

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(ORACLE_DIR, "libqs_oracle.so")
        src = os.path.join(ORACLE_DIR, "qs_oracle.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "oracle"], stdout=subprocess.DEVNULL)
        L = C.CDLL(so)
        L.qso_rank.restype = C.c_uint64
        L.qso_rank.argtypes = [C.c_uint64] * 4
        L.qso_tuple_index.argtypes = [C.c_uint64] * 4
        L.qso_log_score.restype = C.c_double
        L.qso_log_score.argtypes = [C.c_uint64] * 3
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def rank(a, b, c, d):
    return int(lib().qso_rank(a, b, c, d))


def tuple_index(a, b, c, d):
    return int(lib().qso_tuple_index(a, b, c, d))


def log_score(q1, q2, q3):
    return float(lib().qso_log_score(int(q1), int(q2), int(q3)))


def distance_matrix(parent, leaf_id, n_taxa):
    parent = np.ascontiguousarray(parent, np.int32)
    leaf_id = np.ascontiguousarray(leaf_id, np.int32)
    D = np.empty((n_taxa, n_taxa), np.uint16)
    r = lib().qso_distance_matrix(len(parent), _p(parent, C.c_int32), _p(leaf_id, C.c_int32), n_taxa, _p(D, C.c_uint16))
    if r < 0:
        raise RuntimeError(f"qso_distance_matrix failed: {r}")
    return D, r


def _count(fn, n_taxa, flat):
    off = np.ascontiguousarray(flat.node_offsets, np.int64)
    par = np.ascontiguousarray(flat.parent, np.int32)
    leaf = np.ascontiguousarray(flat.leaf_lookup_id, np.int32)
    table = np.zeros((comb(n_taxa, 4), 3), np.uint32)
    r = fn(n_taxa, len(off) - 1, _p(off, C.c_int64), _p(par, C.c_int32), _p(leaf, C.c_int32), _p(table, C.c_uint32))
    if r != 0:
        raise RuntimeError(f"oracle counting failed: {r}")
    return table


def count_clades_compact(n_taxa, flat):
    """Reference -s table semantics: 2 per tree."""
    return _count(lib().qso_count_clades_compact, n_taxa, flat)


def count_clades_fast(n_taxa, flat):
    """Reference fast-table semantics: canonical counts (n <= 64)."""
    return _count(lib().qso_count_clades_fast, n_taxa, flat)


def count_fourpoint(n_taxa, flat):
    return _count(lib().qso_count_fourpoint, n_taxa, flat)


def unrank(r):
    q = (C.c_uint64 * 4)()
    lib().qso_unrank(C.c_uint64(int(r)), q)
    return tuple(int(x) for x in q)


def count_fourpoint_ranks(n_taxa, flat, ranks):
    """Canonical counts of the table entries `ranks` only (uint32 [len(ranks), 3]) -- for n where C(n,4) is out of reach."""
    off = np.ascontiguousarray(flat.node_offsets, np.int64)
    par = np.ascontiguousarray(flat.parent, np.int32)
    leaf = np.ascontiguousarray(flat.leaf_lookup_id, np.int32)
    ranks = np.ascontiguousarray(ranks, np.uint64)
    out = np.zeros((len(ranks), 3), np.uint32)
    r = lib().qso_count_fourpoint_ranks(n_taxa, len(off) - 1, _p(off, C.c_int64), _p(par, C.c_int32), _p(leaf, C.c_int32),
                                        C.c_int64(len(ranks)), _p(ranks, C.c_uint64), _p(out, C.c_uint32))
    if r != 0:
        raise RuntimeError(f"qso_count_fourpoint_ranks failed: {r}")
    return out


def score(ref, table, count_scale=1, cint_bits=16):
    """ref: newick.FlatReference; table: canonical counts uint32[C(n,4),3].
    Returns (lqic, qpic, eqpic, bifurcating)."""
    E = ref.edge_count
    lq, qp, eqp = (np.empty(E, np.float64) for _ in range(3))
    bif = C.c_int(0)
    table = np.ascontiguousarray(table, np.uint32)
    r = lib().qso_score(ref.n_nodes, _p(ref.parent, C.c_int32), _p(ref.leaf_lookup_id, C.c_int32), _p(ref.parent_edge, C.c_int32),
                        _p(ref.child_rank, C.c_int32), ref.n_taxa, _p(table, C.c_uint32), count_scale, cint_bits,
                        _p(lq, C.c_double), _p(qp, C.c_double), _p(eqp, C.c_double), C.byref(bif))
    if r != 0:
        raise RuntimeError(f"qso_score failed: {r}")
    return lq, qp, eqp, bool(bif.value)


def raw_qic(ref, table, count_scale=1, cint_bits=16):
    nq = comb(ref.n_taxa, 4)
    topo = np.empty(nq, np.int8)
    qic = np.empty(nq, np.float64)
    table = np.ascontiguousarray(table, np.uint32)
    r = lib().qso_raw_qic(ref.n_nodes, _p(ref.parent, C.c_int32), _p(ref.leaf_lookup_id, C.c_int32), ref.n_taxa,
                          _p(table, C.c_uint32), count_scale, cint_bits, _p(topo, C.c_int8), _p(qic, C.c_double))
    if r != 0:
        raise RuntimeError(f"qso_raw_qic failed: {r}")
    return topo, qic


def raw_qic_text(ref, table, count_scale=1, cint_bits=16):
    """The -q file as the reference writes it (QuartetScoreComputer.hpp:684, ostream default precision)."""
    topo, qic = raw_qic(ref, table, count_scale, cint_bits)
    n, taxa = ref.n_taxa, ref.taxa
    lines = []
    i = 0
    for a in range(n):
        for b in range(a + 1, n):
            for c in range(b + 1, n):
                for d in range(c + 1, n):
                    t = topo[i]
                    if t == 0:
                        lines.append("(%s,%s|%s,%s): %s" % (taxa[a], taxa[b], taxa[c], taxa[d], "%g" % qic[i]))
                    elif t == 1:
                        lines.append("(%s,%s|%s,%s): %s" % (taxa[a], taxa[c], taxa[b], taxa[d], "%g" % qic[i]))
                    elif t == 2:
                        lines.append("(%s,%s|%s,%s): %s" % (taxa[a], taxa[d], taxa[b], taxa[c], "%g" % qic[i]))
                    i += 1
    return "\n".join(lines) + ("\n" if lines else "")
