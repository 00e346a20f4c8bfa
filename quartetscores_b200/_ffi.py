"""ctypes binding of libqscuda.so (C ABI in include/qscuda.h).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  There is no fallback:
if the shared library is missing, importing a symbol raises, and without a CUDA device
``qs_create`` fails with QS_E_CUDA.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QSCUDA_LIB") or os.path.join(_HERE, "libqscuda.so")   # QSCUDA_LIB: tuning builds (tools/)

QS_OK = 0
QS_MODE_TABLE = 0
QS_MODE_TABLE_FREE = 1
QS_MODE_AUTO = 2
QS_DEVICE_NONE = -1

_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_u64p = C.POINTER(C.c_uint64)
_f64p = C.POINTER(C.c_double)

# name -> (restype, argtypes); mirrors include/qscuda.h one to one
SIGNATURES = {
    "qs_abi_version": (C.c_int, []),
    "qs_strerror": (C.c_char_p, [C.c_int]),
    "qs_last_error": (C.c_char_p, [C.c_void_p]),
    "qs_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "qs_destroy": (C.c_int, [C.c_void_p]),
    "qs_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "qs_set_count_scale": (C.c_int, [C.c_void_p, C.c_int]),
    "qs_set_reference": (C.c_int, [C.c_void_p, C.c_int, _i32p, _i32p, _i32p, _i32p, _i32p]),
    "qs_add_trees": (C.c_int, [C.c_void_p, C.c_int, _i64p, _i32p, _i32p]),
    "qs_clear_trees": (C.c_int, [C.c_void_p]),
    "qs_num_trees": (C.c_int, [C.c_void_p, _i64p]),
    "qs_count": (C.c_int, [C.c_void_p]),
    "qs_score": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _f64p, _f64p, _f64p]),
    "qs_score_num_pairs": (C.c_int, [C.c_void_p, _i64p]),
    "qs_score_inner_nodes": (C.c_int, [C.c_void_p, _i32p, C.c_int64, _i64p]),
    "qs_score_partials": (C.c_int, [C.c_void_p, C.c_int, _f64p, _u64p]),
    "qs_score_finalize": (C.c_int, [C.c_void_p, C.c_int, _f64p, _u64p, _f64p, _f64p, _f64p]),
    "qs_score_scan": (C.c_int, [C.c_void_p, C.c_int]),
    "qs_score_device_partials": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), _i64p]),
    "qs_score_select_winners": (C.c_int, [C.c_void_p]),
    "qs_score_finish": (C.c_int, [C.c_void_p, C.c_int, _f64p, _f64p, _f64p]),
    "qs_get_counts": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]),
    "qs_shard_range": (C.c_int, [C.c_void_p, _u64p, _u64p]),
    "qs_shard_bounds": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), _u64p, _u64p]),
    "qs_rebalance_shards": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "qs_table_resident": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "qs_plan_stats": (C.c_int, [C.c_int, C.c_int, C.c_int, _i64p]),
    "qs_get_distances": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_uint16)]),
    "qs_write_raw_qic": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.c_char_p]),
    "qs_write_raw_qic_shards": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.POINTER(C.c_char_p), C.c_char_p]),
    "qs_newick_flatten": (C.c_int, [C.c_char_p, C.c_size_t, C.c_int, C.POINTER(C.c_char_p), C.c_int, C.POINTER(C.c_void_p), C.c_char_p, C.c_size_t]),
    "qs_flat_trees_view": (C.c_int, [C.c_void_p, _i64p, _i64p, C.POINTER(_i64p), C.POINTER(_i32p), C.POINTER(_i32p)]),
    "qs_flat_trees_free": (None, [C.c_void_p]),
    "qs_add_newick": (C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_char_p), C.c_int, _i64p]),
    "qs_add_newick_file": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_char_p), C.c_int, _i64p]),
    "qs_save_table": (C.c_int, [C.c_void_p, C.c_char_p]),
    "qs_load_table": (C.c_int, [C.c_void_p, C.c_char_p]),
    "qs_last_timing": (C.c_int, [C.c_void_p, _f64p, _f64p, _f64p]),
    "qs_launch_count": (C.c_int, [C.c_void_p, _i64p]),
    "qs_tree_classes": (C.c_int, [C.c_void_p, _i64p, _i64p]),
    "qs_measure_alu_peak": (C.c_int, [C.c_void_p, _f64p, _f64p]),
}

_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). quartetscores_b200 has no CPU fallback."
            )
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)       # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class QSError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"[{code}] {message}")
        self.code = code
