"""The C-ABI boundary (include/qscuda.h) without a GPU: the shared library loads, exports every function the header
declares, the ctypes table mirrors the header one to one, and the product path fails loudly (no CPU fallback) when
there is no CUDA device."""
import ctypes as C
import os
import re

import pytest

from quartetscores_b200 import _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "qscuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"^\s*(?:const\s+char\s*\*|int|void)\s+(qs_\w+)\s*\(", text, flags=re.M)))


def test_header_declares_the_expected_surface():
    names = declared_functions()
    assert len(names) >= 30
    for must in ("qs_create", "qs_set_reference", "qs_add_trees", "qs_count", "qs_score", "qs_get_counts", "qs_write_raw_qic",
                 "qs_score_partials", "qs_score_finalize", "qs_newick_flatten", "qs_add_newick_file", "qs_save_table", "qs_load_table", "qs_destroy"):
        assert must in names


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(_ffi.LIB_PATH)
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, f"declared in include/qscuda.h but not exported by libqscuda.so: {missing}"


def test_ctypes_table_mirrors_the_header():
    names = set(declared_functions())
    assert set(_ffi.SIGNATURES) == names, (sorted(names - set(_ffi.SIGNATURES)), sorted(set(_ffi.SIGNATURES) - names))
    assert _ffi.load().qs_abi_version() == 1


def test_no_cpu_fallback_without_a_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    lib = _ffi.load()
    h = C.c_void_p()
    rc = lib.qs_create(C.byref(h), 10, 2, 0, 0, 0, 1)
    assert rc == -2 and not h.value                      # QS_E_CUDA: nothing is computed on the host instead
    assert b"CUDA" in lib.qs_strerror(rc)
    from quartetscores_b200 import Context, QSError
    with pytest.raises(QSError):
        Context(10, 2)


def test_argument_errors_do_not_crash():
    lib = _ffi.load()
    assert lib.qs_create(None, 10, 2, 0, 0, 0, 1) == -1                  # QS_E_ARG
    h = C.c_void_p()
    assert lib.qs_create(C.byref(h), 10, 3, 0, _ffi.QS_DEVICE_NONE, 0, 1) == -1      # cint_bytes must be 1, 2, 4 or 8
    assert lib.qs_create(C.byref(h), 10, 2, 7, _ffi.QS_DEVICE_NONE, 0, 1) == -1      # unknown mode
    assert lib.qs_count(None) == -1 and lib.qs_destroy(None) in (0, -1)
