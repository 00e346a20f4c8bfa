"""quartetscores_b200 — B200-native (sm_100a) quartet counting and internode-certainty scoring.

Drop-in for the hot path of lutteropp/QuartetScores (QuartetCounterLookup + QuartetLookupTable +
TreeInformation + QuartetScoreComputer).  The compute lives in libqscuda.so (CUDA, C ABI in
include/qscuda.h); this package is the Python host side: Newick plumbing, flattening, the mirror of the
reference's QuartetScoreComputer interface, synthetic inputs and the multi-GPU driver.
"""
from .computer import Context, QuartetScoreComputer, cint_bytes_for, flatten_newick_native  # noqa: F401
from ._ffi import QSError, QS_DEVICE_NONE, QS_MODE_AUTO, QS_MODE_TABLE, QS_MODE_TABLE_FREE  # noqa: F401

__all__ = ["Context", "QuartetScoreComputer", "QSError", "cint_bytes_for", "flatten_newick_native", "QS_MODE_TABLE", "QS_MODE_TABLE_FREE", "QS_MODE_AUTO", "QS_DEVICE_NONE"]
