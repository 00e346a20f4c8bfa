"""Soundness margin of the scan kernel's fp32 pre-filter (kernels/score.cuh: dev_log_score_f32, QS_FILTER_MARGIN).

The kernel drops a quartet from the min-QIC selection only when its fp32 estimate exceeds a bound by more than
QS_FILTER_MARGIN = 2e-6, which is sound while |estimate - exact| stays below half of that.  This restates the
estimate in numpy float32 (same operation order; log2 within an ulp of CUDA's log2f) and measures the error
against the fp64 log_score of the reference (src/QuartetScoreComputer.hpp:135-159) on random and adversarial triples."""
import numpy as np

MARGIN = 2e-6


def est_f32(q1, q2, q3):
    q1f, q2f, q3f = (x.astype(np.float32) for x in (q1, q2, q3))
    s = (q1 + q2 + q3).astype(np.float32)
    inv = np.float32(1.0) / s
    acc = np.zeros_like(s)
    for q in (q1f, q2f, q3f):
        p = q * inv
        with np.errstate(divide="ignore", invalid="ignore"):
            term = np.where(q > 0, p * np.log2(np.where(q > 0, p, 1).astype(np.float32)), np.float32(0)).astype(np.float32)
        acc = (acc + term).astype(np.float32)
    qic = (np.float32(1.0) + acc * np.float32(0.63092975357145743710)).astype(np.float32)
    return np.where((q1 < q2) | (q1 < q3), -qic, qic)


def exact_f64(q1, q2, q3):
    s = (q1 + q2 + q3).astype(np.float64)
    out = np.ones_like(s)
    for q in (q1, q2, q3):
        p = q / s
        with np.errstate(divide="ignore", invalid="ignore"):
            out = out + np.where(q > 0, p * np.log(np.where(q > 0, p, 1)) / np.log(3.0), 0.0)
    return np.where((q1 < q2) | (q1 < q3), -out, out)


def test_fp32_estimate_error_is_far_below_the_filter_margin():
    rng = np.random.default_rng(3)
    cases = []
    for hi in (10, 300, 5000, 65535, 131070):            # count ranges up to 2 x uint16 (count_scale 2)
        cases.append(rng.integers(0, hi + 1, size=(400000, 3)))
    near = rng.integers(1, 131070, size=(400000, 1))      # nearly unanimous and nearly uniform triples
    cases.append(np.concatenate([near, rng.integers(0, 3, size=(400000, 2))], axis=1))
    cases.append(near + rng.integers(0, 3, size=(400000, 3)))
    q = np.concatenate(cases).astype(np.int64)
    q = q[q.sum(axis=1) > 0]
    err = np.abs(est_f32(q[:, 0], q[:, 1], q[:, 2]).astype(np.float64) - exact_f64(q[:, 0], q[:, 1], q[:, 2]))
    assert err.max() < MARGIN / 4, err.max()
