"""Host logic of the counting kernel (CPU only): the task table built by libqscuda (qs_plan_stats runs the
builder and its self-check: tasks tile every item enumeration exactly once, every item decodes to sane taxon ids,
its two matrix rows lie inside the task's staged row ranges, and the quartets that the items keep after their flush
masks add up to exactly the quartets of the range — once over the role-X kinds, once over roles Y and Z)."""
import ctypes as C
from math import comb

import numpy as np
import pytest

from quartetscores_b200 import _ffi
from quartetscores_b200.multi import shard_bounds


def plan_stats(n, b, e):
    st = np.zeros(17, np.int64)
    rc = _ffi.load().qs_plan_stats(n, b, e, st.ctypes.data_as(C.POINTER(C.c_int64)))
    assert rc == 0
    keys = ["x_tasks", "y_tasks", "xo_items", "xd_items", "y_items", "xo_slots", "xd_slots", "y_slots", "rows", "max_rows", "violations", "quartets",
            "xr_items", "xr_slots", "x_compares", "z_items", "y_compares"]
    return dict(zip(keys, (int(x) for x in st)))


@pytest.mark.parametrize("n", [4, 5, 8, 9, 17, 33, 50, 100, 161, 230])
def test_plan_self_check_whole_space(n):
    s = plan_stats(n, 0, n)
    assert s["violations"] == 0
    assert s["quartets"] == comb(n, 4)
    # every quartet lies in exactly one 8x8 block of role X and one of role Y: the blocks must at least cover them
    assert (s["xo_items"] + s["xr_items"]) * 64 + s["xd_items"] * 28 >= comb(n, 4) and s["y_items"] * 64 >= comb(n, 4)
    assert s["x_compares"] >= 2 * comb(n, 4) and s["y_compares"] >= comb(n, 4)


@pytest.mark.parametrize("n,G", [(40, 3), (100, 8), (300, 8)])
def test_plan_self_check_shards(n, G):
    total = 0
    for g in range(G):
        b, e, rb, re_ = shard_bounds(n, g, G)
        s = plan_stats(n, b, e)
        assert s["violations"] == 0 and s["quartets"] == re_ - rb
        total += s["quartets"]
    assert total == comb(n, 4)


def test_plan_efficiency_cfg2_shape():
    """BASELINE config 2 shape: tasks are full and most compared lanes are real quartets."""
    s = plan_stats(100, 0, 100)
    assert s["xo_items"] / s["xo_slots"] > 0.98 and s["y_items"] / s["y_slots"] > 0.95
    useful_x = 2 * s["quartets"] / s["x_compares"]          # the ragged b-block of every (c,d) only runs over its c & 7 valid taxa (XR items)
    assert useful_x > 0.96


def test_plan_self_check_random_ranges():
    """Arbitrary d-ranges (slabs, shards of any width): role Y owns whole d-blocks, role Z the d of the cut blocks or — for a
    range with few whole blocks — every d; the plan's self-check covers the split, the decoding of every item and its rows."""
    rng = np.random.default_rng(7)
    seen_all_z = seen_both = 0
    for _ in range(40):
        n = int(rng.integers(8, 260))
        b = int(rng.integers(0, n - 1))
        e = int(rng.integers(b + 1, n + 1))
        s = plan_stats(n, b, e)
        lo = max(3, b)
        assert s["violations"] == 0, (n, b, e, s)
        assert s["quartets"] == (comb(e, 4) - comb(lo, 4) if e > lo else 0)
        if e > lo:
            assert s["y_compares"] >= s["quartets"]
            seen_all_z += s["z_items"] == s["y_items"]
            seen_both += 0 < s["z_items"] < s["y_items"]
    assert seen_all_z >= 5 and seen_both >= 5
