"""The counting kernel's speed depends on what ptxas makes of `acc -= mask`: a two-input IMAD.IADD (FMA pipe) next to
every HSET2 (ALU pipe).  If the tree loop is unrolled, ptxas fuses the two subtracts of a counter into one three-input
IADD3 on the ALU pipe and the kernel loses 10 %.  This guards the property on the built library;
it needs cuobjdump (CUDA toolkit), not a GPU."""
import os
import re
import shutil
import subprocess

import pytest

from quartetscores_b200 import _ffi


def kernel_sass(pattern):
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([exe, "-sass", _ffi.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    chunks = re.split(r"\n\s*Function : ", out)
    for ch in chunks:
        if re.match(pattern, ch):
            return [l.split(";")[0].strip() for l in ch.splitlines() if re.match(r"\s*/\*[0-9a-f]{4,}\*/", l)]
    pytest.fail(f"no kernel matching {pattern} in {_ffi.LIB_PATH}")


@pytest.mark.parametrize("threads", [512, 256])
def test_counting_loop_pairs_every_hset2_with_an_imad_iadd(threads):
    lines = kernel_sass(rf"_ZN2qs20qs_count_rows_kernelILi{threads}EEEvNS_13CountRowsArgsE")
    lt = [i for i, l in enumerate(lines) if "HSET2.LT" in l]
    assert lt, "role-X loop (HSET2.LT) not found"
    # the role-X tree loops (one copy per number of valid taxa in the b-block, 1..8: ragged blocks run shorter loops): the
    # instructions around their HSET2.LT compares, up to the backward branch that closes the last one
    lo = lt[0] - 40
    hi = next(i for i in range(lt[-1], len(lines)) if " BRA " in lines[i] + " ")
    body = lines[max(lo, 0):hi]
    n_hset = sum("HSET2" in l for l in body)
    n_sub = sum("IMAD.IADD" in l and ", -R" in l for l in body)
    fused = [l for l in body if "IADD3" in l and l.count(", -R") >= 2]
    assert n_hset >= 64 * 36 // 8 and n_sub >= n_hset, (n_hset, n_sub)        # 8 (gt, lt) pairs of packed compares per valid taxon: 8 x (1 + ... + 8) in all copies
    assert not fused, fused[:3]
    assert not any(op in l for l in body for op in ("HADD2.F32", "STL", "LDL")), "spills or conversions inside the tree loop"


def test_counting_kernel_stages_rows_with_tma_bulk_copies():
    lines = kernel_sass(r"_ZN2qs20qs_count_rows_kernelILi512EEEvNS_13CountRowsArgsE")
    assert any(l.split()[1].startswith("UBLKCP") or " UBLKCP" in l for l in lines), "cp.async.bulk (UBLKCP) missing: rows are no longer staged by TMA"
    assert any("SYNCS.PHASECHK" in l for l in lines) and any("SYNCS.ARRIVE" in l for l in lines), "mbarrier pipeline missing"
    assert sum("LDS.128" in l for l in lines) >= 8


def test_counting_kernel_has_no_stack_frame():
    """A stack frame (ptxas spilling task-level values, or a helper that was not inlined) makes every launch reserve local
    memory: with 16 bytes of it the cfg2 step went from 4.9 to 8.1 ms and the end-to-end step from 5.3 to 37.7 ms
    (profiles/r02_w_notes.txt) although the kernel itself ran as fast as before."""
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([exe, "-res-usage", _ffi.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    found = 0
    for m in re.finditer(r"Function (\S*qs_count_rows_kernel\S*):\s*\n\s*REG:(\d+) STACK:(\d+)", out):
        found += 1
        assert int(m.group(3)) == 0, m.group(0)
    assert found == 2
