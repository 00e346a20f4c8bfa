#!/usr/bin/env python3
"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Runs only where oracle/_ref/ has been built (`make -C oracle ref`, needs /root/reference, i.e. the
build container; the GPU box only sees the committed .npz files).  For every case below it writes the
seeded synthetic input to a temp dir, runs

    oracle/_ref/qs_ref_dump   (harness that includes the reference headers; oracle/ref_dump.cpp)
    oracle/_ref/QuartetScores (the reference's own main) for the annotated Newick / -q text

and stores inputs + reference outputs in tests/golden/<case>.npz:
    ref_newick, eval_newick      input text
    counts        uint16/uint32 [C(n,4),3] canonical table (fast mode), rank order
    counts_s      same from the -s (compact) table, run with 1 thread
    lqic,qpic,eqpic  float64[edges] (fast mode);  *_s for the -s run
    out_newick    the reference's -o file (fast mode), out_newick_s for -s
    rawqic        the reference's -q file
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from quartetscores_b200.synth import SyntheticInput  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")

C1_REF = "((A,B),(C,D),(E,(F,G)));"
C1_EVAL = ["((A,B),(C,D),(E,(F,G)));", "((A,C),(B,D),(E,(F,G)));", "((A,B),(C,D),(F,(E,G)));", "(A,B,C,D,(E,(F,G)));", "((A,B),(C,E),(F,G));"]
C2_REF = "((A,B),(C,D),(E,F,G));"
C3_REF = "((A,B),(C,D),E);"
C3_EVAL = ["((A,B),(C,D),E);"] * 150 + ["((A,C),(B,D),E);"] * 50
# rooted gene trees, unary nodes, branch lengths, inner labels, comments (SURVEY §0 "hand set")
C4_REF = "((A:0.1,B:0.2)ab:0.3,(C:1,D:2)cd:1e-2,(E,(F,G)fg:3)x);"
C4_EVAL = ["(((A:1,B:2)n1:0.5,(C,D)),(E,(F,G)));", "((((A,B)),(C,D)),((E,(F,G))));", "(A[c1],(B,(C,(D,(E,(F,G))))));", "((G,F),(E,D),(C,(B,A)));", "((A,B,C),(D,E),(F,G));", "(A,B);", "((A,C),(E,G));"]

CASES = {
    # name: (kwargs for SyntheticInput) or explicit text
    "c1_known_answer": dict(ref=C1_REF, evals=C1_EVAL),
    "c2_multifurcating_ref": dict(ref=C2_REF, evals=C1_EVAL),
    "c3_u8_200trees": dict(ref=C3_REF, evals=C3_EVAL),
    "c4_rooted_unary_labels": dict(ref=C4_REF, evals=C4_EVAL),
    "s16x300_missing_poly": dict(n_taxa=16, n_trees=300, seed=1601, k_max=6, p_missing=0.1, p_contract=0.1),
    "s24x60_u8": dict(n_taxa=24, n_trees=60, seed=2401, k_max=8),
    "s20x40_multiref_missing": dict(n_taxa=20, n_trees=40, seed=2001, k_max=5, p_missing=0.15, multifurcating_reference=0.3),
    "s32x270_spr": dict(n_taxa=32, n_trees=270, seed=3201, k_max=12, nni_fraction=0.2),
    "s13x257_poly": dict(n_taxa=13, n_trees=257, seed=1301, k_max=4, p_contract=0.3),
}


def run_case(name, spec, tmp):
    if "ref" in spec:
        ref_nwk, evals = spec["ref"], spec["evals"]
    else:
        s = SyntheticInput(**spec)
        ref_nwk, evals = s.ref_newick, s.eval_newick
    rp, ep = os.path.join(tmp, "ref.nwk"), os.path.join(tmp, "eval.nwk")
    open(rp, "w").write(ref_nwk + "\n")
    open(ep, "w").write("\n".join(evals) + "\n")
    out = {"ref_newick": ref_nwk, "eval_newick": "\n".join(evals)}
    for suffix, savemem in (("", 0), ("_s", 1)):
        pref = os.path.join(tmp, "dump" + suffix)
        subprocess.check_call([os.path.join(REF, "qs_ref_dump"), rp, ep, pref, str(savemem), "1"], stdout=subprocess.DEVNULL)
        counts = np.fromfile(pref + ".counts.u64", dtype=np.uint64).reshape(-1, 3)
        dt = np.uint16 if counts.max(initial=0) < 65536 else np.uint32
        out["counts" + suffix] = counts.astype(dt)
        sc = np.fromfile(pref + ".scores.f64", dtype=np.float64).reshape(3, -1)
        out["lqic" + suffix], out["qpic" + suffix], out["eqpic" + suffix] = sc[0], sc[1], sc[2]
        if not savemem:
            out["rawqic"] = open(pref + ".rawqic.txt").read()
            out["meta"] = open(pref + ".meta.txt").read()
        op = os.path.join(tmp, "out" + suffix + ".nwk")
        if os.path.exists(op):
            os.remove(op)
        args = [os.path.join(REF, "QuartetScores"), "-r", rp, "-e", ep, "-o", op, "-t", "1"] + (["-s"] if savemem else [])
        subprocess.check_call(args, stdout=subprocess.DEVNULL)
        out["out_newick" + suffix] = open(op).read()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: {len(out['counts'])} quartets, {len(evals)} trees, bytes={os.path.getsize(os.path.join(HERE, name + '.npz'))}")


if __name__ == "__main__":
    only = sys.argv[1:]
    with tempfile.TemporaryDirectory() as tmp:
        for name, spec in CASES.items():
            if only and name not in only:
                continue
            run_case(name, spec, tmp)
