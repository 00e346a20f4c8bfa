#!/usr/bin/env python3
"""Golden digests of the UNMODIFIED reference on BASELINE-shaped inputs that are too big to store in full.

Runs only where oracle/_ref/ has been built (needs /root/reference).  For each case it writes the seeded
synthetic input (the same SyntheticInput bench.py uses), runs oracle/_ref/qs_ref_dump (harness that includes
the reference headers, fast table, all host threads -- the fast table is deterministic at any thread count,
SURVEY App. B3) and stores in tests/golden/big/<case>.npz:

    spec            the SyntheticInput keyword arguments (json)
    counts_sha256   sha256 of the canonical table as little-endian uint16 [C(n,4)][3] in rank order
    counts_sample   every `stride`-th table entry (uint16 [k][3]) and `stride`, for a readable diff when the digest differs
    lqic,qpic,eqpic float64[edges]
    out_newick      the reference's -o file
    rawqic_sha256   sha256 of the reference's -q file, rawqic_lines, rawqic_head (first 20 lines)

cfg2 (100 taxa x 10,000 trees, seed 2000) takes ~10 minutes on 8 cores (the harness counts twice).
"""
import hashlib
import json
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
OUT = os.path.join(HERE, "big")

CASES = {
    # BASELINE.json configs[1], exactly bench.py's WORKLOADS["cfg2"]
    "cfg2_100x10000": dict(n_taxa=100, n_trees=10000, seed=2000, k_max=20),
    # BASELINE.json configs[0]
    "cfg1_50x1000": dict(n_taxa=50, n_trees=1000, seed=1000, k_max=10),
}


def sha_file(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 24), b""):
            h.update(chunk)
    return h.hexdigest()


def run_case(name, spec, tmp, threads):
    s = SyntheticInput(**spec)
    rp, ep = os.path.join(tmp, "ref.nwk"), os.path.join(tmp, "eval.nwk")
    open(rp, "w").write(s.ref_newick + "\n")
    open(ep, "w").write("\n".join(s.eval_newick) + "\n")
    pref = os.path.join(tmp, "dump")
    subprocess.check_call([os.path.join(REF, "qs_ref_dump"), rp, ep, pref, "0", str(threads)], stdout=subprocess.DEVNULL)
    counts = np.fromfile(pref + ".counts.u64", dtype=np.uint64).reshape(-1, 3)
    assert counts.max() < 65536
    c16 = np.ascontiguousarray(counts.astype("<u2"))
    stride = max(1, len(c16) // 4096)
    sc = np.fromfile(pref + ".scores.f64", dtype=np.float64).reshape(3, -1)
    op = os.path.join(tmp, "out.nwk")
    if os.path.exists(op):
        os.remove(op)          # the reference refuses to overwrite (src/QuartetScores.cpp:81-85)
    subprocess.check_call([os.path.join(REF, "QuartetScores"), "-r", rp, "-e", ep, "-o", op, "-t", str(threads)], stdout=subprocess.DEVNULL)
    raw = pref + ".rawqic.txt"
    with open(raw) as f:
        head = [next(f) for _ in range(20)]
    n_lines = sum(1 for _ in open(raw))
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), spec=json.dumps(spec), counts_sha256=hashlib.sha256(c16.tobytes()).hexdigest(),
                        counts_sample=c16[::stride], stride=stride, lqic=sc[0], qpic=sc[1], eqpic=sc[2], out_newick=open(op).read(),
                        rawqic_sha256=sha_file(raw), rawqic_lines=n_lines, rawqic_head="".join(head))
    print(f"{name}: {len(c16)} quartets, sha256 {hashlib.sha256(c16.tobytes()).hexdigest()[:16]}..., {n_lines} raw lines")


if __name__ == "__main__":
    only = sys.argv[1:]
    with tempfile.TemporaryDirectory() as tmp:
        for name, spec in CASES.items():
            if only and name not in only:
                continue
            run_case(name, spec, tmp, os.cpu_count() or 1)
