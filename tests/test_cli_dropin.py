"""Drop-in check on the GPU box: the reference binary and the same main() linked against libqscuda
(oracle/_ref/QuartetScoresB200 = src/QuartetScores.cpp unmodified + integration/QuartetScoreComputerB200.hpp)
must write byte-identical annotated Newick and raw-QIC files.  Both binaries are prebuilt by oracle/Makefile in
the build container (they need the reference sources to compile) and travel with the snapshot."""
import os
import subprocess

import pytest

from conftest import golden_cases

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "QuartetScores")
OUR_BIN = os.path.join(ROOT, "oracle", "_ref", "QuartetScoresB200")
needs_bins = pytest.mark.skipif(not (os.path.exists(REF_BIN) and os.path.exists(OUR_BIN)), reason="oracle/_ref binaries not built (no /root/reference at build time)")


def _run(exe, tmp, tag, ref_nwk, eval_nwk, extra, env=None):
    rp, ep, op, qp = (str(tmp / f"{tag}.{x}") for x in ("ref.nwk", "eval.nwk", "out.nwk", "raw.txt"))
    with open(rp, "w") as f:
        f.write(ref_nwk.strip() + "\n")
    with open(ep, "w") as f:
        f.write(eval_nwk.strip() + "\n")
    cmd = [exe, "-r", rp, "-e", ep, "-o", op, "-q", qp] + extra
    p = subprocess.run(cmd, capture_output=True, text=True, env=dict(os.environ, **(env or {})))
    assert p.returncode == 0, p.stdout + p.stderr
    return open(op).read(), open(qp).read(), p.stdout


@needs_bins
@pytest.mark.parametrize("name", golden_cases())
def test_cli_outputs_byte_identical(name, golden, tmp_path):
    g = golden(name)
    want_out, want_raw, _ = _run(REF_BIN, tmp_path, "ref", g["ref_newick"], g["eval_newick"], ["-t", "2"])
    got_out, got_raw, stdout = _run(OUR_BIN, tmp_path, "our", g["ref_newick"], g["eval_newick"], ["-t", "2"])
    assert got_out == want_out
    assert got_raw == want_raw
    assert got_out.strip() == g["out_newick"].strip()
    assert "Elapsed time:" in stdout


@needs_bins
@pytest.mark.parametrize("name", ["c1_known_answer", "c3_u8_200trees", "s16x300_missing_poly"])
def test_cli_savemem_flag(name, golden, tmp_path):
    """-s: the reference stores doubled, CINT-wrapped counts; -t 1 because its -s path races."""
    g = golden(name)
    rp, ep = str(tmp_path / "ref.nwk"), str(tmp_path / "eval.nwk")
    open(rp, "w").write(g["ref_newick"].strip() + "\n")
    open(ep, "w").write(g["eval_newick"].strip() + "\n")
    outs = []
    for exe, tag in ((REF_BIN, "ref"), (OUR_BIN, "our")):
        op = str(tmp_path / f"{tag}.out.nwk")
        p = subprocess.run([exe, "-r", rp, "-e", ep, "-o", op, "-s", "-t", "1"], capture_output=True, text=True)
        assert p.returncode == 0, p.stdout + p.stderr
        outs.append(open(op).read())
    assert outs[0] == outs[1]
    assert outs[1].strip() == g["out_newick_s"].strip()


@needs_bins
def test_cli_refuses_existing_output(golden, tmp_path):
    g = golden("c1_known_answer")
    rp, ep, op = str(tmp_path / "r.nwk"), str(tmp_path / "e.nwk"), str(tmp_path / "o.nwk")
    open(rp, "w").write(g["ref_newick"].strip() + "\n")
    open(ep, "w").write(g["eval_newick"].strip() + "\n")
    open(op, "w").write("x")
    p = subprocess.run([OUR_BIN, "-r", rp, "-e", ep, "-o", op], capture_output=True, text=True)
    assert p.returncode == 1 and "already exists" in p.stdout            # src/QuartetScores.cpp:81-85


@needs_bins
def test_cli_two_shards_in_one_process(golden, tmp_path):
    """QS_NUM_GPUS=2 on a box with one GPU is refused by qs_create (device 1 missing); with enough GPUs it must agree."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    g = golden("s32x270_spr")
    one, _, _ = _run(OUR_BIN, tmp_path, "g1", g["ref_newick"], g["eval_newick"], [], env={"QS_NUM_GPUS": "1"})
    rp, ep, op = str(tmp_path / "r.nwk"), str(tmp_path / "e.nwk"), str(tmp_path / "o2.nwk")
    open(rp, "w").write(g["ref_newick"].strip() + "\n")
    open(ep, "w").write(g["eval_newick"].strip() + "\n")
    p = subprocess.run([OUR_BIN, "-r", rp, "-e", ep, "-o", op], capture_output=True, text=True, env=dict(os.environ, QS_NUM_GPUS="2"))
    assert p.returncode == 0, p.stdout + p.stderr
    assert open(op).read() == one


# ---- round 2: the drop-in semantics the reference has and round 1 lacked (VERDICT r01 "What's missing" 1-2) --------------

def _files(g, tmp_path):
    rp, ep = str(tmp_path / "ref.nwk"), str(tmp_path / "eval.nwk")
    open(rp, "w").write(g["ref_newick"].strip() + "\n")
    open(ep, "w").write(g["eval_newick"].strip() + "\n")
    return rp, ep


@needs_bins
@pytest.mark.parametrize("name", ["c1_known_answer", "s16x300_missing_poly", "s20x40_multiref_missing"])
def test_cli_savemem_with_raw_qic(name, golden, tmp_path):
    """-s together with -q (BASELINE config 5 names it; src/QuartetScores.cpp:120-122 writes the raw file with either table)"""
    g = golden(name)
    rp, ep = _files(g, tmp_path)
    got = {}
    for exe, tag in ((REF_BIN, "ref"), (OUR_BIN, "our")):
        op, qp = str(tmp_path / f"{tag}.out.nwk"), str(tmp_path / f"{tag}.raw.txt")
        p = subprocess.run([exe, "-r", rp, "-e", ep, "-o", op, "-q", qp, "-s", "-t", "1"], capture_output=True, text=True, env=dict(os.environ, QS_SLAB_BYTES="20000"))
        assert p.returncode == 0, p.stdout + p.stderr
        got[tag] = (open(op).read(), open(qp).read(), p.stdout)
    assert got["our"][0] == got["ref"][0] and got["our"][1] == got["ref"][1]
    assert "Using memory-efficient Lookup table" in got["our"][2] and "Using memory-efficient Lookup table" in got["ref"][2]


@needs_bins
def test_cli_memory_policy_lines_and_automatic_table_free_fallback(golden, tmp_path):
    """QuartetScoreComputer.hpp:724-745: the same estimate lines and table choice as the reference on this host; and when the
    table does not fit the DEVICE (forced here) the run falls back to table-free slabs instead of failing — same outputs, -q too."""
    g = golden("s32x270_spr")
    want_out, want_raw, ref_stdout = _run(REF_BIN, tmp_path, "ref", g["ref_newick"], g["eval_newick"], ["-t", "2"])
    got_out, got_raw, stdout = _run(OUR_BIN, tmp_path, "our", g["ref_newick"], g["eval_newick"], ["-t", "2"],
                                    env={"QS_TABLE_BYTES_LIMIT": "1000", "QS_SLAB_BYTES": "40000"})
    assert got_out == want_out and got_raw == want_raw
    pick = lambda text: [l for l in text.splitlines() if "Lookup table" in l or "Estimated" in l]
    assert pick(stdout) == pick(ref_stdout) and any("Using runtime-efficient Lookup table" in l for l in pick(stdout))


@needs_bins
@pytest.mark.parametrize("extra", [[], ["-s"]])
def test_cli_several_shards_with_raw_qic(extra, golden, tmp_path):
    """three shards in one process (all on GPU 0 here; one per GPU in production): same annotated tree, same -q file"""
    g = golden("s16x300_missing_poly")
    want_out, want_raw, _ = _run(REF_BIN, tmp_path, "ref", g["ref_newick"], g["eval_newick"], ["-t", "1"] + extra)
    got_out, got_raw, _ = _run(OUR_BIN, tmp_path, "our", g["ref_newick"], g["eval_newick"], ["-t", "2"] + extra, env={"QS_DEVICES": "0,0,0", "QS_SLAB_BYTES": "20000"})
    assert got_out == want_out and got_raw == want_raw
