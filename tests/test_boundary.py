"""The drop-in boundary, proved by compilation: the body of run_and_report() of the reference's src/locarna.cc (RnaData, TraceController,
ArcMatches, ScoringParams / Scoring, AlignerParams / Aligner, align / trace / get_alignment) is cut from the reference tree at build time
and compiled UNMODIFIED against include/locarna_b200_compat.hh (Makefile target `refmain` -> locarna_b200/bin/locarna_refmain_b200).
On the GPU the program's stdout and --clustal output must equal the reference binary's (fixtures of tests/golden/)."""
import json
import os
import subprocess

import pytest

from golden_util import GOLD, out_dir, prefetch, run

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "locarna_b200", "bin", "locarna_refmain_b200")
REF = "/root/reference/src/locarna.cc"
CASES = json.load(open(os.path.join(GOLD, "reference_outputs.json")))["cli"][::2] + [
    c for c in json.load(open(os.path.join(GOLD, "locarna_cli_options.json")))][::2] + [   # every other case: tests/test_gpu_cli.py runs all of them through locarna_b200

    c for c in json.load(open(os.path.join(GOLD, "normalized_outputs.json")))[::3] if c["rc"] == 0 or "simultaneously" not in c["stderr"]] + [
    dict(c, clustal=None) for c in json.load(open(os.path.join(GOLD, "kbest_outputs.json")))[::3]] + [
    c for c in json.load(open(os.path.join(GOLD, "anchors_outputs.json")))[::4]] + [
    dict(c, clustal=None) for c in json.load(open(os.path.join(GOLD, "maxdiffaln_outputs.json")))[::6] if c["rc"] == 0] + [
    dict(c, clustal=None) for c in json.load(open(os.path.join(GOLD, "ribosum_outputs.json")))[::4] if c["rc"] == 0]   # every third case: the CLI tests run all of them


@pytest.mark.skipif(not os.path.exists(REF), reason="the reference tree is only present in the build container")
def test_reference_pipeline_compiles_against_the_compat_header():
    r = subprocess.run(["make", "-C", ROOT, "refmain"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    inc = open(os.path.join(ROOT, "build", "refmain", "refmain_block.inc")).read()
    # the block is the reference's text: a few of its statements, as written there
    for needle in ("std::make_unique<Aligner>(", "AlignerParams::trace_controller(&trace_controller)", "Scoring scoring(",
                   "ScoringParams::exp_probA(my_exp_probA)", "trace_controller.restrict_by_anchors(seq_constraints);", "score = aligner->align();"):
        assert needle in inc, needle
    assert len(inc.splitlines()) > 350
    assert os.access(BIN, os.X_OK)


def test_refmain_fails_loudly_without_a_device():
    """Host-side objects (RnaData, ScoringParams, ...) work anywhere; the aligner needs the GPU - no CPU fallback."""
    if not os.access(BIN, os.X_OK):
        pytest.skip("locarna_refmain_b200 not built (no reference tree at build time)")
    import ctypes
    try:
        ctypes.CDLL("libcuda.so.1")
        pytest.skip("a CUDA driver is present")
    except OSError:
        pass
    r = subprocess.run([BIN, os.path.join(GOLD, "g0.pp"), os.path.join(GOLD, "g1.pp")], capture_output=True, text=True)
    assert r.returncode == 255 and "no CPU fallback" in r.stderr
    r = subprocess.run([BIN, os.path.join(GOLD, "nonexistent.pp"), os.path.join(GOLD, "g1.pp")], capture_output=True, text=True)
    assert r.returncode == 255 and "failed to read from file" in r.stderr      # the reference's own error text (locarna.cc:470-476)


def _cmd(i, case):
    return ([BIN, os.path.join(GOLD, case["A"]), os.path.join(GOLD, case["B"]), "--clustal", os.path.join(out_dir(), "refmain%d.aln" % i)] + case["args"], GOLD)


@pytest.fixture(scope="module")
def commands_started():
    """All cases are started together, a few processes at a time (golden_util.prefetch)."""
    if os.access(BIN, os.X_OK):
        prefetch([_cmd(i, c) for i, c in enumerate(CASES)])


@pytest.mark.gpu
@pytest.mark.parametrize("i,case", list(enumerate(CASES)), ids=["%s-%s" % ("_".join(c["args"]) or "default", c["A"]) for c in CASES])
def test_refmain_output_matches_reference_binary(i, case, commands_started):
    if not os.access(BIN, os.X_OK):
        pytest.skip("locarna_refmain_b200 not built (no reference tree at build time)")
    r = run(*_cmd(i, case))
    assert r.returncode == case["rc"], r.stderr
    assert r.stdout == case["stdout"]
    if case["rc"] == 0 and case["clustal"] is not None:
        assert open(os.path.join(out_dir(), "refmain%d.aln" % i)).read() == case["clustal"]
