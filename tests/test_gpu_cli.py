"""The `locarna`-compatible command line front end against stdout / --clustal output of the reference's own binary."""
import json
import os
import subprocess

import pytest

from golden_util import GOLD

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "locarna_b200", "bin", "locarna_b200")
CASES = json.load(open(os.path.join(GOLD, "reference_outputs.json")))["cli"]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%s-%s" % ("_".join(c["args"]) or "default", c["A"]))
def test_cli_output_matches_reference(case, tmp_path):
    clu = str(tmp_path / "out.aln")
    r = subprocess.run([CLI, os.path.join(GOLD, case["A"]), os.path.join(GOLD, case["B"]), "--clustal", clu] + case["args"],
                       capture_output=True, text=True)
    assert r.returncode == case["rc"], r.stderr
    assert r.stdout == case["stdout"]
    assert open(clu).read() == case["clustal"]


CASES_OPT = json.load(open(os.path.join(GOLD, "locarna_cli_options.json")))


@pytest.mark.parametrize("case", CASES_OPT, ids=lambda c: "%s-%s" % ("_".join(c["args"]) or "default", c["A"]))
def test_cli_exp_prob_maxbpspan_arcmatch_scores(case, tmp_path):
    """--exp-prob, --maxBPspan and --write-arcmatch-scores against the reference binary (tools/make_golden_cli_options.py)."""
    clu, ams = str(tmp_path / "out.aln"), str(tmp_path / "out.ams")
    a, b = os.path.join(GOLD, case["A"]), os.path.join(GOLD, case["B"])
    sto = str(tmp_path / "out.sto")
    r = subprocess.run([CLI, a, b, "--clustal", clu, "--stockholm", sto] + case["args"], capture_output=True, text=True)
    assert r.returncode == case["rc"], r.stderr
    assert r.stdout == case["stdout"]
    assert open(clu).read() == case["clustal"]
    assert open(sto).read() == case["stockholm"]
    w = subprocess.run([CLI, a, b, "--write-arcmatch-scores", ams] + case["args"], capture_output=True, text=True)
    assert w.returncode == case["ams_rc"], w.stderr
    assert w.stdout == case["ams_stdout"]                      # writes the file and exits without aligning (locarna.cc:705-720)
    assert open(ams).read() == case["arcmatch_scores"]


def test_cli_rejects_unimplemented_modes():
    r = subprocess.run([CLI, os.path.join(GOLD, "g0.pp"), os.path.join(GOLD, "g1.pp"), "--mea-alignment"], capture_output=True, text=True)
    assert r.returncode == 255 and "does not implement" in r.stderr


CASES_KBEST = json.load(open(os.path.join(GOLD, "kbest_outputs.json")))


@pytest.mark.parametrize("case", CASES_KBEST, ids=lambda c: "%s-%s" % ("_".join(c["args"]), c["A"]))
def test_cli_kbest(case):
    """--kbest k / --better t: k-best alignments by interval splitting (Aligner::suboptimal, aligner.cc:1383-1514; restricted top levels
    on the resident D table) against the reference binary's stdout (tools/make_golden_kbest.py)."""
    r = subprocess.run([CLI, os.path.join(GOLD, case["A"]), os.path.join(GOLD, case["B"])] + case["args"], capture_output=True, text=True)
    assert r.returncode == case["rc"], r.stderr
    assert r.stdout == case["stdout"]


CASES_NORM = json.load(open(os.path.join(GOLD, "normalized_outputs.json")))


@pytest.mark.parametrize("case", CASES_NORM, ids=lambda c: "%s-%s" % ("_".join(c["args"]), c["A"]))
def test_cli_normalized_penalized(case, tmp_path):
    """--normalized L (Dinkelbach iteration, aligner.cc:1522-1597) and --penalized PP (aligner.cc:1599-1622) against the reference binary:
    stdout (score + alignment), stderr of the rejected combinations and the clustal file (tools/make_golden_normalized.py)."""
    clu = str(tmp_path / "out.aln")
    r = subprocess.run([CLI, os.path.join(GOLD, case["A"]), os.path.join(GOLD, case["B"]), "--clustal", clu] + case["args"], capture_output=True, text=True)
    assert r.returncode == case["rc"], r.stderr
    assert r.stdout == case["stdout"]
    if case["rc"] != 0:
        assert r.stderr == case["stderr"]
    else:
        assert open(clu).read() == case["clustal"]


CLI_P = os.path.join(ROOT, "locarna_b200", "bin", "locarna_p_b200")
CASES_P = json.load(open(os.path.join(GOLD, "locarna_p_cli.json")))


def _close_lines(a: str, b: str, n_int: int) -> bool:
    """Same lines; the leading n_int integer fields identical, the probability equal to the 6 printed digits (+- one unit)."""
    la, lb = a.splitlines(), b.splitlines()
    if len(la) != len(lb):
        return False
    for x, y in zip(la, lb):
        fx, fy = x.split(), y.split()
        if fx[:n_int] != fy[:n_int] or abs(float(fx[n_int]) - float(fy[n_int])) > 2e-6 * abs(float(fy[n_int])):
            return False
    return True


@pytest.mark.parametrize("case", CASES_P, ids=lambda c: "%s-%s" % ("_".join(c["args"]) or "default", c["A"]))
def test_locarna_p_cli_matches_reference(case, tmp_path):
    """stdout and the probability files of the reference's own locarna_p binary (tests/golden/locarna_p_cli.json)."""
    am, bm = str(tmp_path / "am"), str(tmp_path / "bm")
    r = subprocess.run([CLI_P, os.path.join(GOLD, case["A"]), os.path.join(GOLD, case["B"]), "--write-arcmatch-probs", am, "--write-basematch-probs", bm]
                       + case["args"], capture_output=True, text=True)
    assert r.returncode == case["rc"], r.stderr
    assert r.stdout == case["stdout"]
    assert _close_lines(open(am).read(), case["am"], 4)
    assert _close_lines(open(bm).read(), case["bm"], 2)


CLI_TREE = os.path.join(ROOT, "locarna_b200", "bin", "mlocarna_tree_b200")


def test_mlocarna_tree_stage_archaea(tmp_path):
    """BASELINE config 1 through the one-process guide-tree front end: the 21 pairwise scores of the reference's `locarna` binary
    (mlocarna's tree-stage flags), mlocarna's result.matrix format and the UPGMA tree of the reference's Perl module."""
    from locarna_b200 import allpairs
    gold = json.load(open(os.path.join(GOLD, "reference_outputs.json")))
    arch, tree = gold["archaea"], gold["trees"][0]
    assert tree["names"] == arch["names"]
    (tmp_path / "results").mkdir(); (tmp_path / "scores").mkdir()
    files = [os.path.join(GOLD, "archaea", n + ".pp") for n in arch["names"]]
    r = subprocess.run([CLI_TREE, "--tgtdir", str(tmp_path)] + files, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(tmp_path / "results" / "result.matrix").read() == allpairs.format_matrix(tree["matrix"])
    assert open(tmp_path / "results" / "result.tree").read() == tree["newick"] + ";\n"
    pairs = [tuple(p) for p in arch["pairs"]]
    assert pairs == allpairs.all_vs_all(len(files))
    assert open(tmp_path / "scores" / "scores-0").read() == allpairs.format_score_list(pairs, arch["scores"])
    # explicit flags equal to the defaults, results on stdout
    r2 = subprocess.run([CLI_TREE, "--struct-weight", "200", "-D", "30", "--noLP", "-p", "0.001"] + files, capture_output=True, text=True)
    assert r2.returncode == 0 and r2.stdout == allpairs.format_matrix(tree["matrix"]) + tree["newick"] + ";\n"


def test_mlocarna_tree_stage_shares_and_gpus(tmp_path):
    """The pair list split mlocarna's way over processes (--compute-pairwise-scores k/N, src/Utils/mlocarna:2321-2344; cost-balanced
    shares) and over devices (--gpus): the partial score lists together are exactly the full list, and --gpus 1 equals the default."""
    gold = json.load(open(os.path.join(GOLD, "reference_outputs.json")))
    arch = gold["archaea"]
    files = [os.path.join(GOLD, "archaea", n + ".pp") for n in arch["names"]]
    want = {(a, b): s for (a, b), s in zip([tuple(p) for p in arch["pairs"]], arch["scores"])}
    got = {}
    for k in (1, 2, 3):
        out = tmp_path / ("scores-%d" % k)
        r = subprocess.run([CLI_TREE, "--compute-pairwise-scores", "%d/3" % k, "--score-list", str(out)] + files, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        for line in open(out):
            a, b, s = line.split()
            assert (int(a), int(b)) not in got
            got[(int(a), int(b))] = int(s)
    assert got == want
    r1 = subprocess.run([CLI_TREE, "--gpus", "1"] + files, capture_output=True, text=True)
    r0 = subprocess.run([CLI_TREE] + files, capture_output=True, text=True)
    assert r1.returncode == 0 and r1.stdout == r0.stdout and r0.stdout


CASES_REFALN = json.load(open(os.path.join(GOLD, "maxdiffaln_outputs.json")))


@pytest.mark.parametrize("case", CASES_REFALN, ids=lambda c: "%s-%s" % ("_".join(x[:12] for x in c["args"]), c["A"]))
def test_cli_reference_alignment_bands(case):
    """--max-diff d with --max-diff-pw-aln / --max-diff-aln: band around a reference alignment (TraceController from a MultipleAlignment,
    trace_controller.cc:406-539), then the probability envelope inside it; stdout / error exits of the reference binary
    (tools/make_golden_maxdiffaln.py)."""
    r = subprocess.run([CLI, case["A"], case["B"]] + case["args"], capture_output=True, text=True, cwd=GOLD)
    assert r.returncode == case["rc"], r.stderr
    assert r.stdout == case["stdout"]
    if case["rc"] != 0:
        assert r.stderr == case["stderr"]


CASES_RIBOSUM = json.load(open(os.path.join(GOLD, "ribosum_outputs.json")))


@pytest.mark.parametrize("case", CASES_RIBOSUM, ids=lambda c: "%s-%s" % ("_".join(c["args"][1:]), c["A"]))
def test_cli_ribosum_file(case, tmp_path):
    """--ribosum-file with a matrix other than the built-in one (tests/golden/synthetic.ribosum, tools/make_golden_ribosum.py): base-match
    and arc-match score tables from the file (RibosumFreq, ribosum.cc:40-200, :324-331; scoring.cc:141-198, :369-438), stdout and the
    arc-match scores of the reference binary; a file that is not a ribosum matrix is refused with the reference's message."""
    r = subprocess.run([CLI, case["A"], case["B"]] + case["args"], capture_output=True, text=True, cwd=GOLD)
    assert r.returncode == case["rc"], r.stderr
    assert r.stdout == case["stdout"]
    if case["rc"] != 0:
        assert r.stderr == case["stderr"]
        return
    ams = str(tmp_path / "out.ams")
    w = subprocess.run([CLI, case["A"], case["B"], "--write-arcmatch-scores", ams] + case["args"], capture_output=True, text=True, cwd=GOLD)
    assert w.returncode == 0, w.stderr
    assert open(ams).read() == case["arcmatch_scores"]


CASES_PP = json.load(open(os.path.join(GOLD, "pp_outputs.json")))


@pytest.mark.parametrize("case", CASES_PP, ids=lambda c: "%s-%s" % ("_".join(c["args"]) or "default", c["A"]))
def test_cli_pp_output(case, tmp_path):
    """--pp: the alignment with its consensus dot plot in PP 2.0 format (the hand-over to mlocarna's progressive stage; consensus
    constructor of RnaData rna_data.cc:104-126 / :1474-1578, write_pp :1242-1350), byte-equal to the reference binary's file incl. the
    order of the base pair lines (tools/make_golden_pp.py)."""
    pp = str(tmp_path / "out.pp")
    r = subprocess.run([CLI, case["A"], case["B"], "--pp", pp, "-q"] + case["args"], capture_output=True, text=True, cwd=GOLD)
    assert r.returncode == case["rc"], r.stderr
    assert open(pp).read() == case["pp"]
