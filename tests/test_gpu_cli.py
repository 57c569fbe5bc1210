"""The `locarna`-compatible command line front end against stdout / --clustal output of the reference's own binary.

Every case is one process of the front end; the commands of the whole module are started together, a few at a time
(golden_util.prefetch), and each test then checks the result of its own command."""
import json
import os
import subprocess

import pytest

from golden_util import GOLD, out_dir, prefetch, run

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "locarna_b200", "bin", "locarna_b200")
CLI_P = os.path.join(ROOT, "locarna_b200", "bin", "locarna_p_b200")
CLI_TREE = os.path.join(ROOT, "locarna_b200", "bin", "mlocarna_tree_b200")


def _load(name):
    return json.load(open(os.path.join(GOLD, name)))


def _f(tag, i, ext):
    """Output file of case i of the test family `tag` (fixed name: the command is started before the test runs)."""
    return os.path.join(out_dir(), "%s%d.%s" % (tag, i, ext))


CASES = list(enumerate(_load("reference_outputs.json")["cli"]))
CASES_OPT = list(enumerate(_load("locarna_cli_options.json")))
CASES_KBEST = list(enumerate(_load("kbest_outputs.json")))
CASES_NORM = list(enumerate(_load("normalized_outputs.json")))
CASES_P = list(enumerate(_load("locarna_p_cli.json")))
CASES_REFALN = list(enumerate(_load("maxdiffaln_outputs.json")))
CASES_RIBOSUM = list(enumerate(_load("ribosum_outputs.json")))
CASES_PP = list(enumerate(_load("pp_outputs.json")))


def _ab(case):
    return [os.path.join(GOLD, case["A"]), os.path.join(GOLD, case["B"])]


# the commands of each test family: (argv, cwd)
def _cmd_plain(i, c):
    return ([CLI] + _ab(c) + ["--clustal", _f("plain", i, "aln")] + c["args"], None)


def _cmd_opt(i, c):
    return ([CLI] + _ab(c) + ["--clustal", _f("opt", i, "aln"), "--stockholm", _f("opt", i, "sto")] + c["args"], None)


def _cmd_opt_ams(i, c):
    return ([CLI] + _ab(c) + ["--write-arcmatch-scores", _f("opt", i, "ams")] + c["args"], None)


def _cmd_kbest(i, c):
    return ([CLI] + _ab(c) + c["args"], None)


def _cmd_norm(i, c):
    return ([CLI] + _ab(c) + ["--clustal", _f("norm", i, "aln")] + c["args"], None)


def _cmd_p(i, c):
    return ([CLI_P] + _ab(c) + ["--write-arcmatch-probs", _f("p", i, "am"), "--write-basematch-probs", _f("p", i, "bm")] + c["args"], None)


def _cmd_refaln(i, c):
    return ([CLI, c["A"], c["B"]] + c["args"], GOLD)


def _cmd_ribosum(i, c):
    return ([CLI, c["A"], c["B"]] + c["args"], GOLD)


def _cmd_ribosum_ams(i, c):
    return ([CLI, c["A"], c["B"], "--write-arcmatch-scores", _f("ribosum", i, "ams")] + c["args"], GOLD)


def _cmd_pp(i, c):
    return ([CLI, c["A"], c["B"], "--pp", _f("pp", i, "pp"), "-q"] + c["args"], GOLD)


@pytest.fixture(scope="module", autouse=True)
def _all_commands_started():
    jobs = []
    for fn, cases in ((_cmd_plain, CASES), (_cmd_opt, CASES_OPT), (_cmd_opt_ams, CASES_OPT), (_cmd_kbest, CASES_KBEST), (_cmd_norm, CASES_NORM),
                      (_cmd_p, CASES_P), (_cmd_refaln, CASES_REFALN), (_cmd_ribosum, CASES_RIBOSUM), (_cmd_pp, CASES_PP)):
        jobs += [fn(i, c) for i, c in cases]
    jobs += [_cmd_ribosum_ams(i, c) for i, c in CASES_RIBOSUM if c["rc"] == 0]
    prefetch(jobs)


def _ids(cases, cut=None):
    return ["%s-%s" % ("_".join((x[:cut] if cut else x) for x in c["args"]) or "default", c["A"]) for _, c in cases]


@pytest.mark.parametrize("i,case", CASES, ids=_ids(CASES))
def test_cli_output_matches_reference(i, case):
    r = run(*_cmd_plain(i, case))
    assert r.returncode == case["rc"], r.stderr
    assert r.stdout == case["stdout"]
    assert open(_f("plain", i, "aln")).read() == case["clustal"]


@pytest.mark.parametrize("i,case", CASES_OPT, ids=_ids(CASES_OPT))
def test_cli_exp_prob_maxbpspan_arcmatch_scores(i, case):
    """--exp-prob, --maxBPspan and --write-arcmatch-scores against the reference binary (tools/make_golden_cli_options.py)."""
    r = run(*_cmd_opt(i, case))
    assert r.returncode == case["rc"], r.stderr
    assert r.stdout == case["stdout"]
    assert open(_f("opt", i, "aln")).read() == case["clustal"]
    assert open(_f("opt", i, "sto")).read() == case["stockholm"]
    w = run(*_cmd_opt_ams(i, case))
    assert w.returncode == case["ams_rc"], w.stderr
    assert w.stdout == case["ams_stdout"]                      # writes the file and exits without aligning (locarna.cc:705-720)
    assert open(_f("opt", i, "ams")).read() == case["arcmatch_scores"]


def test_cli_rejects_unimplemented_modes():
    r = subprocess.run([CLI, os.path.join(GOLD, "g0.pp"), os.path.join(GOLD, "g1.pp"), "--mea-alignment"], capture_output=True, text=True)
    assert r.returncode == 255 and "does not implement" in r.stderr


@pytest.mark.parametrize("i,case", CASES_KBEST, ids=_ids(CASES_KBEST))
def test_cli_kbest(i, case):
    """--kbest k / --better t: k-best alignments by interval splitting (Aligner::suboptimal, aligner.cc:1383-1514; restricted top levels
    on the resident D table) against the reference binary's stdout (tools/make_golden_kbest.py)."""
    r = run(*_cmd_kbest(i, case))
    assert r.returncode == case["rc"], r.stderr
    assert r.stdout == case["stdout"]


@pytest.mark.parametrize("i,case", CASES_NORM, ids=_ids(CASES_NORM))
def test_cli_normalized_penalized(i, case):
    """--normalized L (Dinkelbach iteration, aligner.cc:1522-1597) and --penalized PP (aligner.cc:1599-1622) against the reference binary:
    stdout (score + alignment), stderr of the rejected combinations and the clustal file (tools/make_golden_normalized.py)."""
    r = run(*_cmd_norm(i, case))
    assert r.returncode == case["rc"], r.stderr
    assert r.stdout == case["stdout"]
    if case["rc"] != 0:
        assert r.stderr == case["stderr"]
    else:
        assert open(_f("norm", i, "aln")).read() == case["clustal"]


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


@pytest.mark.parametrize("i,case", CASES_P, ids=_ids(CASES_P))
def test_locarna_p_cli_matches_reference(i, case):
    """stdout and the probability files of the reference's own locarna_p binary (tests/golden/locarna_p_cli.json)."""
    r = run(*_cmd_p(i, case))
    assert r.returncode == case["rc"], r.stderr
    assert r.stdout == case["stdout"]
    assert _close_lines(open(_f("p", i, "am")).read(), case["am"], 4)
    assert _close_lines(open(_f("p", i, "bm")).read(), case["bm"], 2)


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


@pytest.mark.parametrize("i,case", CASES_REFALN, ids=_ids(CASES_REFALN, 12))
def test_cli_reference_alignment_bands(i, case):
    """--max-diff d with --max-diff-pw-aln / --max-diff-aln: band around a reference alignment (TraceController from a MultipleAlignment,
    trace_controller.cc:406-539), then the probability envelope inside it; stdout / error exits of the reference binary
    (tools/make_golden_maxdiffaln.py)."""
    r = run(*_cmd_refaln(i, case))
    assert r.returncode == case["rc"], r.stderr
    assert r.stdout == case["stdout"]
    if case["rc"] != 0:
        assert r.stderr == case["stderr"]


@pytest.mark.parametrize("i,case", CASES_RIBOSUM, ids=["%s-%s" % ("_".join(c["args"][1:]), c["A"]) for _, c in CASES_RIBOSUM])
def test_cli_ribosum_file(i, case):
    """--ribosum-file with a matrix other than the built-in one (tests/golden/synthetic.ribosum, tools/make_golden_ribosum.py): base-match
    and arc-match score tables from the file (RibosumFreq, ribosum.cc:40-200, :324-331; scoring.cc:141-198, :369-438), stdout and the
    arc-match scores of the reference binary; a file that is not a ribosum matrix is refused with the reference's message."""
    r = run(*_cmd_ribosum(i, case))
    assert r.returncode == case["rc"], r.stderr
    assert r.stdout == case["stdout"]
    if case["rc"] != 0:
        assert r.stderr == case["stderr"]
        return
    w = run(*_cmd_ribosum_ams(i, case))
    assert w.returncode == 0, w.stderr
    assert open(_f("ribosum", i, "ams")).read() == case["arcmatch_scores"]


@pytest.mark.parametrize("i,case", CASES_PP, ids=_ids(CASES_PP))
def test_cli_pp_output(i, case):
    """--pp: the alignment with its consensus dot plot in PP 2.0 format (the hand-over to mlocarna's progressive stage; consensus
    constructor of RnaData rna_data.cc:104-126 / :1474-1578, write_pp :1242-1350), byte-equal to the reference binary's file incl. the
    order of the base pair lines (tools/make_golden_pp.py)."""
    r = run(*_cmd_pp(i, case))
    assert r.returncode == case["rc"], r.stderr
    assert open(_f("pp", i, "pp")).read() == case["pp"]
