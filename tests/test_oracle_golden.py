"""The oracle port against fixtures produced by the compiled reference itself (tools/make_golden.py)."""
import os

import pytest

from golden_util import GOLD, digest, full_edges, load_cases
from oracle import oracle as O

CASES = load_cases()


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%s-%s-%s" % (c["case"], c["A"], c["B"]))
def test_port_matches_reference_fixture(case):
    a, b = os.path.join(GOLD, case["A"]), os.path.join(GOLD, case["B"])
    r = O.port_align(a, b, case["flags"])
    assert r["score"] == case["score"]
    assert r["min_col"] == case["min_col"] and r["max_col"] == case["max_col"]
    am_rows = [list(x[:4]) + [s, d] for x, s, d in zip(r["am"], r["am_score"], r["D"])]
    assert len(am_rows) == case["n_am"]
    assert am_rows[:12] == case["am_head"]
    assert digest(am_rows) == case["am_sha256"]
    assert full_edges(r["edges"], r["lenA"], r["lenB"]) == case["edges_full"]
    assert r["structA"] == case["structA"] and r["structB"] == case["structB"]
    assert r["rowA"] == case["rowA"] and r["rowB"] == case["rowB"]


@pytest.mark.skipif(not O.have_ref(), reason="compiled reference (oracle/_ref) not present")
def test_port_matches_live_reference(synth_dir):
    """Where the compiled reference exists, compare the port with it live on a fresh input as well."""
    a, b = synth_dir["cfg3"][:2]
    for flags in ({}, {"noLP": True, "max-diff-am": 30}, {"sequ-local": True}):
        p, r = O.port_align(a, b, flags), O.ref_align(a, b, flags)
        for k in ("score", "min_col", "max_col", "am", "am_score", "D", "rowA", "rowB", "structA", "structB"):
            assert p[k] == r[k], (flags, k)


def test_archaea_all_vs_all_scores():
    """BASELINE config 1: the 21 pairwise scores of Data/Examples/archaea.fa (synthetic dot plots), mlocarna tree-stage flags."""
    import json
    g = json.load(open(os.path.join(GOLD, "reference_outputs.json")))["archaea"]
    paths = [os.path.join(GOLD, "archaea", n + ".pp") for n in g["names"]]
    for (a, b), score, rowA, rowB in zip(g["pairs"], g["scores"], g["rowA"], g["rowB"]):
        r = O.port_align(paths[a], paths[b], g["flags"])
        assert r["score"] == score
        assert r["rowA"] == rowA and r["rowB"] == rowB


def test_port_inside_p_matches_reference_fixture():
    """LocARNA-P inside: the port's Z and inside table against the compiled reference's AlignerP<double>
    (tests/golden/locarna_p_outputs.json, tools/make_golden_p.py). The port follows the reference's expression order, so the
    values agree to the last bit here; the test allows 1e-12 relative for other libm builds."""
    import json
    for case in json.load(open(os.path.join(GOLD, "locarna_p_outputs.json"))):
        r = O.port_inside_p(os.path.join(GOLD, case["A"]), os.path.join(GOLD, case["B"]), case["flags"], case["pf_scale"])
        assert abs(r["Z"] - case["Z"]) <= 1e-12 * abs(case["Z"]), case["A"]
        assert len(r["D"]) == len(case["D"])
        assert all(abs(x - y) <= 1e-12 * abs(y) for x, y in zip(r["D"], case["D"]))


@pytest.mark.skipif(not O.have_ref(), reason="compiled reference (oracle/_ref) not present")
def test_port_inside_p_matches_live_reference(synth_dir):
    a, b = synth_dir["cfg3"][:2]
    for flags in ({"pf-double": True, "min-trace-probability": 1e-5}, {"pf-double": True, "min-trace-probability": 1e-5, "max-diff-am": 20}):
        p, r = O.port_inside_p(a, b, flags), O.ref_inside_p(a, b, flags)
        assert p["Z"] == r["Z"] and p["D"] == r["pfD"], flags


def test_port_probs_p_matches_reference_fixture():
    """LocARNA-P outside + probabilities: the port's arc-match and base-match probabilities against the compiled reference's
    (the lists locarna_p writes with --write-arcmatch-probs / --write-basematch-probs, thresholds 0.001)."""
    import json
    for case in json.load(open(os.path.join(GOLD, "locarna_p_outputs.json"))):
        r = O.port_probs_p(os.path.join(GOLD, case["A"]), os.path.join(GOLD, case["B"]), case["flags"], case["pf_scale"])
        am = {tuple(case["am"][k]): r["am_prob"][k] for k in range(len(case["am"])) if r["am_prob"][k] >= 0.001}
        ref_am = {tuple(x[:4]): x[4] for x in case["am_probs"]}
        assert set(am) == set(ref_am), case["A"]
        assert all(abs(am[k] - ref_am[k]) <= 1e-12 * ref_am[k] for k in ref_am)
        bm = {(i, j): r["bm"][i][j] for i in range(1, r["lenA"] + 1) for j in range(1, r["lenB"] + 1) if r["bm"][i][j] >= 0.001}
        ref_bm = {(x[0], x[1]): x[2] for x in case["bm_probs"]}
        assert set(bm) == set(ref_bm), case["A"]
        assert all(abs(bm[k] - ref_bm[k]) <= 1e-12 * ref_bm[k] for k in ref_bm)


@pytest.mark.parametrize("pair", [("g0.pp", "g1.pp"), ("g2.pp", "g3.pp"), ("st0.pp", "st1.pp")])
@pytest.mark.parametrize("flags", [{}, {"noLP": True}, {"indel-opening": 0}, {"max-diff": 12}, {"no-ribosum": True, "indel-opening": -300}, {"tau": 0}])
def test_gap_free_frame_is_score_and_trace_preserving(pair, flags, monkeypatch):
    """The product runs profile pairs in the gap-free frame (locarna_b200/csrc/host_model.h ProfileTables: gap costs moved into the base
    match and arc-match scores, gap extension 0). The restatement, run with position-dependent gap costs once as the reference computes
    and once in that frame (Scoring::to_gap_free_frame), must give the same score, the same D for every arc match and the same trace."""
    import os
    from oracle import oracle as O
    from golden_util import GOLD
    a, b = os.path.join(GOLD, pair[0]), os.path.join(GOLD, pair[1])
    monkeypatch.setenv("LOCARNA_PORT_TEST_GAPS", "1")
    plain = O.port_align(a, b, flags)
    monkeypatch.setenv("LOCARNA_PORT_GAP_FREE_FRAME", "1")
    framed = O.port_align(a, b, flags)
    assert plain["am_score"] == framed["am_score"]
    assert plain["score"] == framed["score"] and plain["D"] == framed["D"] and plain["edges"] == framed["edges"]


PROFILE_CASES = __import__("json").load(open(os.path.join(GOLD, "profiles_outputs.json")))


@pytest.mark.parametrize("case", PROFILE_CASES, ids=["%s-%s-%s" % (c["A"], c["B"], "_".join(c["args"]) or "default") for c in PROFILE_CASES])
def test_port_profile_input_matches_reference_fixture(case):
    """Profile (multi-row) input: the port's scoring of alignment columns (scoring.cc:141-198, :272-311, :369-438; stral_score.cc:29-44)
    against the compiled reference (tests/golden/profiles_outputs.json, tools/make_golden_profiles.py): band, arc matches with scores,
    every D, score and alignment edges."""
    r = O.port_align(os.path.join(GOLD, case["A"]), os.path.join(GOLD, case["B"]), case["flags"])
    assert r["min_col"] == case["min_col"] and r["max_col"] == case["max_col"]
    am_rows = [list(x[:4]) + [s, d] for x, s, d in zip(r["am"], r["am_score"], r["D"])]
    assert len(am_rows) == case["n_am"] and am_rows[:5] == case["am_head"]
    assert digest(am_rows) == case["am_sha256"]
    assert r["score"] == case["score"]
    assert full_edges(r["edges"], r["lenA"], r["lenB"]) == case["edges_full"]
