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
