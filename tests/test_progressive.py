"""Progressive alignment along the guide tree (locarna_b200/progressive.py = mlocarna's perform_progressive_steps over the pairwise
front end): tree parsing, the order / operands / names of the steps, and the complete run on the archaea example - with the reference's
own binary on the CPU (pins the driver) and with locarna_b200 on the GPU (every intermediate profile alignment and its consensus dot
plot, byte for byte, against tests/golden/progressive_archaea.json made by tools/make_golden_progressive.py)."""
import json
import os

import pytest

from golden_util import GOLD
from locarna_b200 import progressive as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASE = json.load(open(os.path.join(GOLD, "progressive_archaea.json")))
REF_LOCARNA = os.path.join(ROOT, "oracle", "_ref", "locarna")


def _leaf(label):
    return os.path.join(GOLD, "archaea", label + ".pp")


def test_newick_and_plan():
    t = P.parse_newick("((a:1.5,'b c':2)x:0.5,(d,e));")
    assert [c.label for c in t.children] == ["x", None] and [c.label for c in t.children[0].children] == ["a", "b c"]
    steps, final = P.plan(t, lambda l: l + ".pp", "im")
    # post order, first child first; names as lib/perl/MLocarna.pm:104-131 (label appended, non-alphanumerics -> "_", "-<i>" on a clash)
    assert [(s.op1, s.op2, s.target, s.size) for s in steps] == [("a.pp", "b c.pp", "im/intermediatex", 2), ("d.pp", "e.pp", "im/intermediate", 2),
                                                                 ("im/intermediatex.pp", "im/intermediate.pp", "im/intermediate-1", 4)]
    assert final == "im/intermediate-1"
    assert P.plan(t, lambda l: l, "im", max_alignment_size=3)[1] is None          # src/Utils/mlocarna:3684
    with pytest.raises(ValueError):
        P.plan(P.parse_newick("(a,b,c);"), lambda l: l, "im")                       # "the guide tree must be binary"
    cmd = P.command(steps[0], ["--noLP"], locarna="locarna")
    assert cmd == ["locarna", "a.pp", "b c.pp", "--noLP", "--clustal=im/intermediatex.aln", "--pp=im/intermediatex.pp", "-q"]   # mlocarna:2683-2699


def _check(d):
    assert open(os.path.join(d, "results", "result.aln")).read() == CASE["result_aln"]
    assert open(os.path.join(d, "results", "result.pp")).read() == CASE["result_pp"]
    for s in CASE["steps"]:
        assert open(os.path.join(d, "intermediates", s["name"] + ".aln")).read() == s["aln"], s["name"]
        assert open(os.path.join(d, "intermediates", s["name"] + ".pp")).read() == s["pp"], s["name"]


@pytest.mark.skipif(not os.access(REF_LOCARNA, os.X_OK), reason="the compiled reference is only present where oracle/_ref was built")
def test_driver_with_the_reference_binary(tmp_path):
    tree = P.parse_newick(CASE["newick"] + ";")
    cmds, final = P.run(tree, _leaf, str(tmp_path), CASE["args"], locarna=REF_LOCARNA)
    assert len(cmds) == 6 and os.path.basename(final) == CASE["final"]
    _check(str(tmp_path))


@pytest.mark.gpu
def test_progressive_alignment_equals_reference(tmp_path):
    tree = P.parse_newick(CASE["newick"] + ";")
    P.run(tree, _leaf, str(tmp_path), CASE["args"])
    _check(str(tmp_path))
