"""UPGMA guide tree against newick strings produced by the reference's own Perl module (tools/make_golden.py)."""
import json
import os

from golden_util import GOLD
from locarna_b200 import capi


def test_upgma_matches_reference_perl():
    trees = json.load(open(os.path.join(GOLD, "reference_outputs.json")))["trees"]
    assert len(trees) >= 4
    for t in trees:
        assert capi.upgma_newick(t["names"], t["matrix"]) == t["newick"], t["names"]
