"""--stacking / --new-stacking on the GPU against the compiled reference (tests/golden/stacking_outputs.json from tools/make_golden_stacking.py):
score, the complete D table in the reference's arc-match order (with Scoring::arcmatch of every arc match) and the alignment edges,
through every D-fill kernel."""
import json
import os

import pytest

from golden_util import GOLD, digest, full_edges
from locarna_b200 import capi

pytestmark = pytest.mark.gpu
CASES = json.load(open(os.path.join(GOLD, "stacking_outputs.json")))


@pytest.mark.parametrize("mode", ["auto", "dep", "levels"])
@pytest.mark.parametrize("case", CASES, ids=lambda c: "%s-%s" % ("_".join("%s=%s" % kv for kv in c["flags"].items()), c["A"]))
def test_stacking_matches_reference(case, mode, monkeypatch):
    monkeypatch.setenv("LB200_DFILL", mode)
    assert case["D_differs_from_unstacked"] > 0                     # the fixture exercises the stacked terms
    ctx = capi.Context(0, case["flags"])
    a, b = ctx.add_pp(os.path.join(GOLD, case["A"])), ctx.add_pp(os.path.join(GOLD, case["B"]))
    ctx.add_pair(a, b)
    ctx.run(capi.RUN_TRACE | capi.RUN_KEEP_D)
    assert ctx.scores()[0] == case["score"]
    am, score, D = ctx.arcmatches(0, with_D=True)
    rows = [list(x) + [s, d] for x, s, d in zip(am, score, D)]
    assert len(rows) == case["n_am"] and rows[:8] == case["am_head"]
    assert digest(rows) == case["am_sha256"]
    edges, sa, sb = ctx.alignment(0)
    inf = ctx.info(0)
    assert full_edges(edges, inf.lenA, inf.lenB) == case["edges_full"]
    ctx.close()
