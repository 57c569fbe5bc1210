"""The CUDA path (through the C ABI) against fixtures produced by the compiled reference itself."""
import os

import pytest

from golden_util import GOLD, digest, full_edges, load_cases
from locarna_b200 import capi

pytestmark = pytest.mark.gpu
CASES = load_cases()


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%s-%s-%s" % (c["case"], c["A"], c["B"]))
def test_gpu_matches_reference_fixture(case):
    ctx = capi.Context(0, case["flags"])
    ia, ib = ctx.add_pp(os.path.join(GOLD, case["A"])), ctx.add_pp(os.path.join(GOLD, case["B"]))
    ctx.add_pair(ia, ib)
    ctx.run(capi.RUN_TRACE)
    assert ctx.scores()[0] == case["score"]
    lo, hi = ctx.band(0)
    assert lo == case["min_col"] and hi == case["max_col"]
    am, score, D = ctx.arcmatches(0, with_D=True)
    am_rows = [list(x) + [s, d] for x, s, d in zip(am, score, D)]
    assert len(am_rows) == case["n_am"]
    assert digest(am_rows) == case["am_sha256"]
    edges, sa, sb = ctx.alignment(0)
    inf = ctx.info(0)
    assert full_edges(edges, inf.lenA, inf.lenB) == case["edges_full"]
    ctx.close()


def test_archaea_all_vs_all_scores_gpu():
    """BASELINE config 1 on the GPU: one batch of the 21 pairs, scores against the reference fixture."""
    import json
    g = json.load(open(os.path.join(GOLD, "reference_outputs.json")))["archaea"]
    ctx = capi.Context(0, g["flags"])
    ids = [ctx.add_pp(os.path.join(GOLD, "archaea", n + ".pp")) for n in g["names"]]
    for a, b in g["pairs"]:
        ctx.add_pair(ids[a], ids[b])
    ctx.run(capi.RUN_TRACE)
    assert ctx.scores() == g["scores"]
    ctx.close()
