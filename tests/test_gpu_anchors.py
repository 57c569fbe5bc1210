"""Anchor constraints ("#A<k>" annotation of the PP inputs, strict semantics; AnchorConstraints + TraceController::restrict_by_anchors +
the allowed_match filter of ArcMatches) against the compiled reference (tests/golden/anchors_outputs.json, tools/make_golden_anchors.py):
band, arc matches with scores and the complete D table, score and alignment through every D-fill kernel; stdout / clustal of the CLI."""
import json
import os
import subprocess

import pytest

from golden_util import GOLD, digest, full_edges, out_dir, prefetch, run
from locarna_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "locarna_b200", "bin", "locarna_b200")
CASES = json.load(open(os.path.join(GOLD, "anchors_outputs.json")))
IDS = ["%s-%s" % (c["A"], "_".join(c["args"]) or "default") for c in CASES]


def _band(case, lo, hi):
    want_hi = list(case["max_col"])
    if case["flags"].get("min-trace-probability", 1) == 0:
        want_hi[0] = min(want_hi[0], want_hi[1])   # row 0 is cut back to row 1 (host_model.cc restrict_band_by_anchors): read by no band cell
    return lo == case["min_col"] and hi == want_hi


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_anchor_band_and_arc_matches_on_the_host(case):
    """Host mirror (no GPU): band after anchors + envelope and the arc-match list."""
    ctx = capi.Context(capi.DEVICE_NONE, case["flags"])
    a, b = ctx.add_pp(os.path.join(GOLD, case["A"])), ctx.add_pp(os.path.join(GOLD, case["B"]))
    assert ctx.seq_anchors(a) != "" and ctx.seq_anchors(b) != ""
    ctx.add_pair(a, b)
    ctx.prepare()
    assert _band(case, *ctx.band(0))
    assert len(ctx.arcmatches(0)[0]) == case["n_am"]
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["auto", "dep", "levels"])
@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_anchors_match_reference(case, mode, monkeypatch):
    monkeypatch.setenv("LB200_DFILL", mode)
    ctx = capi.Context(0, case["flags"])
    a, b = ctx.add_pp(os.path.join(GOLD, case["A"])), ctx.add_pp(os.path.join(GOLD, case["B"]))
    ctx.add_pair(a, b)
    ctx.run(capi.RUN_TRACE | capi.RUN_KEEP_D)
    assert _band(case, *ctx.band(0))
    assert ctx.scores()[0] == case["score"]
    am, score, D = ctx.arcmatches(0, with_D=True)
    rows = [list(x) + [s, d] for x, s, d in zip(am, score, D)]
    assert len(rows) == case["n_am"] and digest(rows) == case["am_sha256"]
    edges, sa, sb = ctx.alignment(0)
    inf = ctx.info(0)
    assert full_edges(edges, inf.lenA, inf.lenB) == case["edges_full"]
    ctx.close()


CLI_CASES = list(enumerate(CASES))[::3]


def _cmd(i, case):
    return ([CLI, case["A"], case["B"], "--clustal", os.path.join(out_dir(), "anchors%d.aln" % i)] + case["args"], GOLD)


@pytest.fixture(scope="module")
def commands_started():
    """All CLI cases are started together, a few processes at a time (golden_util.prefetch)."""
    prefetch([_cmd(i, c) for i, c in CLI_CASES])


@pytest.mark.gpu
@pytest.mark.parametrize("i,case", CLI_CASES, ids=IDS[::3])
def test_cli_with_anchors(i, case, commands_started):
    r = run(*_cmd(i, case))
    assert r.returncode == case["rc"], r.stderr
    assert r.stdout == case["stdout"]
    assert open(os.path.join(out_dir(), "anchors%d.aln" % i)).read() == case["clustal"]


@pytest.mark.gpu
def test_anchor_combinations_that_are_refused():
    a, b = os.path.join(GOLD, "ana0.pp"), os.path.join(GOLD, "ana1.pp")
    for flags in ({"sequ-local": True}, {"struct-local": True}, {"free-endgaps": "++++"}):
        ctx = capi.Context(0, flags)
        ctx.add_pair(ctx.add_pp(a), ctx.add_pp(b))
        with pytest.raises(capi.Error, match="anchor constraints are supported for global alignment"):
            ctx.run()
        ctx.close()
    ctx = capi.Context(0, {})                       # different name sets: refused, never silently ignored
    ctx.add_pair(ctx.add_pp(a), ctx.add_pp(os.path.join(GOLD, "anc1.pp")))
    with pytest.raises(capi.Error, match="only one of the two sequences"):
        ctx.run()
    ctx.close()
    ctx = capi.Context(0, {})                       # one sequence without names: no constraints at all (anchor_constraints.cc:27-31)
    ctx.add_pair(ctx.add_pp(a), ctx.add_pp(os.path.join(GOLD, "g1.pp")))
    ctx.run()
    plain = capi.Context(0, {})
    plain.add_pair(plain.add_pp(os.path.join(GOLD, "g0.pp")), plain.add_pp(os.path.join(GOLD, "g1.pp")))
    plain.run()
    assert ctx.scores() == plain.scores()
    ctx.close(); plain.close()
