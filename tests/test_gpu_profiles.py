"""Profile input (alignments with consensus dot plots: what mlocarna's progressive stage aligns after the guide tree) against the
compiled reference (tests/golden/profiles_outputs.json, tools/make_golden_profiles.py): Scoring for alignment columns (averaged
sigma, gap costs scaled by the gap frequency, averaged ribosum arc-match scores; scoring.cc:141-198, :272-311, :369-438), the envelope
with the column-averaged STRAL score (stral_score.cc:29-44), band, arc matches with scores, the complete D table, score and alignment
edges; stdout / clustal / --pp output of the command line front end (all rows of both inputs, "A." / "B." prefixes on a name clash,
consensus dot plot weighted by the row counts)."""
import json
import os

import pytest

from golden_util import GOLD, digest, full_edges, out_dir, prefetch, run
from locarna_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "locarna_b200", "bin", "locarna_b200")
CASES = json.load(open(os.path.join(GOLD, "profiles_outputs.json")))
IDS = ["%s-%s-%s" % (c["A"], c["B"], "_".join(c["args"]) or "default") for c in CASES]


def test_profile_rows_are_read():
    ctx = capi.Context(capi.DEVICE_NONE, {})
    a, g = ctx.add_pp(os.path.join(GOLD, "prof_c.pp")), ctx.add_pp(os.path.join(GOLD, "g4.pp"))
    rows = ctx.seq_rows(a)
    assert [n for n, _ in rows] == ["g0", "g1", "g4"] and len({len(s) for _, s in rows}) == 1 and any("-" in s for _, s in rows)
    assert len(rows[0][1]) == ctx.seq_length(a)
    assert len(ctx.seq_rows(g)) == 1 and "-" not in ctx.seq_rows(g)[0][1]
    ctx.close()


@pytest.mark.parametrize("flags", [{}, {"noLP": True, "max-diff": 12}, {"no-ribosum": True, "indel-opening": 0}, {"tau": 0, "min-prob": 0.01}])
def test_position_specific_tables_of_single_sequences_equal_the_tables_by_symbol(flags):
    """A context that holds a profile scores ALL its pairs position by position. For two single sequences that must give what the
    tables by symbol code give: same band, same arc matches, same scores."""
    plain = capi.Context(capi.DEVICE_NONE, flags)
    plain.add_pair(plain.add_pp(os.path.join(GOLD, "g2.pp")), plain.add_pp(os.path.join(GOLD, "g3.pp")))
    plain.prepare()
    mixed = capi.Context(capi.DEVICE_NONE, flags)
    mixed.add_pp(os.path.join(GOLD, "prof_a.pp"))
    mixed.add_pair(mixed.add_pp(os.path.join(GOLD, "g2.pp")), mixed.add_pp(os.path.join(GOLD, "g3.pp")))
    mixed.prepare()
    assert plain.band(0) == mixed.band(0)
    assert plain.arcmatches(0) == mixed.arcmatches(0) and len(plain.arcmatches(0)[0]) > 0
    plain.close(); mixed.close()


def test_profile_input_is_refused_where_the_frame_does_not_hold():
    """Local alignment and free end gaps do not commute with the gap-free frame of the profile path: refused when the bands are derived."""
    for flags in ({"sequ-local": True}, {"struct-local": True}, {"free-endgaps": "+---"}):
        ctx = capi.Context(capi.DEVICE_NONE, flags)
        ctx.add_pair(ctx.add_pp(os.path.join(GOLD, "prof_a.pp")), ctx.add_pp(os.path.join(GOLD, "g4.pp")))
        with pytest.raises(capi.Error, match="profile .* input is supported for global alignment"):
            ctx.prepare()
        ctx.close()


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_profile_band_and_arc_match_scores_on_the_host(case):
    """Host mirror (no GPU): band after the envelope over alignment columns and the arc-match list with the averaged scores."""
    ctx = capi.Context(capi.DEVICE_NONE, case["flags"])
    a, b = ctx.add_pp(os.path.join(GOLD, case["A"])), ctx.add_pp(os.path.join(GOLD, case["B"]))
    ctx.add_pair(a, b)
    ctx.prepare()
    lo, hi = ctx.band(0)
    assert lo == case["min_col"] and hi == case["max_col"]
    am, score = ctx.arcmatches(0)
    rows = [list(x) + [s] for x, s in zip(am, score)]
    assert len(rows) == case["n_am"] and rows[:5] == [r[:5] for r in case["am_head"]]
    assert digest(rows) == case["am_scores_sha256"]
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["auto", "levels"])
@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_profiles_match_reference(case, mode, monkeypatch):
    monkeypatch.setenv("LB200_DFILL", mode)
    ctx = capi.Context(0, case["flags"])
    a, b = ctx.add_pp(os.path.join(GOLD, case["A"])), ctx.add_pp(os.path.join(GOLD, case["B"]))
    ctx.add_pair(a, b)
    ctx.run(capi.RUN_TRACE | capi.RUN_KEEP_D)
    lo, hi = ctx.band(0)
    assert lo == case["min_col"] and hi == case["max_col"]
    assert ctx.scores()[0] == case["score"]
    am, score, D = ctx.arcmatches(0, with_D=True)
    rows = [list(x) + [s, d] for x, s, d in zip(am, score, D)]
    assert len(rows) == case["n_am"] and rows[:5] == case["am_head"]
    assert digest(rows) == case["am_sha256"]
    edges, sa, sb = ctx.alignment(0)
    inf = ctx.info(0)
    assert full_edges(edges, inf.lenA, inf.lenB) == case["edges_full"]
    ctx.close()


@pytest.mark.gpu
def test_profile_combinations_that_are_refused():
    a, b = os.path.join(GOLD, "prof_a.pp"), os.path.join(GOLD, "prof_b.pp")
    for flags in ({"sequ-local": True}, {"struct-local": True}, {"free-endgaps": "++++"}):
        ctx = capi.Context(0, flags)
        ctx.add_pair(ctx.add_pp(a), ctx.add_pp(b))
        with pytest.raises(capi.Error, match="profile .* input is supported for global alignment"):
            ctx.run()
        ctx.close()


CLI_CASES = list(enumerate(CASES))


def _cmd(i, case):
    return ([CLI, case["A"], case["B"], "--clustal", os.path.join(out_dir(), "prof%d.aln" % i), "--pp", os.path.join(out_dir(), "prof%d.pp" % i)] + case["args"], GOLD)


@pytest.fixture(scope="module")
def commands_started():
    """All CLI cases are started together, a few processes at a time (golden_util.prefetch)."""
    prefetch([_cmd(i, c) for i, c in CLI_CASES])


@pytest.mark.gpu
@pytest.mark.parametrize("i,case", CLI_CASES, ids=IDS)
def test_cli_with_profiles(i, case, commands_started):
    r = run(*_cmd(i, case))
    assert r.returncode == case["rc"], r.stderr
    assert r.stdout == case["stdout"]
    assert open(os.path.join(out_dir(), "prof%d.aln" % i)).read() == case["clustal"]
    assert open(os.path.join(out_dir(), "prof%d.pp" % i)).read() == case["pp"]
