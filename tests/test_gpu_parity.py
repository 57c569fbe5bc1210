"""GPU parity: D table and score of the CUDA path (through the C ABI) against the oracle port, bit-exact."""
import os

import pytest

from locarna_b200 import capi
from oracle import oracle as O

pytestmark = pytest.mark.gpu

FLAGSETS = [
    {},
    {"noLP": True, "max-diff-am": 30},
    {"sequ-local": True},
    {"free-endgaps": "++++"},
    {"free-endgaps": "+-+-"},
    {"free-endgaps": "-+-+", "noLP": True},
    {"min-trace-probability": 0, "max-diff": 20},
    {"min-trace-probability": 0},
    {"max-diff-at-am": 25, "min-prob": 0.01},
    {"no-ribosum": True, "indel-opening": 0, "tau": 100},
    {"unpaired-penalty": 10, "struct-weight": 150, "noLP": True},
    {"indel-opening": 40, "indel": -120},  # positive opening: explicit border initialisation in the D-fill kernel
    {"struct-local": True},
    {"struct-local": True, "exclusion": -100, "noLP": True},
    {"struct-local": True, "sequ-local": True, "exclusion": -250},
]


def gpu_align(pairs, flags, run_flags=capi.RUN_KEEP_D):
    ctx = capi.Context(device=0, flags=flags)
    ids = {}
    for a, b in pairs:
        for f in (a, b):
            if f not in ids:
                ids[f] = ctx.add_pp(f)
        ctx.add_pair(ids[a], ids[b])
    ctx.run(run_flags)
    return ctx


def check_pairs(pairs, flags, want_kind=None):
    ctx = gpu_align(pairs, flags)
    if want_kind is not None:
        assert ctx.dfill_kind == want_kind and ctx.rows_fallbacks == 0, (ctx.dfill_kind, ctx.rows_fallbacks)
    scores = ctx.scores()
    for k, (a, b) in enumerate(pairs):
        ref = O.port_align(a, b, flags, do_trace=False)
        am, score, D = ctx.arcmatches(k, with_D=True)
        assert am == [x[:4] for x in ref["am"]]
        assert score == ref["am_score"]
        bad = [i for i in range(len(D)) if D[i] != ref["D"][i]]
        assert not bad, "D mismatch at %d of %d arc matches, first %s: gpu %s ref %s" % (
            len(bad), len(D), am[bad[0]], D[bad[0]], ref["D"][bad[0]])
        assert scores[k] == ref["score"], (flags, a, b)
        assert ctx.info(k).cells == ref["cells"]
    ctx.close()


@pytest.mark.parametrize("flags", FLAGSETS)
def test_d_table_and_score(synth_dir, flags):
    pairs = [tuple(synth_dir["cfg2"][:2]), tuple(synth_dir["cfg3"][:2]), tuple(synth_dir["cfg3"][2:4]),
             tuple(synth_dir["short"][:2]), tuple(synth_dir["short"][2:4]), tuple(synth_dir["short"][4:6]),
             (synth_dir["short"][0], synth_dir["cfg3"][5])]
    check_pairs(pairs, flags)


def test_batch_order_independent(synth_dir):
    """Scores do not depend on batch composition / order."""
    fam = synth_dir["cfg3"]
    pairs = [(fam[i], fam[j]) for i in range(4) for j in range(i)]
    flags = {"noLP": True, "max-diff-am": 30}
    c1 = gpu_align(pairs, flags, capi.RUN_SCORE_ONLY)
    c2 = gpu_align(list(reversed(pairs)), flags, capi.RUN_SCORE_ONLY)
    assert c1.scores() == list(reversed(c2.scores()))
    for k, (a, b) in enumerate(pairs):
        assert c1.scores()[k] == O.port_align(a, b, flags, do_trace=False)["score"]


def test_chunked_batches_equal_single_batch(synth_dir, monkeypatch):
    """Large batches are streamed in chunks (LB200_CHUNK_PAIRS): same scores and alignments as one batch."""
    fam = synth_dir["cfg3"]
    pairs = [(fam[i], fam[j]) for i in range(6) for j in range(i)]
    flags = {"noLP": True, "max-diff-am": 30}
    c1 = gpu_align(pairs, flags, capi.RUN_TRACE)
    monkeypatch.setenv("LB200_CHUNK_PAIRS", "4")
    c2 = gpu_align(pairs, flags, capi.RUN_TRACE)
    assert c1.scores() == c2.scores()
    for k in range(len(pairs)):
        assert c1.alignment(k) == c2.alignment(k)
        assert c1.info(k).cells == c2.info(k).cells


NON_SL_FLAGSETS = [f for f in FLAGSETS if not f.get("struct-local")]


@pytest.mark.parametrize("mode,pack", [("dep", "1"), ("dep", "0"), ("levels", "1"), ("levels", "0"), ("auto", "1"), ("auto2", "1")])
@pytest.mark.parametrize("flags", NON_SL_FLAGSETS)
def test_schedules_and_entry_formats(synth_dir, monkeypatch, flags, mode, pack):
    """The box-by-box D fill (LB200_DFILL=dep: dependency-driven persistent launch, levels: one launch per level group), each with the
    packed 8-byte and the 16-byte entry stream, and the row-grouped kernel (auto: wherever it applies; auto2: two columns per lane
    forced): D table, score and cell count bit-exact against the oracle."""
    if mode == "auto2":
        monkeypatch.setenv("LB200_ROWS_FORCE_NC", "2")
    monkeypatch.setenv("LB200_DFILL", "auto" if mode == "auto2" else mode)
    monkeypatch.setenv("LB200_PACK", pack)
    pairs = [tuple(synth_dir["cfg2"][:2]), tuple(synth_dir["cfg3"][:2]), tuple(synth_dir["short"][:2]),
             (synth_dir["short"][0], synth_dir["cfg3"][5])]
    # the row-grouped kernel applies to bands of at most ~60 columns per anti-diagonal and non-positive opening (explicit borders)
    rows_ok = flags.get("indel-opening", -750) <= 0 and (flags.get("min-trace-probability", 1) > 0 or "max-diff" in flags)
    check_pairs(pairs, flags, {"dep": 1, "levels": 0}.get(mode, 2 if rows_ok else None))


@pytest.mark.parametrize("sb", ["1", "3"])
def test_dependency_schedule_pair_blocks_and_traceback(synth_dir, monkeypatch, sb):
    """Dependency-driven schedule with the claim order regrouped by pair blocks (LB200_SB_PAIRS), all-vs-all family batch with
    traceback: same scores and alignments as the level schedule, scores equal to the oracle's."""
    fam = synth_dir["cfg3"]
    pairs = [(fam[i], fam[j]) for i in range(6) for j in range(i)]
    flags = {"noLP": True, "max-diff-am": 30}
    monkeypatch.setenv("LB200_DFILL", "levels")
    c1 = gpu_align(pairs, flags, capi.RUN_TRACE)
    monkeypatch.setenv("LB200_DFILL", "dep")
    monkeypatch.setenv("LB200_SB_PAIRS", sb)
    c2 = gpu_align(pairs, flags, capi.RUN_TRACE)
    assert c1.scores() == c2.scores()
    for k, (a, b) in enumerate(pairs):
        assert c1.alignment(k) == c2.alignment(k)
        if k % 4 == 0:
            assert c2.scores()[k] == O.port_align(a, b, flags, do_trace=False)["score"]
