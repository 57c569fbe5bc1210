"""Host-side logic of the product (PP reader, base pairs, band incl. 80-bit envelope, arc matches, arc-match
scores) against the oracle port, on a host-only context (no GPU needed)."""
import itertools
import os

import pytest

from locarna_b200 import capi
from oracle import oracle as O

FLAGSETS = [
    {},
    {"noLP": True, "max-diff-am": 30},
    {"sequ-local": True},
    {"free-endgaps": "++++"},
    {"min-trace-probability": 0, "max-diff": 20},
    {"max-diff-at-am": 25, "min-prob": 0.01},
    {"no-ribosum": True, "indel-opening": 0, "tau": 100},
    {"unpaired-penalty": 10, "struct-weight": 150},
]


def host_setup(a, b, flags):
    ctx = capi.Context(device=capi.DEVICE_NONE, flags=flags)
    ia, ib = ctx.add_pp(a), ctx.add_pp(b)
    ctx.add_pair(ia, ib)
    ctx.prepare()
    return ctx


@pytest.mark.parametrize("flags", FLAGSETS)
def test_band_arcmatches_scores(synth_dir, flags):
    fams = [synth_dir["cfg2"][:2], synth_dir["cfg3"][:2], synth_dir["short"][:2], synth_dir["short"][2:4]]
    for a, b in fams:
        ctx = host_setup(a, b, flags)
        ref = O.port_align(a, b, flags, setup_only=True)
        lo, hi = ctx.band(0)
        assert lo == ref["min_col"] and hi == ref["max_col"]
        am, score = ctx.arcmatches(0)
        assert am == [x[:4] for x in ref["am"]]
        assert score == ref["am_score"]
        ctx.close()


def test_run_without_device_fails_loudly(synth_dir):
    a, b = synth_dir["short"][:2]
    ctx = host_setup(a, b, {})
    with pytest.raises(capi.Error):
        ctx.run()


def _cli_flags(args):
    """reference CLI arguments of tests/golden/locarna_cli_options.json -> capi flag dict"""
    names = {"-e": "exp-prob", "--exp-prob": "exp-prob", "--maxBPspan": "maxBPspan", "--max-diff-am": "max-diff-am",
             "--max-bps-length-ratio": "max-bps-length-ratio"}
    flags, k = {}, 0
    while k < len(args):
        if args[k] == "--noLP":
            flags["noLP"] = True; k += 1
        elif args[k] in ("-P", "--pos-output", "-L", "-q", "--local-file-output"):      # output only
            k += 1
        elif args[k] == "--width":
            k += 2
        elif args[k] in ("--sequ-local", "--struct-local"):
            flags[args[k][2:]] = args[k + 1] == "true"; k += 2
        else:
            v = args[k + 1]
            flags[names[args[k]]] = float(v) if "." in v else int(v); k += 2
    return flags


def test_exp_prob_maxbpspan_arcmatch_scores_vs_reference_binary():
    """--exp-prob (arc weights, scoring.cc:201-265 with locarna.cc:662-663) and --maxBPspan (rna_data.cc:1078): arc matches and
    Scoring::arcmatch against the `--write-arcmatch-scores` output of the reference's own binary (fixture)."""
    import json, os
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    for case in json.load(open(os.path.join(gold, "locarna_cli_options.json"))):
        ctx = host_setup(os.path.join(gold, case["A"]), os.path.join(gold, case["B"]), _cli_flags(case["args"]))
        am, score = ctx.arcmatches(0)
        mine = "".join("%d %d %d %d %d\n" % (x[0], x[1], x[2], x[3], s) for x, s in zip(am, score))
        assert mine == case["arcmatch_scores"], case["args"]
        ctx.close()


def test_max_bps_length_ratio_refuses_ties_at_the_cut():
    """drop_worst_bps (rna_data.cc:1580-1601) pops a heap filled in hash order: among equally probable pairs at the cut the
    reference's choice is unspecified, so such an input is refused instead of guessed; a clear cut is accepted."""
    seq = "GGGGAAAACCCC"
    ctx = capi.Context(device=capi.DEVICE_NONE, flags={"max-bps-length-ratio": 0.2})    # keep = int(0.2 * 12) = 2
    ctx.add_seq("ok", seq, [(1, 12, 0.9), (2, 11, 0.8), (3, 10, 0.7), (4, 9, 0.7)])
    with pytest.raises(capi.Error):
        ctx.add_seq("tie", seq, [(1, 12, 0.9), (2, 11, 0.7), (3, 10, 0.7), (4, 9, 0.6)])
    ctx.close()


def test_band_from_reference_alignment_equals_reference():
    """lb200_band_from_alignment (TraceRange from a pairwise reference alignment, trace_controller.cc:44-215) against the compiled
    reference's TraceController (ref_harness --max-diff-pw-aln, envelope off), for the reference's own alignment, a random other
    alignment and other gap symbols."""
    import random
    import pytest
    from locarna_b200 import capi
    if not O.have_ref():
        pytest.skip("compiled reference not available")
    rng = random.Random(5)
    G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    for a, b in (("g0.pp", "g1.pp"), ("g4.pp", "g5.pp")):
        pa, pb = os.path.join(G, a), os.path.join(G, b)
        r = O.ref_align(pa, pb, {}, dump="aln")
        sa, sb = r["rowA"].replace("-", ""), r["rowB"].replace("-", "")
        L = max(len(sa), len(sb)) + 4

        def regap(s):
            pos = set(rng.sample(range(L), L - len(s)))
            it = iter(s)
            return "".join("-" if k in pos else next(it) for k in range(L))
        for A, B in ((r["rowA"], r["rowB"]), (regap(sa), regap(sb)), (r["rowA"].replace("-", "~"), r["rowB"].replace("-", "."))):
            for delta in (0, 1, 4, 15):
                ref = O.ref_align(pa, pb, {"max-diff": delta, "max-diff-pw-aln": A + "&" + B, "min-trace-probability": 0}, dump="band", do_trace=False)
                assert capi.band_from_alignment(len(sa), len(sb), A, B, delta) == (ref["min_col"], ref["max_col"])
                # --max-diff-relax: relaxed merging (consensus trace of the trace ranges, widened by delta; trace_controller.cc:214-311, :485-511)
                ref = O.ref_align(pa, pb, {"max-diff": delta, "max-diff-pw-aln": A + "&" + B, "max-diff-relax": True, "min-trace-probability": 0},
                                  dump="band", do_trace=False)
                assert capi.band_from_alignment(len(sa), len(sb), A, B, delta, relaxed=True) == (ref["min_col"], ref["max_col"])
    with pytest.raises(capi.Error):
        capi.band_from_alignment(5, 4, "ACGU-", "AC-GU", 1)       # rows do not spell out sequences of these lengths


def test_ribosum_file_reader():
    """lb200_set_ribosum_file: the reference's own matrix file gives the built-in tables (arc-match scores unchanged); the synthetic
    matrix of tests/golden changes them; a file in another format is refused."""
    import pytest
    from locarna_b200 import capi
    G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

    def scores(path):
        ctx = capi.Context(capi.DEVICE_NONE, {})
        if path:
            ctx.set_ribosum_file(path)
        ctx.add_pair(ctx.add_pp(os.path.join(G, "g0.pp")), ctx.add_pp(os.path.join(G, "g1.pp")))
        ctx.prepare()
        out = ctx.arcmatches(0)
        ctx.close()
        return out
    builtin = scores(None)
    assert scores("RIBOSUM85_60") == builtin
    ref_file = "/root/reference/Data/Matrices/RIBOSUM85_60"
    if os.path.exists(ref_file):
        assert scores(ref_file) == builtin
    syn = scores(os.path.join(G, "synthetic.ribosum"))
    assert syn != builtin          # other scores (and another band: the envelope uses the matrix' base similarities)
    with pytest.raises(capi.Error, match="Cannot parse ribosum input"):
        scores(os.path.join(G, "g0.pp"))


def test_seqs_copy_shares_parsed_inputs(synth_dir):
    """lb200_seqs_copy: a second context takes over the parsed sequences (bands and arc matches identical to parsing again);
    contexts that filter their inputs differently are refused."""
    import pytest
    from locarna_b200 import capi
    paths = synth_dir["short"][:3]
    flags = {"noLP": True, "max-diff-am": 30}
    a = capi.Context(capi.DEVICE_NONE, flags)
    first = a.add_pps(paths)
    b = capi.Context(capi.DEVICE_NONE, flags)
    assert b.copy_seqs_from(a) == 0 and b.copy_seqs_from(a) == len(paths)        # appended, like add_pps
    for ctx, off in ((a, first), (b, len(paths))):
        ctx.add_pair(off + 1, off + 0)
        ctx.add_pair(off + 2, off + 1)
        ctx.prepare()
    for k in range(2):
        assert a.band(k) == b.band(k) and a.arcmatches(k) == b.arcmatches(k)
    c = capi.Context(capi.DEVICE_NONE, {"min-prob": 0.01})
    with pytest.raises(capi.Error, match="filter their inputs differently"):
        c.copy_seqs_from(a)
    for ctx in (a, b, c):
        ctx.close()
