"""BASELINE configs 3 and 4 at their stated sizes against results of the COMPILED REFERENCE (tests/golden/cfg34_reference.json, made
by tools/make_golden_cfg34.py from oracle/_ref/ref_harness), and LocARNA-P at 300 nt against the oracle.

  config 3: 10,000 synthetic pairs of related ~100-nt RNAs, locarna defaults, global: every score (sha256 over all of them, an
            explicit sample of the reference's values, and a live oracle sample)
  config 4: two synthetic 1,500-nt RNAs, --struct-local: score, complete D table, alignment edges, structure strings
"""
import hashlib
import importlib.util
import json
import os

import pytest

from locarna_b200 import capi
from golden_util import full_edges
from oracle import oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "cfg34_reference.json")))


def _gen():
    spec = importlib.util.spec_from_file_location("make_golden_cfg34", os.path.join(ROOT, "tools", "make_golden_cfg34.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def sha(items):
    return hashlib.sha256(",".join(str(x) for x in items).encode()).hexdigest()


def test_config3_ten_thousand_pairs(tmp_path_factory):
    g = GOLD["cfg3"]
    n = g["n_pairs"]
    paths = _gen().cfg3_paths(n, str(tmp_path_factory.mktemp("cfg3")))
    ctx = capi.Context(0, g["flags"])
    first = ctx.add_pps(paths)
    ctx.add_pairs([(first + 2 * k, first + 2 * k + 1) for k in range(n)])
    ctx.run()
    scores = ctx.scores()
    assert ctx.rows_fallbacks == 0
    for k, want in g["sample"].items():                                   # explicit values of the compiled reference
        assert scores[int(k)] == want, k
    assert sha(scores) == g["scores_sha256"]                             # all 10,000 scores
    for k in range(0, n, n // 16):                                        # live oracle sample incl. the reference's cell count
        ref = O.port_align(paths[2 * k], paths[2 * k + 1], g["flags"], do_trace=False)
        assert scores[k] == ref["score"] and ctx.info(k).cells == ref["cells"], k
    ctx.close()


def test_config4_struct_local_1500nt(tmp_path_factory):
    g = GOLD["cfg4"]
    paths = _gen().cfg4_paths(g["n"], str(tmp_path_factory.mktemp("cfg4")))
    ctx = capi.Context(0, g["flags"])
    a, b = ctx.add_pp(paths[0]), ctx.add_pp(paths[1])
    ctx.add_pair(a, b)
    ctx.run(capi.RUN_TRACE | capi.RUN_KEEP_D)
    assert ctx.scores()[0] == g["score"]
    am, sc, D = ctx.arcmatches(0, with_D=True)
    assert len(am) == g["n_arcmatches"]
    assert sha(D) == g["D_sha256"]
    edges, sa, sb = ctx.alignment(0)
    inf = ctx.info(0)
    full = [tuple(e) for e in full_edges(edges, inf.lenA, inf.lenB)]    # incl. locality gaps, as Alignment::alignment_edges(false)
    assert len(full) == g["n_edges"] and sha(full) == g["edges_sha256"]
    # the structure annotation follows from the arc matches on the path: same count of base pairs in both structures as in the edges
    assert sa.count("(") == sa.count(")") == sb.count("(") == sb.count(")") > 0
    ctx.close()


def test_locarna_p_300nt_against_oracle(tmp_path_factory):
    """LocARNA-P complete (inside, outside, probabilities) on one 300-nt pair of the config-5 family: Z, every arc-match probability and
    every base-match probability within 1e-6 relative of the oracle (which is bit-identical with the compiled reference on the
    fixtures of tests/golden/locarna_p_outputs.json)."""
    from locarna_b200 import synth
    paths = synth.make_family(str(tmp_path_factory.mktemp("p300")), 5, 2, 300)
    flags = {"pf-double": True, "min-trace-probability": 1e-5}
    ctx = capi.Context(0, flags)
    a, b = ctx.add_pp(paths[0]), ctx.add_pp(paths[1])
    ctx.add_pair(a, b)
    ctx.run_pf_probs(1.0, 0.001)
    ref = O.port_probs_p(paths[0], paths[1], flags, 1.0, 0.001)
    z = ctx.partition_function(0)
    assert abs(z - ref["Z"]) <= 1e-6 * abs(ref["Z"])
    close = lambda x, y: abs(x - y) <= 1e-6 * max(abs(x), abs(y)) or max(abs(x), abs(y)) < 1e-12
    amp = ctx.arcmatch_probs(0)
    assert len(amp) == len(ref["am_prob"]) and all(close(x, y) for x, y in zip(amp, ref["am_prob"]))
    bm = ctx.basematch_probs(0)
    assert all(close(x, y) for rx, ry in zip(bm, ref["bm"]) for x, y in zip(rx, ry))
    ctx.close()


def test_config4_locarna_p_1500nt(tmp_path_factory):
    """BASELINE config 4, second half: LocARNA-P match probabilities of the two 1,500-nt RNAs against the COMPILED REFERENCE
    (tests/golden/cfg4_locarna_p.json, made by tools/make_golden_p1500.py: AlignerP<double>, about ten minutes on the CPU): the
    partition function, and every arc-match / base-match probability >= 0.001, within 1e-6 relative."""
    path = os.path.join(ROOT, "tests", "golden", "cfg4_locarna_p.json")
    if not os.path.exists(path):
        pytest.skip("fixture not generated")
    g = json.load(open(path))
    from locarna_b200 import synth
    paths = synth.make_family(str(tmp_path_factory.mktemp("cfg4p")), 4, 2, g["n"])
    ctx = capi.Context(0, g["flags"])
    a, b = ctx.add_pp(paths[0]), ctx.add_pp(paths[1])
    ctx.add_pair(a, b)
    ctx.run_pf_probs(1.0, 0.001)
    z = ctx.partition_function(0)
    assert abs(z - g["Z"]) <= 1e-6 * abs(g["Z"])
    am, _ = ctx.arcmatches(0)
    assert len(am) == g["n_arcmatches"]
    amp = ctx.arcmatch_probs(0)
    got = {tuple(am[k]): amp[k] for k in range(len(am))}
    rel = lambda x, y: abs(x - y) / max(abs(x), abs(y), 1e-300)
    assert g["am_probs"] and max(rel(got[tuple(x[:4])], x[4]) for x in g["am_probs"]) <= 1e-6
    # nothing above the reference's output threshold that the reference does not list (its list is >= 0.001)
    listed = {tuple(x[:4]) for x in g["am_probs"]}
    assert all(p < 0.001 * (1 + 1e-6) for k, p in got.items() if k not in listed)
    bm = ctx.basematch_probs(0)
    assert g["bm_probs"] and max(rel(bm[i][j], p) for i, j, p in g["bm_probs"]) <= 1e-6
    ctx.close()
