"""LocARNA-P inside pass on the GPU (FP64) against the oracle: partition function and the inside value of every arc match.

Tolerance: 1e-6 relative (BASELINE.json north_star); the oracle port itself is bit-identical with the compiled reference's
AlignerP<double> (tests/test_oracle_golden.py::test_port_inside_p_matches_reference). The GPU sums the arc-match terms of a cell
with shared-memory atomics and contracts a*b+c into FMAs, so the last bits differ; observed deviations are ~1e-15."""
import glob
import json
import os

import pytest

from locarna_b200 import capi, synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-6
# locarna_p's own defaults for the band: envelope in double, min-trace-probability 1e-5 (locarna_p.cc:171, :285-294)
PFLAGS = {"pf-double": True, "min-trace-probability": 1e-5}
FLAGSETS = [PFLAGS, dict(PFLAGS, **{"max-diff-am": 6}), {"pf-double": True, "min-trace-probability": 0, "max-diff": 8},
            dict(PFLAGS, **{"temperature-alipf": 150, "struct-weight": 150}), dict(PFLAGS, **{"no-ribosum": True, "indel-opening": 0})]


def rel(a, b):
    return abs(a - b) / max(abs(a), abs(b), 1e-300)


def _check(files, pairs, flags, pf_scale=1.0):
    ctx = capi.Context(0, flags)
    ids = [ctx.add_pp(f) for f in files]
    for a, b in pairs:
        ctx.add_pair(ids[a], ids[b])
    ctx.run_pf(pf_scale)
    worst = 0.0
    for k, (a, b) in enumerate(pairs):
        ref = O.port_inside_p(files[a], files[b], flags, pf_scale)
        z = ctx.partition_function(k)
        assert rel(z, ref["Z"]) <= TOL, (flags, a, b, z, ref["Z"])
        D = ctx.arcmatch_pf(k)
        assert len(D) == len(ref["D"])
        for x, y in zip(D, ref["D"]):
            assert rel(x, y) <= TOL or (x == 0 and y == 0), (flags, a, b, x, y)
            worst = max(worst, rel(x, y) if (x or y) else 0.0)
        worst = max(worst, rel(z, ref["Z"]))
    ctx.close()
    return worst


@pytest.mark.parametrize("flags", FLAGSETS)
def test_inside_golden_inputs(flags):
    files = sorted(glob.glob(os.path.join(GOLD, "g*.pp")))
    pairs = [(a, b) for a in range(len(files)) for b in range(a + 1)]
    _check(files, pairs, flags)


def test_inside_pf_scale():
    files = sorted(glob.glob(os.path.join(GOLD, "g*.pp")))
    _check(files, [(0, 1), (2, 3)], PFLAGS, pf_scale=8.0)


def test_inside_against_reference_fixture():
    """Z of the compiled reference's AlignerP<double> (tests/golden/locarna_p_outputs.json, made by tools/make_golden_p.py)."""
    cases = json.load(open(os.path.join(GOLD, "locarna_p_outputs.json")))
    for case in cases:
        ctx = capi.Context(0, case["flags"])
        a, b = ctx.add_pp(os.path.join(GOLD, case["A"])), ctx.add_pp(os.path.join(GOLD, case["B"]))
        ctx.add_pair(a, b)
        ctx.run_pf(case["pf_scale"])
        assert rel(ctx.partition_function(0), case["Z"]) <= TOL, case
        D = ctx.arcmatch_pf(0)
        assert len(D) == len(case["D"])
        assert all(rel(x, y) <= TOL or (x == 0 and y == 0) for x, y in zip(D, case["D"]))
        ctx.close()


def test_inside_longer_pairs(tmp_path):
    paths = synth.make_family(str(tmp_path / "fam"), 77, 4, 120)
    pairs = [(a, b) for a in range(4) for b in range(a)]
    _check(paths, pairs, dict(PFLAGS, **{"max-diff-am": 30}))


def test_rejects_unsupported_modes():
    files = sorted(glob.glob(os.path.join(GOLD, "g*.pp")))
    ctx = capi.Context(0, {"noLP": True})
    a, b = ctx.add_pp(files[0]), ctx.add_pp(files[1])
    ctx.add_pair(a, b)
    with pytest.raises(capi.Error):
        ctx.run_pf(1.0)
    ctx.close()
