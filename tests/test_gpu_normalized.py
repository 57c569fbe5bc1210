"""Normalized and penalized alignment of a BATCH of pairs through the C ABI (lb200_run_normalized / lb200_run_penalized): every pair
runs its own Dinkelbach iteration (aligner.cc:1522-1597). Expected scores: the reference binary's stdout in
tests/golden/normalized_outputs.json (tools/make_golden_normalized.py)."""
import json
import os

import pytest

from golden_util import GOLD
from locarna_b200 import capi

pytestmark = pytest.mark.gpu
CASES = json.load(open(os.path.join(GOLD, "normalized_outputs.json")))
PAIRS = [("g0.pp", "g1.pp"), ("g2.pp", "g3.pp"), ("g4.pp", "g5.pp"), ("st0.pp", "st1.pp")]


def _golden(args):
    out = []
    for a, _ in PAIRS:
        c = [c for c in CASES if c["args"] == args and c["A"] == a][0]
        out.append(int(c["stdout"].split("\n")[0].split()[1]))
    return out


def _ctx(flags):
    ctx = capi.Context(0, flags)
    for a, b in PAIRS:
        ctx.add_pair(ctx.add_pp(os.path.join(GOLD, a)), ctx.add_pp(os.path.join(GOLD, b)))
    return ctx


@pytest.mark.parametrize("L", [0, 50, 200])
def test_normalized_batch(L):
    ctx = _ctx({"sequ-local": True})
    ctx.run_normalized(L)
    assert list(ctx.scores()) == _golden(["--normalized", str(L)])
    for k in range(len(PAIRS)):
        edges, sa, sb = ctx.alignment(k)
        assert len(edges) > 0
    ctx.close()


@pytest.mark.parametrize("pp", [0, 20, -15])
def test_penalized_batch(pp):
    ctx = _ctx({"sequ-local": True})
    ctx.run_penalized(pp)
    assert list(ctx.scores()) == _golden(["--penalized", str(pp)])
    ctx.close()


def test_penalized_zero_is_plain_local_alignment():
    ctx = _ctx({"sequ-local": True})
    ctx.run(capi.RUN_TRACE)
    want, al = list(ctx.scores()), [ctx.alignment(k) for k in range(len(PAIRS))]
    ctx.run_penalized(0)
    assert list(ctx.scores()) == want
    for k in range(len(PAIRS)):
        got = ctx.alignment(k)
        assert got == al[k]
    ctx.close()


def test_normalized_needs_sequ_local_and_rejects_struct_local():
    ctx = _ctx({})
    with pytest.raises(capi.Error):
        ctx.run_normalized(100)
    ctx.close()
    ctx = _ctx({"sequ-local": True, "struct-local": True})
    with pytest.raises(capi.Error):
        ctx.run_normalized(100)
    ctx.close()
