"""Host-side logic of the product (PP reader, base pairs, band incl. 80-bit envelope, arc matches, arc-match
scores) against the oracle port, on a host-only context (no GPU needed)."""
import itertools

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
