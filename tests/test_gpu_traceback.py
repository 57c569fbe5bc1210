"""GPU traceback parity: alignment edges and structure strings against the oracle port, bit-exact."""
import pytest

from locarna_b200 import capi
from oracle import oracle as O

pytestmark = pytest.mark.gpu

FLAGSETS = [
    {},
    {"noLP": True, "max-diff-am": 30},
    {"sequ-local": True},
    {"sequ-local": True, "noLP": True},
    {"free-endgaps": "++++"},
    {"free-endgaps": "+-+-"},
    {"free-endgaps": "-+-+", "noLP": True},
    {"min-trace-probability": 0, "max-diff": 20},
    {"no-ribosum": True, "indel-opening": 0, "tau": 100},
    {"indel-opening": 40, "indel": -120},
    {"struct-weight": 400, "tau": 0},
    {"struct-local": True},
    {"struct-local": True, "exclusion": -100, "noLP": True},
    {"struct-local": True, "sequ-local": True, "exclusion": -250},
    {"struct-local": True, "exclusion": -1000, "free-endgaps": "++++"},
]


@pytest.mark.parametrize("flags", FLAGSETS)
def test_alignment_edges(synth_dir, flags):
    pairs = [tuple(synth_dir["cfg2"][:2]), tuple(synth_dir["cfg3"][:2]), tuple(synth_dir["cfg3"][2:4]), tuple(synth_dir["cfg3"][4:6]),
             tuple(synth_dir["short"][:2]), tuple(synth_dir["short"][2:4]), tuple(synth_dir["short"][4:6]),
             (synth_dir["short"][0], synth_dir["cfg3"][5])]
    ctx = capi.Context(0, flags)
    ids = {}
    for a, b in pairs:
        for f in (a, b):
            if f not in ids:
                ids[f] = ctx.add_pp(f)
        ctx.add_pair(ids[a], ids[b])
    ctx.run(capi.RUN_TRACE)
    scores = ctx.scores()
    for k, (a, b) in enumerate(pairs):
        ref = O.port_align(a, b, flags, do_trace=True)
        assert scores[k] == ref["score"]
        edges, sa, sb = ctx.alignment(k)
        assert edges == ref["edges"], (flags, a, b)
        assert sa == ref["strA"] and sb == ref["strB"]
    ctx.close()
