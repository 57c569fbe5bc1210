"""Edge cases on the GPU path against the oracle: no base pairs, very short sequences, no valid arc match, identical inputs,
non-ACGU symbols, ragged lengths."""
import os

import pytest

from locarna_b200 import capi, synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _pp(tmp_path, name, seq, pairs):
    p = str(tmp_path / (name + ".pp"))
    synth.write_pp(p, name, seq, pairs)
    return p


@pytest.fixture()
def files(tmp_path):
    f = {}
    f["nopairs"] = _pp(tmp_path, "nopairs", "ACGUACGUACGUAAGGCC", [])
    f["one"] = _pp(tmp_path, "one", "A", [])
    f["two"] = _pp(tmp_path, "two", "GC", [])
    f["four"] = _pp(tmp_path, "four", "GAAC", [(1, 4, 0.9)])
    f["hairpin"] = _pp(tmp_path, "hairpin", "GGGGAAAACCCC", [(1, 12, 0.9), (2, 11, 0.95), (3, 10, 0.9), (4, 9, 0.5)])
    f["hairpin2"] = _pp(tmp_path, "hairpin2", "GGGCAAAAGCCC", [(1, 12, 0.8), (2, 11, 0.85), (3, 10, 0.7)])
    f["iupac"] = _pp(tmp_path, "iupac", "GGNGAARACYCC", [(1, 12, 0.9), (2, 11, 0.95), (3, 10, 0.9)])
    f["lonely"] = _pp(tmp_path, "lonely", "GAAAAAACAAAAGAAAAC", [(1, 8, 0.6), (13, 18, 0.7)])
    seq = synth.random_sequence(70, 4242)
    f["long"] = _pp(tmp_path, "long", seq, synth.dotplot(seq, seed=4242))
    return f


PAIRS = [("nopairs", "hairpin"), ("hairpin", "nopairs"), ("one", "hairpin"), ("hairpin", "one"), ("two", "four"), ("four", "four"),
         ("hairpin", "hairpin"), ("hairpin", "hairpin2"), ("iupac", "hairpin"), ("hairpin2", "iupac"), ("iupac", "iupac"), ("lonely", "hairpin"),
         ("lonely", "lonely"), ("long", "hairpin"), ("one", "two"), ("long", "nopairs")]
FLAGSETS = [{}, {"noLP": True}, {"sequ-local": True}, {"free-endgaps": "++++"}, {"struct-local": True, "exclusion": -50},
            {"min-trace-probability": 0}, {"max-diff": 3, "min-trace-probability": 0}, {"max-diff-am": 0}, {"min-prob": 0.8}]


@pytest.mark.parametrize("flags", FLAGSETS)
def test_edge_pairs(files, flags):
    ctx = capi.Context(0, flags)
    ids = {k: ctx.add_pp(v) for k, v in files.items()}
    for a, b in PAIRS:
        ctx.add_pair(ids[a], ids[b])
    ctx.run(capi.RUN_TRACE)
    scores = ctx.scores()
    for k, (a, b) in enumerate(PAIRS):
        ref = O.port_align(files[a], files[b], flags)
        assert scores[k] == ref["score"], (flags, a, b, scores[k], ref["score"])
        lo, hi = ctx.band(k)
        assert lo == ref["min_col"] and hi == ref["max_col"], (flags, a, b)
        am, sc, D = ctx.arcmatches(k, with_D=True)
        assert am == [x[:4] for x in ref["am"]] and sc == ref["am_score"] and D == ref["D"], (flags, a, b)
        edges, sa, sb = ctx.alignment(k)
        assert edges == ref["edges"], (flags, a, b, edges, ref["edges"])
        assert sa == ref["strA"] and sb == ref["strB"], (flags, a, b)
    ctx.close()
