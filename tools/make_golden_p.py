#!/usr/bin/env python
"""Generate tests/golden/locarna_p_outputs.json: outputs of the REFERENCE's AlignerP<double> (oracle/_ref/ref_harness --pf-probs, the
reference's own unmodified sources) for the committed golden inputs: partition function, inside table D (reference arc-match
index order), and the arc-match / base-match probabilities of the outside pass (kept for the outside port of a later round).
Runs in the build container only."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
P = {"pf-double": True, "min-trace-probability": 1e-5}
CASES = [("g0.pp", "g1.pp", P, 1.0), ("g2.pp", "g3.pp", P, 1.0), ("g4.pp", "g5.pp", dict(P, **{"max-diff-am": 8}), 1.0),
         ("g3.pp", "g0.pp", {"pf-double": True, "min-trace-probability": 0, "max-diff": 6}, 4.0),
         ("g1.pp", "g2.pp", dict(P, **{"temperature-alipf": 150}), 1.0)]


def main():
    assert O.have_ref(), "build the reference first: make -C oracle ref"
    out = []
    for a, b, flags, scale in CASES:
        r = O.ref_inside_p(os.path.join(GOLD, a), os.path.join(GOLD, b), flags, scale, probs=True)
        out.append({"A": a, "B": b, "flags": flags, "pf_scale": scale, "Z": r["Z"], "D": r["pfD"], "am": [list(x[:4]) for x in r["am"]],
                    "am_probs": [list(x) for x in r["am_probs"]], "bm_probs": [list(x) for x in r["bm_probs"]]})
        print(a, b, "Z", r["Z"], "arc matches", len(r["pfD"]), "am probs", len(r["am_probs"]), "bm probs", len(r["bm_probs"]))
    json.dump(out, open(os.path.join(GOLD, "locarna_p_outputs.json"), "w"), separators=(",", ":"))


if __name__ == "__main__":
    main()
