"""Golden results for --stacking / --new-stacking from the COMPILED REFERENCE (oracle/_ref/ref_harness: stack weights scoring.cc:201-248,
fill_D_entries[_noLP] aligner.cc:600-607 / :641-655, trace_arcmatch[_noLP] :984-1003 / :1046-1050). Inputs: tests/golden/st*.pp with joint
probabilities (#STACK, fourth column; written by locarna_b200/synth.py). Output: tests/golden/stacking_outputs.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from golden_util import GOLD, digest
from locarna_b200 import synth
from oracle import oracle as O

CASES = [
    {"stacking": True, "exp-prob": 0.01},
    {"stacking": True, "exp-prob": 0.01, "noLP": True},
    {"new-stacking": True, "exp-prob": 0.01},
    {"stacking": True, "new-stacking": True, "exp-prob": 0.005, "noLP": True, "max-diff-am": 20},
    {"stacking": True, "exp-prob": 0.01, "struct-local": True},
    {"stacking": True, "exp-prob": 0.02, "sequ-local": True},
    {"stacking": True, "exp-prob": 0.01, "free-endgaps": "++++", "noLP": True},
]

if __name__ == "__main__":
    lens = [64, 70, 88, 81]
    paths = []
    for k, n in enumerate(lens):
        seq = synth.random_sequence(n, 9000 + k) if k % 2 == 0 else synth.mutate(synth.random_sequence(lens[k - 1], 9000 + k - 1), 0.75, 9000 + k)
        p = os.path.join(GOLD, "st%d.pp" % k)
        synth.make_pp(p, "st%d" % k, seq, seed=9000 + k, stacking=True)
        paths.append(p)
    out = []
    for flags in CASES:
        for a, b in ((0, 1), (2, 3)):
            r = O.ref_align(paths[a], paths[b], flags)
            plain = O.ref_align(paths[a], paths[b], {k: v for k, v in flags.items() if k not in ("stacking", "new-stacking")}, dump="am,D", do_trace=False)
            am_rows = [list(x[:4]) + [s, d] for x, s, d in zip(r["am"], r["am_score"], r["D"])]
            out.append({"flags": flags, "A": os.path.basename(paths[a]), "B": os.path.basename(paths[b]), "score": r["score"], "n_am": len(am_rows),
                        "am_sha256": digest(am_rows), "am_head": am_rows[:8], "edges_full": [list(e) for e in r["edges_full"]],
                        "D_differs_from_unstacked": sum(1 for x, y in zip(r["D"], plain["D"]) if x != y)})
            print(flags, out[-1]["A"], "score", r["score"], "D entries changed by stacking:", out[-1]["D_differs_from_unstacked"])
    json.dump(out, open(os.path.join(GOLD, "stacking_outputs.json"), "w"), indent=0)
