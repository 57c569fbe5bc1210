"""Golden outputs of the compiled reference for PROFILE input (alignments with consensus dot plots, the step after the guide tree in
mlocarna's progressive alignment): the profile inputs themselves are written by the reference binary (`locarna --pp`), then
profile-profile and profile-sequence alignments -> tests/golden/profiles_outputs.json: stdout / clustal / --pp output of
oracle/_ref/locarna, band + arc matches with scores + D + alignment edges of ref_harness (Scoring for alignment columns:
scoring.cc:141-198, :272-311, :369-438)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O
sys.path.insert(0, os.path.join(ROOT, "tests"))
from golden_util import digest

GOLD = os.path.join(ROOT, "tests", "golden")


def ref(args):
    return subprocess.run([O.REF_LOCARNA] + args, capture_output=True, text=True, cwd=GOLD)


def main():
    # profile inputs: two-row alignments of the golden pairs, a three-row one on top of the first
    for out, a, b in (("prof_a.pp", "g0.pp", "g1.pp"), ("prof_b.pp", "g2.pp", "g3.pp"), ("prof_s.pp", "st0.pp", "st1.pp"), ("prof_c.pp", "prof_a.pp", "g4.pp")):
        r = ref([a, b, "--pp", out, "-q"])
        assert r.returncode == 0, r.stderr
    pairs = [("prof_a.pp", "prof_b.pp"), ("prof_a.pp", "g4.pp"), ("g5.pp", "prof_b.pp"), ("prof_c.pp", "prof_b.pp"), ("prof_s.pp", "st2.pp"), ("prof_a.pp", "prof_c.pp")]
    argsets = [([], {}), (["--noLP"], {"noLP": True}), (["--max-diff", "12"], {"max-diff": 12}), (["--min-trace-probability", "0"], {"min-trace-probability": 0}),
               (["--indel-opening", "-300", "--struct-weight", "120"], {"indel-opening": -300, "struct-weight": 120}),
               (["--use-ribosum", "false", "--indel-opening", "0"], {"no-ribosum": True, "indel-opening": 0}),
               (["--tau", "0", "-p", "0.01"], {"tau": 0, "min-prob": 0.01})]
    out = []
    for fa, fb in pairs:
        for args, flags in argsets:
            p = ref([fa, fb, "--clustal", "tmp.aln", "--pp", "tmp.pp"] + args)
            h = O.ref_align(os.path.join(GOLD, fa), os.path.join(GOLD, fb), flags, dump="band,am,D,aln")
            rows = [list(x[:4]) + [sc, d] for x, sc, d in zip(h["am"], h["am_score"], h["D"])]
            case = {"args": args, "flags": flags, "A": fa, "B": fb, "rc": p.returncode, "stdout": p.stdout, "min_col": h["min_col"], "max_col": h["max_col"],
                    "n_am": h["n_am"], "am_sha256": digest(rows), "am_scores_sha256": digest([r[:5] for r in rows]), "am_head": rows[:5], "score": h["score"], "edges_full": [list(e) for e in h["edges_full"]]}
            for key, f in (("clustal", "tmp.aln"), ("pp", "tmp.pp")):
                path = os.path.join(GOLD, f)
                case[key] = open(path).read() if os.path.exists(path) else None
                if os.path.exists(path):
                    os.unlink(path)
            out.append(case)
    with open(os.path.join(GOLD, "profiles_outputs.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", len(out), "cases")
    for c in out:
        print(c["A"], c["B"], c["args"], c["rc"], c["score"], c["n_am"])


if __name__ == "__main__":
    main()
