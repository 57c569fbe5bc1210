"""A synthetic matrix in the extended ribosum format (seeded random frequencies, not derived from any shipped matrix) ->
tests/golden/synthetic.ribosum, and the reference binary's output with --ribosum-file on it -> tests/golden/ribosum_outputs.json."""
import json
import math
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O

GOLD = os.path.join(ROOT, "tests", "golden")
B = "ACGU"


def main():
    rng = np.random.default_rng(2024)
    f_base = rng.dirichlet([8, 7, 9, 7])
    f_non = rng.dirichlet([9, 6, 7, 7])
    noise = rng.normal(0, 0.6, (4, 4)); noise = (noise + noise.T) / 2 + np.eye(4) * 1.4
    f_match = np.outer(f_non, f_non) * np.exp(noise); f_match /= f_match.sum()
    f_pair = np.full((4, 4), 0.004) + rng.random((4, 4)) * 0.002
    for x, y, v in (("A", "U", .14), ("U", "A", .15), ("C", "G", .25), ("G", "C", .27), ("G", "U", .07), ("U", "G", .07)):
        f_pair[B.index(x), B.index(y)] = v
    f_pair /= f_pair.sum()
    fp = f_pair.reshape(16)
    n2 = rng.normal(0, 0.8, (16, 16)); n2 = (n2 + n2.T) / 2 + np.eye(16) * 1.8
    f_am = np.outer(fp, fp) * np.exp(n2); f_am /= f_am.sum()
    bm = np.log2(f_match / np.outer(f_base, f_base))
    am = np.log2(f_am / np.outer(fp, fp))
    arcs = [x + y for x in B for y in B]
    out = ["RIBOSUM_SYNTHETIC", "", "\t" + "\t".join(B)]
    for i in range(4):
        out.append(B[i] + "\t" + "\t".join("%.2f" % bm[i, j] for j in range(i + 1)))
    out += ["", "", "\t" + "\t".join(arcs)]
    for i in range(16):
        out.append(arcs[i] + "\t" + "\t".join("%.2f" % am[i, j] for j in range(i + 1)))
    out += ["", "", "BASE FREQUENCIES", " ".join("%.5e" % x for x in f_base) + " ", "", "BASE NONSTRUCTURAL FREQUENCIES", " ".join("%.5e" % x for x in f_non) + " ", "",
            "BASE PAIR FREQUENCIES"] + [" ".join("%.5e" % x for x in row) + " " for row in f_pair] + ["", "", "BASE MATCH FREQUENCIES",
            " ".join("%.5e" % x for x in f_match.reshape(16)) + " ", "", "BASE PAIR MATCH FREQUENCIES"] + [" ".join("%.5e" % x for x in row) + " " for row in f_am] + [""]
    path = os.path.join(GOLD, "synthetic.ribosum")
    open(path, "w").write("\n".join(out))
    cases = []
    for a, b in (("g0.pp", "g1.pp"), ("g2.pp", "g3.pp"), ("st0.pp", "st1.pp")):
        for args in ([], ["--noLP"], ["--sequ-local", "true", "--tau", "100"], ["--struct-local", "true"], ["--min-trace-probability", "0.01"]):
            args = ["--ribosum-file", "synthetic.ribosum"] + args
            ams = os.path.join(GOLD, "tmp.ams")
            p = subprocess.run([O.REF_LOCARNA, a, b] + args, capture_output=True, text=True, cwd=GOLD)
            w = subprocess.run([O.REF_LOCARNA, a, b, "--write-arcmatch-scores", "tmp.ams"] + args, capture_output=True, text=True, cwd=GOLD)
            cases.append({"args": args, "A": a, "B": b, "rc": p.returncode, "stdout": p.stdout, "stderr": p.stderr, "arcmatch_scores": open(ams).read()})
            os.unlink(ams)
    p = subprocess.run([O.REF_LOCARNA, "g0.pp", "g1.pp", "--ribosum-file", "g0.pp"], capture_output=True, text=True, cwd=GOLD)
    cases.append({"args": ["--ribosum-file", "g0.pp"], "A": "g0.pp", "B": "g1.pp", "rc": p.returncode, "stdout": p.stdout, "stderr": p.stderr, "arcmatch_scores": None})
    json.dump(cases, open(os.path.join(GOLD, "ribosum_outputs.json"), "w"), indent=0)
    for c in cases:
        print(c["args"], c["A"], c["rc"], c["stdout"].split("\n")[0], c["stderr"].strip()[:90])


if __name__ == "__main__":
    main()
