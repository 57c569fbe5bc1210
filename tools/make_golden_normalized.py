"""Golden outputs of the reference's own `locarna` binary (oracle/_ref/locarna, run in the build container) for normalized
(--normalized L, aligner.cc:1522-1597) and penalized (--penalized PP, aligner.cc:1599-1622) alignment
-> tests/golden/normalized_outputs.json (inputs: tests/golden/g*.pp, st*.pp)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O

GOLD = os.path.join(ROOT, "tests", "golden")

ARGS = (["--normalized", "0"], ["--normalized", "50"], ["--normalized", "200"], ["--normalized", "1000", "--noLP"],
        ["--normalized", "100", "--sequ-local", "true", "--max-diff", "20"], ["--normalized", "100", "--stacking", "--exp-prob", "0.001"],
        ["--normalized", "30", "--indel-opening", "-300", "--indel", "-250"], ["--normalized", "100", "--sequ-local", "false"],
        ["--normalized", "100", "--struct-local", "true"], ["--normalized", "10", "--penalized", "10"],
        ["--penalized", "0"], ["--penalized", "20"], ["--penalized", "75", "--noLP"], ["--penalized", "-15"],
        ["--penalized", "30", "--sequ-local", "false"], ["--penalized", "25", "--sequ-local", "false", "--free-endgaps", "++--"],
        ["--penalized", "40", "--max-diff-am", "15", "--min-prob", "0.01"])
PAIRS = (("g0.pp", "g1.pp"), ("g2.pp", "g3.pp"), ("g4.pp", "g5.pp"), ("st0.pp", "st1.pp"))


def main():
    out = []
    for args in ARGS:
        for a, b in PAIRS:
            clu = os.path.join(GOLD, "tmp.aln")
            if os.path.exists(clu):
                os.unlink(clu)
            r = subprocess.run([O.REF_LOCARNA, os.path.join(GOLD, a), os.path.join(GOLD, b), "--clustal", clu] + args, capture_output=True, text=True)
            out.append({"args": args, "A": a, "B": b, "rc": r.returncode, "stdout": r.stdout, "stderr": r.stderr,
                        "clustal": open(clu).read() if os.path.exists(clu) else None})
            if os.path.exists(clu):
                os.unlink(clu)
    with open(os.path.join(GOLD, "normalized_outputs.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", len(out), "cases")
    for c in out:
        print(c["args"], c["A"], c["rc"], c["stdout"].split("\n")[0], c["stderr"].strip()[:80])


if __name__ == "__main__":
    main()
