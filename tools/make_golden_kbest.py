"""Golden outputs of the reference's own `locarna` binary (oracle/_ref/locarna, run in the build container) for k-best alignment by
interval splitting (--kbest k / --better t; Aligner::suboptimal, aligner.cc:1383-1514) -> tests/golden/kbest_outputs.json."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O

GOLD = os.path.join(ROOT, "tests", "golden")
ARGS = (["--kbest", "3", "--sequ-local", "true"], ["--kbest", "6", "--sequ-local", "true", "--noLP"], ["--kbest", "2"], ["--kbest", "4"],
        ["--kbest", "4", "--sequ-local", "true", "--normalized", "50"], ["--better", "900", "--sequ-local", "true"],
        ["--kbest", "3", "--sequ-local", "true", "-P"], ["--kbest", "3", "--sequ-local", "true", "--stacking", "--exp-prob", "0.001"],
        ["--kbest", "5", "--sequ-local", "true", "--max-diff", "10"], ["--kbest", "3", "--free-endgaps", "++++"], ["--kbest", "1", "--sequ-local", "true"])
PAIRS = (("g0.pp", "g1.pp"), ("g2.pp", "g3.pp"), ("g4.pp", "g5.pp"), ("st0.pp", "st1.pp"))


def main():
    out = []
    for args in ARGS:
        for a, b in PAIRS:
            r = subprocess.run([O.REF_LOCARNA, os.path.join(GOLD, a), os.path.join(GOLD, b)] + args, capture_output=True, text=True, timeout=600)
            out.append({"args": args, "A": a, "B": b, "rc": r.returncode, "stdout": r.stdout})
            print(args, a, r.returncode, len(r.stdout.splitlines()), "lines")
    with open(os.path.join(GOLD, "kbest_outputs.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", len(out), "cases")


if __name__ == "__main__":
    main()
