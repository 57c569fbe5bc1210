"""Golden stdout and probability files of the reference's own `locarna_p` binary (oracle/_ref/locarna_p, run in the build container)
-> tests/golden/locarna_p_cli.json (inputs: tests/golden/g*.pp). Used by tests/test_gpu_cli.py::test_locarna_p_cli_matches_reference."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O

GOLD = os.path.join(ROOT, "tests", "golden")
REF_LOCARNA_P = os.path.join(os.path.dirname(O.REF_LOCARNA), "locarna_p")
CASES = [
    ([], "g0.pp", "g1.pp"),
    (["--min-am-prob", "0.01", "--min-bm-prob", "0.05"], "g2.pp", "g3.pp"),
    (["--max-diff-am", "8", "--pf-scale", "2"], "g4.pp", "g5.pp"),
    (["--temperature-alipf", "150", "--include-am-in-bm"], "g3.pp", "g0.pp"),
    (["--exp-prob", "0.01", "--maxBPspan", "30"], "g2.pp", "g3.pp"),
    (["-e", "0.002", "--max-bps-length-ratio", "1.0"], "g0.pp", "g1.pp"),
]


def main():
    out = []
    for args, a, b in CASES:
        am, bm = os.path.join(GOLD, "tmp.am"), os.path.join(GOLD, "tmp.bm")
        r = subprocess.run([REF_LOCARNA_P, os.path.join(GOLD, a), os.path.join(GOLD, b), "--write-arcmatch-probs", am, "--write-basematch-probs", bm] + args,
                           capture_output=True, text=True)
        out.append({"args": args, "A": a, "B": b, "rc": r.returncode, "stdout": r.stdout, "am": open(am).read(), "bm": open(bm).read()})
        os.unlink(am); os.unlink(bm)
    with open(os.path.join(GOLD, "locarna_p_cli.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", len(out), "cases")


if __name__ == "__main__":
    main()
