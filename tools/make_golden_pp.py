"""Golden --pp outputs (alignment + consensus dot plot in PP 2.0 format) of the reference's own `locarna` binary (oracle/_ref/locarna,
run in the build container) -> tests/golden/pp_outputs.json."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O

GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    out = []
    for a, b in (("g0.pp", "g1.pp"), ("g2.pp", "g3.pp"), ("st0.pp", "st1.pp"), ("ana0.pp", "ana1.pp")):
        for args in ([], ["--sequ-local", "true"], ["--sequ-local", "true", "--local-file-output"], ["--exp-prob", "0.01"], ["--min-prob", "0.01", "--noLP"],
                     ["--stacking", "--exp-prob", "0.001"]):
            if a.startswith("an") and "--sequ-local" in args:
                continue
            pp = os.path.join(GOLD, "tmp_out.pp")
            p = subprocess.run([O.REF_LOCARNA, a, b, "--pp", "tmp_out.pp", "-q"] + args, capture_output=True, text=True, cwd=GOLD)
            out.append({"args": args, "A": a, "B": b, "rc": p.returncode, "pp": open(pp).read()})
            os.unlink(pp)
            print(a, args, p.returncode, len(out[-1]["pp"].splitlines()), "lines", "#STACK" in out[-1]["pp"])
    json.dump(out, open(os.path.join(GOLD, "pp_outputs.json"), "w"), indent=0)


if __name__ == "__main__":
    main()
