"""Golden outputs of the reference's own `locarna` binary (oracle/_ref/locarna, run in the build container) for the options
--exp-prob, --maxBPspan, --max-bps-length-ratio, --pos-output and --write-arcmatch-scores -> tests/golden/locarna_cli_options.json (inputs: tests/golden/g*.pp)."""
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
    for args in (["--exp-prob", "0.01"], ["--maxBPspan", "30"], ["-e", "0.0005", "--maxBPspan", "45", "--noLP", "--max-diff-am", "20"], [],
                 ["--max-bps-length-ratio", "1.0"], ["-P"], ["--pos-output", "-L", "--sequ-local", "true"], ["-P", "-q", "--struct-local", "true"],
                 ["--width", "30", "--local-file-output", "--sequ-local", "true"]):
        for a, b in (("g0.pp", "g1.pp"), ("g2.pp", "g3.pp")):
            clu, ams, sto = os.path.join(GOLD, "tmp.aln"), os.path.join(GOLD, "tmp.ams"), os.path.join(GOLD, "tmp.sto")
            pa, pb = os.path.join(GOLD, a), os.path.join(GOLD, b)
            r = subprocess.run([O.REF_LOCARNA, pa, pb, "--clustal", clu, "--stockholm", sto] + args, capture_output=True, text=True)
            w = subprocess.run([O.REF_LOCARNA, pa, pb, "--write-arcmatch-scores", ams] + args, capture_output=True, text=True)
            out.append({"args": args, "A": a, "B": b, "rc": r.returncode, "stdout": r.stdout, "clustal": open(clu).read(), "stockholm": open(sto).read(),
                        "ams_rc": w.returncode, "ams_stdout": w.stdout, "arcmatch_scores": open(ams).read()})
            os.unlink(clu); os.unlink(ams); os.unlink(sto)
    with open(os.path.join(GOLD, "locarna_cli_options.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", len(out), "cases")


if __name__ == "__main__":
    main()
