"""Golden outputs of the reference's own `locarna` binary (oracle/_ref/locarna, run in the build container) for bands around a
reference alignment: --max-diff d with --max-diff-pw-aln "rowA&rowB" or --max-diff-aln file (TraceController from a MultipleAlignment,
trace_controller.cc:406-539) -> tests/golden/maxdiffaln_outputs.json + tests/golden/ref_*.aln."""
import json
import os
import random
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O

GOLD = os.path.join(ROOT, "tests", "golden")
PAIRS = (("g0.pp", "g1.pp"), ("g2.pp", "g3.pp"), ("st0.pp", "st1.pp"))
EXTRA = ([], ["--noLP"], ["--sequ-local", "true"], ["--min-trace-probability", "0"], ["--struct-local", "true"])


def regap(s, L, rng):
    pos = set(rng.sample(range(L), L - len(s)))
    it = iter(s)
    return "".join("-" if k in pos else next(it) for k in range(L))


def main():
    rng = random.Random(11)
    out = []
    for a, b in PAIRS:
        pa, pb = os.path.join(GOLD, a), os.path.join(GOLD, b)
        r = O.ref_align(pa, pb, {}, dump="aln")
        names = (a[:-3], b[:-3])
        rows = [(r["rowA"], r["rowB"])]
        sa, sb = r["rowA"].replace("-", ""), r["rowB"].replace("-", "")
        L = max(len(sa), len(sb)) + 3
        rows.append((regap(sa, L, rng), regap(sb, L, rng)))           # some other alignment of the two sequences
        for v, (ra, rb) in enumerate(rows):
            aln = os.path.join(GOLD, "ref_%s_%s_%d.aln" % (names[0], names[1], v))
            with open(aln, "w") as f:
                f.write("CLUSTAL W --- reference alignment for --max-diff-aln\n\n")
                for k in range(0, len(ra), 40):
                    f.write("%-18s %s\n%-18s %s\n\n" % (names[0], ra[k:k + 40], names[1], rb[k:k + 40]))
            combos = [(0, [], "pw"), (2, [], "pw"), (2, [], "file"), (6, ["--noLP"], "file"), (6, ["--sequ-local", "true"], "pw"),
                      (6, ["--min-trace-probability", "0"], "file"), (6, ["--struct-local", "true"], "pw"), (20, [], "file"),
                      (3, ["--max-diff-relax"], "pw"), (8, ["--max-diff-relax", "--noLP"], "file")]
            for delta, extra, mode in combos:
                args = ["--max-diff", str(delta)] + extra + (["--max-diff-pw-aln", ra + "&" + rb] if mode == "pw" else ["--max-diff-aln", os.path.basename(aln)])
                p = subprocess.run([O.REF_LOCARNA, a, b] + args, capture_output=True, text=True, cwd=GOLD)
                out.append({"args": args, "A": a, "B": b, "rc": p.returncode, "stdout": p.stdout, "stderr": p.stderr})
    # argument errors (locarna.cc:504-545)
    for args in (["--max-diff", "5", "--max-diff-pw-aln", "AC-G"], ["--max-diff", "5", "--max-diff-pw-aln", "ACG&AC"],
                 ["--max-diff", "5", "--max-diff-pw-aln", "A&C", "--max-diff-aln", "ref_g0_g1_0.aln"]):
        p = subprocess.run([O.REF_LOCARNA, "g0.pp", "g1.pp"] + args, capture_output=True, text=True, cwd=GOLD)
        out.append({"args": args, "A": "g0.pp", "B": "g1.pp", "rc": p.returncode, "stdout": p.stdout, "stderr": p.stderr})
    with open(os.path.join(GOLD, "maxdiffaln_outputs.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", len(out), "cases;", sum(1 for c in out if c["rc"] != 0), "with non-zero exit")
    for c in out:
        if c["rc"] != 0:
            print(c["args"][:4], c["rc"], c["stderr"].strip()[:100])


if __name__ == "__main__":
    main()
