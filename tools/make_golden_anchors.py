"""Golden outputs of the compiled reference for anchor constraints ("#A<k>" annotation in the PP inputs, strict semantics): stdout /
clustal of oracle/_ref/locarna, band + arc matches + D + alignment of ref_harness -> tests/golden/anchors_outputs.json and the
annotated inputs tests/golden/an*.pp (made from g*.pp / st*.pp: anchors on / next to columns of the unconstrained alignment)."""
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


def annotate(src, dst, rows):
    out, done = [], False
    for l in open(src).read().split("\n"):
        out.append(l)
        if not done and l and not l.startswith("#") and len(l.split()) == 2:
            out += ["#A%d %s" % (k + 1, r) for k, r in enumerate(rows)]
            done = True
    open(dst, "w").write("\n".join(out))


def pos(row, k):
    return sum(1 for c in row[:k + 1] if c != "-")


def main():
    inputs = []
    for tag, (a, b), shift, two_rows in (("a", ("g0.pp", "g1.pp"), 0, False), ("b", ("g0.pp", "g1.pp"), 2, False), ("c", ("g2.pp", "g3.pp"), -1, True),
                                         ("d", ("st0.pp", "st1.pp"), 1, False)):
        r = O.ref_align(os.path.join(GOLD, a), os.path.join(GOLD, b), {}, dump="aln")
        cols = [k for k, (x, y) in enumerate(zip(r["rowA"], r["rowB"])) if x != "-" and y != "-"]
        picks = [cols[len(cols) // 5], cols[2 * len(cols) // 5], cols[3 * len(cols) // 5], cols[4 * len(cols) // 5]]
        names = ["A1", "A2", "B1", "C7"] if two_rows else ["A", "B", "C", "D"]
        rows_a = [["."] * r["lenA"] for _ in range(2 if two_rows else 1)]
        rows_b = [["."] * r["lenB"] for _ in range(2 if two_rows else 1)]
        for nm, k in zip(names, picks):
            pa, pb = pos(r["rowA"], k), min(r["lenB"], max(1, pos(r["rowB"], k) + shift))
            for row in range(len(rows_a)):
                rows_a[row][pa - 1] = nm[row]
                rows_b[row][pb - 1] = nm[row]
        fa, fb = "an%s0.pp" % tag, "an%s1.pp" % tag
        annotate(os.path.join(GOLD, a), os.path.join(GOLD, fa), ["".join(x) for x in rows_a])
        annotate(os.path.join(GOLD, b), os.path.join(GOLD, fb), ["".join(x) for x in rows_b])
        inputs.append((fa, fb))
    out = []
    for fa, fb in inputs:
        for args, flags in (([], {}), (["--noLP"], {"noLP": True}), (["--max-diff", "12"], {"max-diff": 12}), (["--min-trace-probability", "0"], {"min-trace-probability": 0}),
                            (["--indel-opening", "-300", "--struct-weight", "120"], {"indel-opening": -300, "struct-weight": 120})):
            clu = os.path.join(GOLD, "tmp.aln")
            p = subprocess.run([O.REF_LOCARNA, fa, fb, "--clustal", "tmp.aln"] + args, capture_output=True, text=True, cwd=GOLD)
            h = O.ref_align(os.path.join(GOLD, fa), os.path.join(GOLD, fb), flags, dump="band,am,D,aln")
            rows = [list(x[:4]) + [sc, d] for x, sc, d in zip(h["am"], h["am_score"], h["D"])]
            out.append({"args": args, "flags": flags, "A": fa, "B": fb, "rc": p.returncode, "stdout": p.stdout, "clustal": open(clu).read() if os.path.exists(clu) else None,
                        "min_col": h["min_col"], "max_col": h["max_col"], "n_am": h["n_am"], "am_sha256": digest(rows), "score": h["score"], "edges_full": [list(e) for e in h["edges_full"]]})
            if os.path.exists(clu):
                os.unlink(clu)
    # a name in one sequence only, relaxed anchors, anchors with local alignment: the B200 path refuses them (documented)
    with open(os.path.join(GOLD, "anchors_outputs.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", len(out), "cases")
    for c in out:
        print(c["A"], c["args"], c["rc"], c["score"], c["n_am"])


if __name__ == "__main__":
    main()
