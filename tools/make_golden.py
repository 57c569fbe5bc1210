#!/usr/bin/env python
"""Generate tests/golden/*: inputs (PP 2.0) and the outputs of the REFERENCE ITSELF for them.

Runs in the build container only: it executes oracle/_ref/ref_harness, i.e. the reference's own unmodified sources
compiled by oracle/Makefile. The fixtures pin the oracle port (tests/test_oracle_golden.py, CPU) and the CUDA path
(tests/test_gpu_golden.py) wherever /root/reference is absent.

Per case: score, band, arc matches in reference index order with Scoring::arcmatch and the D entries, alignment edges
(full, incl. locality gaps), structure strings and the gapped rows. Large tables are stored as SHA-256 digests of a
canonical text dump plus their first entries.
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from locarna_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

CASES = [
    ("default", {}),
    ("mlocarna_tree", {"noLP": True, "max-diff-am": 30}),
    ("sequ_local", {"sequ-local": True}),
    ("free_endgaps", {"free-endgaps": "++++"}),
    ("max_diff", {"min-trace-probability": 0, "max-diff": 10}),
    ("noribo_linear", {"no-ribosum": True, "indel-opening": 0, "tau": 100}),
]


def digest(obj) -> str:
    return hashlib.sha256(json.dumps(obj, separators=(",", ":")).encode()).hexdigest()


def main():
    os.makedirs(GOLD, exist_ok=True)
    assert O.have_ref(), "build the reference first: make -C oracle ref"
    # inputs: three short related pairs and one length-mismatched pair
    seqs = []
    lens = [36, 44, 52, 60, 72, 30]
    for k, n in enumerate(lens):
        seq = synth.random_sequence(n, 9000 + k) if k % 2 == 0 or k == 5 else synth.mutate(seqs[-1][1], 0.75, 9000 + k)
        name = "g%d" % k
        path = os.path.join(GOLD, name + ".pp")
        synth.make_pp(path, name, seq, seed=9000 + k, density=1.6)
        seqs.append((name, seq, path))
    pairs = [(0, 1), (2, 3), (4, 5), (3, 0)]
    out = {"generator": "tools/make_golden.py via oracle/_ref/ref_harness (LocARNA 2.0.1 reference sources, unmodified)", "cases": []}
    for cname, flags in CASES:
        for a, b in pairs:
            r = O.ref_align(seqs[a][2], seqs[b][2], flags)
            am_rows = [list(x[:4]) + [s, d] for x, s, d in zip(r["am"], r["am_score"], r["D"])]
            case = {
                "case": cname, "flags": flags, "A": os.path.basename(seqs[a][2]), "B": os.path.basename(seqs[b][2]),
                "score": r["score"], "min_col": r["min_col"], "max_col": r["max_col"],
                "n_am": len(am_rows), "am_sha256": digest(am_rows), "am_head": am_rows[:12],
                "edges_full": [list(e) for e in r["edges_full"]],
                "structA": r["structA"], "structB": r["structB"], "rowA": r["rowA"], "rowB": r["rowB"],
            }
            out["cases"].append(case)
    # config 1 of BASELINE.json: the 7 RNAs of Data/Examples/archaea.fa, all-vs-all with the flags of mlocarna's guide-tree
    # stage. No RNAfold in this image: the dot plots are synthetic (synth.dotplot on the real sequences).
    fa = "/root/reference/Data/Examples/archaea.fa"
    if os.path.exists(fa):
        names, cur = [], None
        for line in open(fa):
            line = line.strip()
            if line.startswith(">"):
                cur = [line[1:], ""]; names.append(cur)
            elif line and cur is not None:
                cur[1] += line
        adir = os.path.join(GOLD, "archaea")
        os.makedirs(adir, exist_ok=True)
        paths = []
        for k, (name, seq) in enumerate(names):
            pth = os.path.join(adir, name + ".pp")
            synth.make_pp(pth, name, seq.upper().replace("T", "U"), seed=100 + k, density=1.8)
            paths.append(pth)
        flags = {"struct-weight": 200, "max-diff-am": 30, "noLP": True, "min-prob": 0.001}
        # mlocarna pair order: A = later sequence (mlocarna:3577-3604)
        pl = [(a, b) for a in range(len(paths)) for b in range(a)]
        res = O.ref_batch([(paths[a], paths[b]) for a, b in pl], flags, dump="aln", timing=False)
        out["archaea"] = {"flags": flags, "names": [n for n, _ in names], "pairs": [list(p) for p in pl],
                          "scores": [r["score"] for r in res], "rowA": [r["rowA"] for r in res], "rowB": [r["rowB"] for r in res]}
    # guide trees by the reference's own Perl module (lib/perl/MLocarna/Tree.pm), for the archaea scores and a matrix with ties
    import random, subprocess as sp
    def perl_upgma(names, m):
        script = 'use lib "/root/reference/lib/perl"; use MLocarna::Tree; my @n=(%s); my $m=[%s]; my $t=new MLocarna::Tree("UPGMA",\\@n,$m); print $t->to_newick();' % (
            ",".join('"%s"' % x for x in names), ",".join("[" + ",".join(str(v) for v in row) + "]" for row in m))
        return sp.run(["perl", "-e", script], capture_output=True, text=True, check=True).stdout
    trees = []
    if "archaea" in out:
        from locarna_b200 import allpairs
        g = out["archaea"]
        m = allpairs.assemble_matrix(len(g["names"]), [tuple(p) for p in g["pairs"]], g["scores"])
        trees.append({"names": g["names"], "matrix": m, "newick": perl_upgma(g["names"], m)})
    rnd = random.Random(7)
    for n in (2, 5, 12):
        names = ["seq%d" % k for k in range(n)]
        m = [[0] * n for _ in range(n)]
        for i in range(n):
            for j in range(i):
                m[i][j] = m[j][i] = rnd.choice([-500, -120, 0, 300, 300, 750, 1200])
        trees.append({"names": names, "matrix": m, "newick": perl_upgma(names, m)})
    trees.append({"names": ["a:b", "it's", "plain"], "matrix": [[0, 5, 1], [5, 0, 2], [1, 2, 0]], "newick": perl_upgma(["a:b", "it's", "plain"], [[0, 5, 1], [5, 0, 2], [1, 2, 0]])})
    out["trees"] = trees
    # stdout and --clustal file of the reference's own `locarna` binary (oracle/_ref/locarna) for the CLI parity test
    import subprocess
    cli = []
    for args in ([], ["--noLP", "--max-diff-am", "30"], ["--sequ-local", "true"], ["--free-endgaps", "++++", "--width", "40"],
                 ["--write-structure"], ["-L", "--sequ-local", "true"], ["-q"]):
        for a, b in [(0, 1), (2, 3), (4, 5)]:
            clu = os.path.join(GOLD, "tmp.aln")
            r = subprocess.run([O.REF_LOCARNA, seqs[a][2], seqs[b][2], "--clustal", clu] + args, capture_output=True, text=True)
            cli.append({"args": args, "A": os.path.basename(seqs[a][2]), "B": os.path.basename(seqs[b][2]), "rc": r.returncode,
                        "stdout": r.stdout, "clustal": open(clu).read()})
            os.unlink(clu)
    out["cli"] = cli
    with open(os.path.join(GOLD, "reference_outputs.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
