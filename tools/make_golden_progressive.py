"""Golden result of the progressive stage for the archaea example: the driver of locarna_b200/progressive.py run with the REFERENCE's
locarna binary (oracle/_ref/locarna) along the reference guide tree of tests/golden/reference_outputs.json ->
tests/golden/progressive_archaea.json (result.aln, result.pp and every intermediate alignment)."""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from locarna_b200 import progressive as P

GOLD = os.path.join(ROOT, "tests", "golden")
ARGS = ["--max-diff-am", "30"]


def main():
    g = json.load(open(os.path.join(GOLD, "reference_outputs.json")))
    tree = P.parse_newick(g["trees"][0]["newick"] + ";")
    d = tempfile.mkdtemp()
    cmds, final = P.run(tree, lambda l: os.path.join(GOLD, "archaea", l + ".pp"), d, ARGS, locarna=os.path.join(ROOT, "oracle", "_ref", "locarna"))
    out = {"args": ARGS, "newick": g["trees"][0]["newick"], "final": os.path.basename(final), "steps": [],
           "result_aln": open(os.path.join(d, "results", "result.aln")).read(), "result_pp": open(os.path.join(d, "results", "result.pp")).read()}
    for c in cmds:
        tgt = [x for x in c if x.startswith("--clustal=")][0][len("--clustal="):-4]
        out["steps"].append({"name": os.path.basename(tgt), "aln": open(tgt + ".aln").read(), "pp": open(tgt + ".pp").read()})
    json.dump(out, open(os.path.join(GOLD, "progressive_archaea.json"), "w"), indent=0)
    print("wrote", len(out["steps"]), "steps;", out["result_aln"])


if __name__ == "__main__":
    main()
