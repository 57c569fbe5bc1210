"""Progressive alignment along a guide tree: the loop of mlocarna after the guide-tree stage (src/Utils/mlocarna:3660-3716
perform_progressive_steps, :2683-2715 call_locarna, :2476-2500 result files), over the `locarna`-compatible front end of this package.

The guide tree (newick, e.g. results/result.tree of `mlocarna_tree_b200`) is walked in post order (lib/perl/MLocarna/Tree.pm:410-419);
every inner node aligns the results of its two children - PP files of single sequences at the leaves, profile PP files (alignment +
consensus dot plot, written by `--pp`) further up - with

    locarna_b200 <childA> <childB> <locarna flags> --clustal=<tgt>.aln --pp=<tgt>.pp -q

exactly as mlocarna calls its pairwise aligner; the alignment of the root is copied to results/result.aln and results/result.pp.
Intermediate files are named as mlocarna names them (lib/perl/MLocarna.pm:104-131: "intermediate", "intermediate-1", ...).
Each step is one profile-profile / profile-sequence alignment on the GPU; consensus dot plots are the averaged ones
(`--consensus-structure none`, what mlocarna passes by default), the optional RNAalifold-based variant needs ViennaRNA: out of scope.

    python -m locarna_b200.progressive --treefile T --input-dir DIR --tgtdir OUT [--dry-run] [-- locarna flags]
"""
from __future__ import annotations

import argparse
import os
import re
import shutil
import subprocess
import sys
from dataclasses import dataclass, field

ROOT = os.path.dirname(os.path.abspath(__file__))
LOCARNA = os.path.join(ROOT, "bin", "locarna_b200")


@dataclass
class Node:
    label: str | None = None
    children: list = field(default_factory=list)


def parse_newick(text: str) -> Node:
    """Newick with optional quoted labels and branch lengths (lib/perl/MLocarna/Tree.pm:324-382); lengths are ignored."""
    s = text.strip()
    if s.endswith(";"):
        s = s[:-1]
    pos = 0

    def label() -> str | None:
        nonlocal pos
        if pos < len(s) and s[pos] == "'":
            end = pos + 1
            out = []
            while end < len(s):
                if s[end] == "'":
                    if end + 1 < len(s) and s[end + 1] == "'":
                        out.append("'"); end += 2; continue
                    break
                out.append(s[end]); end += 1
            pos = end + 1
            name = "".join(out)
        else:
            m = re.match(r"[^,():;]*", s[pos:])
            name = m.group(0).strip()
            pos += m.end()
        if pos < len(s) and s[pos] == ":":           # branch length
            m = re.match(r":[^,();]*", s[pos:])
            pos += m.end()
        return name or None

    def node() -> Node:
        nonlocal pos
        n = Node()
        if pos < len(s) and s[pos] == "(":
            pos += 1
            while True:
                n.children.append(node())
                if pos < len(s) and s[pos] == ",":
                    pos += 1
                    continue
                if pos < len(s) and s[pos] == ")":
                    pos += 1
                    break
                raise ValueError("malformed newick at position %d" % pos)
        n.label = label()
        return n

    root = node()
    if pos != len(s):
        raise ValueError("trailing characters in newick at position %d" % pos)
    return root


@dataclass
class Step:
    op1: str      # first operand (PP file of the first child)
    op2: str
    target: str   # <intermediate dir>/<name>, without extension
    size: int     # sequences in the resulting alignment


def _intermediate_name(taken: dict, label: str | None) -> str:
    name = "intermediate" + (re.sub(r"[^A-Za-z0-9]", "_", label) if label is not None else "")
    suf, i = "", 0
    while name + suf in taken:
        i += 1
        suf = "-%d" % i
    taken[name + suf] = None
    return name + suf


def plan(tree: Node, leaf_file, intermediate_dir: str, max_alignment_size: int | None = None):
    """Steps in the order mlocarna performs them, and the base name of the final alignment (None if the size limit stopped it).
    leaf_file(label) -> PP file of a leaf."""
    steps: list = []
    taken: dict = {}

    def visit(n: Node):
        data = [visit(c) for c in n.children]
        if any(d is None for d in data):
            return None
        if len(n.children) >= 2:
            if len(n.children) > 2:
                raise ValueError("the guide tree must be binary")
            size = data[0][1] + data[1][1]
            if max_alignment_size is not None and size > max_alignment_size:
                return None
            tgt = os.path.join(intermediate_dir, _intermediate_name(taken, n.label))
            steps.append(Step(data[0][0], data[1][0], tgt, size))
            return (tgt + ".pp", size)
        if len(n.children) == 0:
            return (leaf_file(n.label), 1)
        return data[0]

    res = visit(tree)
    return steps, (re.sub(r"\.pp$", "", res[0]) if res is not None else None)


def command(step: Step, locarna_args, locarna: str = LOCARNA, stockholm: bool = False, verbose: bool = False):
    cmd = [locarna, step.op1, step.op2] + list(locarna_args) + ["--clustal=%s.aln" % step.target]
    if stockholm:
        cmd.append("--stockholm=%s.stk" % step.target)
    cmd.append("--pp=%s.pp" % step.target)
    if not verbose:
        cmd.append("-q")
    return cmd


def run(tree: Node, leaf_file, tgtdir: str, locarna_args=(), locarna: str = LOCARNA, stockholm: bool = False, verbose: bool = False,
        dry_run: bool = False, max_alignment_size: int | None = None):
    inter, results = os.path.join(tgtdir, "intermediates"), os.path.join(tgtdir, "results")
    steps, final = plan(tree, leaf_file, inter, max_alignment_size)
    cmds = [command(s, locarna_args, locarna, stockholm, verbose) for s in steps]
    if dry_run:
        return cmds, final
    os.makedirs(inter, exist_ok=True)
    os.makedirs(results, exist_ok=True)
    for s, cmd in zip(steps, cmds):
        if verbose:
            print("Align %s + %s --> %s" % (s.op1, s.op2, s.target))
        r = subprocess.run(cmd)
        if r.returncode != 0:
            raise RuntimeError("Command %s failed" % " ".join(cmd))
        if not (os.path.exists(s.target + ".aln") and os.path.exists(s.target + ".pp")):
            raise RuntimeError("Pairwise aligner (%s) failed to write alignment to file" % locarna)
    if final is not None:
        shutil.copy(final + ".aln", os.path.join(results, "result.aln"))
        shutil.copy(final + ".pp", os.path.join(results, "result.pp"))
        if stockholm:
            shutil.copy(final + ".stk", os.path.join(results, "result.stk"))
    return cmds, final


def main(argv=None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    extra = []
    if "--" in argv:
        k = argv.index("--")
        argv, extra = argv[:k], argv[k + 1:]
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--treefile", required=True)
    ap.add_argument("--input-dir", required=True, help="PP files of the sequences: <label> or <label>.pp")
    ap.add_argument("--tgtdir", required=True)
    ap.add_argument("--stockholm", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--dry-run", action="store_true", help="print the commands instead of running them")
    ap.add_argument("--max-alignment-size", type=int, default=None)
    a = ap.parse_args(argv)
    tree = parse_newick(open(a.treefile).read())

    def leaf_file(label):
        for cand in (os.path.join(a.input_dir, label), os.path.join(a.input_dir, label + ".pp")):
            if os.path.exists(cand):
                return cand
        raise FileNotFoundError("no PP file for leaf %r in %s" % (label, a.input_dir))

    cmds, final = run(tree, leaf_file, a.tgtdir, extra, stockholm=a.stockholm, verbose=a.verbose, dry_run=a.dry_run, max_alignment_size=a.max_alignment_size)
    if a.dry_run:
        for c in cmds:
            print(" ".join(c))
    return 0


if __name__ == "__main__":
    sys.exit(main())
