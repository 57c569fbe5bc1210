"""The `locarna`-compatible command line front end against stdout / --clustal output of the reference's own binary."""
import json
import os
import subprocess

import pytest

from golden_util import GOLD

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "locarna_b200", "bin", "locarna_b200")
CASES = json.load(open(os.path.join(GOLD, "reference_outputs.json")))["cli"]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%s-%s" % ("_".join(c["args"]) or "default", c["A"]))
def test_cli_output_matches_reference(case, tmp_path):
    clu = str(tmp_path / "out.aln")
    r = subprocess.run([CLI, os.path.join(GOLD, case["A"]), os.path.join(GOLD, case["B"]), "--clustal", clu] + case["args"],
                       capture_output=True, text=True)
    assert r.returncode == case["rc"], r.stderr
    assert r.stdout == case["stdout"]
    assert open(clu).read() == case["clustal"]


def test_cli_rejects_unimplemented_modes():
    r = subprocess.run([CLI, os.path.join(GOLD, "g0.pp"), os.path.join(GOLD, "g1.pp"), "--stacking"], capture_output=True, text=True)
    assert r.returncode == 255 and "does not implement" in r.stderr
