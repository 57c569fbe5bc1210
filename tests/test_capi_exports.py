"""The C-ABI library loads and exports every symbol that include/locarna_b200.h declares (no compute)."""
import os
import re

from locarna_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "locarna_b200.h")).read()
    declared = set(re.findall(r"\b(lb200_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    lib = capi.load()
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, missing
    assert declared == set(capi.EXPORTS)


def test_params_struct_layout_matches_header():
    """lb200_default_params through ctypes: the Python mirror of lb200_params has the header's field order and sizes (a mismatch
    would shift every later field)."""
    p = capi.make_params({})
    assert (p.min_prob, p.max_diff_am, p.max_diff_at_am, p.max_diff) == (0.001, -1, -1, -1)
    assert (p.struct_weight, p.indel, p.indel_opening, p.tau, p.match, p.mismatch, p.use_ribosum) == (200, -150, -750, 50, 50, 0, 1)
    assert p.free_endgaps == b"----" and p.pf_double == 0
    assert (p.exp_prob, p.max_bps_length_ratio, p.max_bp_span) == (-1.0, 0.0, -1)


def test_front_ends_fail_loudly_without_a_device():
    """No CPU fallback: on a box without a CUDA device the command line front ends exit with an error instead of computing."""
    import subprocess
    import pytest
    try:
        ctx = capi.Context(0, {})
        ctx.close()
        pytest.skip("a CUDA device is present")
    except capi.Error:
        pass
    gold = os.path.join(ROOT, "tests", "golden")
    for exe in ("locarna_b200", "locarna_p_b200", "mlocarna_tree_b200"):
        r = subprocess.run([os.path.join(ROOT, "locarna_b200", "bin", exe), os.path.join(gold, "g0.pp"), os.path.join(gold, "g1.pp")],
                           capture_output=True, text=True)
        assert r.returncode == 255 and "ERROR" in r.stderr, (exe, r.returncode, r.stderr)
