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
