"""locarna_b200: B200-native (sm_100a) pairwise sequence-structure alignment, drop-in for LocARNA's Aligner path.

The product is the C-ABI shared library ``liblocarna_b200.so`` (include/locarna_b200.h); ``capi`` binds it
with ctypes for the tests and the benchmark, ``synth`` generates synthetic PP 2.0 inputs.
"""
from . import capi, synth  # noqa: F401
