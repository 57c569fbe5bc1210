"""Extract the body of run_and_report() of the reference's src/locarna.cc - from "bool skip_aligning" down to "arc_matches.reset()" - into
a file that locarna_b200/csrc/cli/refmain_driver.cc includes verbatim (Makefile target `refmain`). Nothing of the reference is
committed: the block is cut from /root/reference at build time. Usage: extract_ref_block.py <locarna.cc> <out.inc>"""
import sys

src, out = sys.argv[1], sys.argv[2]
lines = open(src).read().split("\n")
start = next(k for k, l in enumerate(lines) if "bool skip_aligning = false" in l)
end = next(k for k, l in enumerate(lines) if k > start and "arc_matches.reset()" in l)
block = lines[start:end + 1]
assert any("std::make_unique<Aligner>" in l for l in block) and any("Scoring scoring(" in l for l in block)
open(out, "w").write("// lines %d-%d of %s, verbatim\n" % (start + 1, end + 1, src) + "\n".join(block) + "\n")
print("extracted lines %d-%d (%d lines)" % (start + 1, end + 1, len(block)))
