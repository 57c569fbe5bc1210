"""BASELINE config 4 (first half): one structure-local alignment of two synthetic 1500-nt RNAs, GPU vs the oracle."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from locarna_b200 import capi, synth
from oracle import oracle as O

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
t = time.time(); paths = synth.make_family("/tmp/lb200_cfg4_%d" % n, 4, 2, n); print("synth %.1fs" % (time.time() - t), flush=True)
flags = {"struct-local": True}
ctx = capi.Context(0, flags)
a, b = ctx.add_pp(paths[0]), ctx.add_pp(paths[1])
ctx.add_pair(a, b)
t = time.time(); ctx.prepare(); print("band (host envelope) %.1fs" % (time.time() - t), flush=True)
t = time.time(); ctx.run(capi.RUN_TRACE); w = time.time() - t
inf = ctx.info(0)
print("GPU: wall %.2fs kernel %.1f ms, K=%d tasks=%d cells=%.3g score=%s edges=%d" % (w, ctx.kernel_ms, inf.n_arcmatches, inf.n_tasks, inf.cells, ctx.scores()[0], inf.n_edges), flush=True)
t = time.time(); ctx.run(capi.RUN_TRACE); print("GPU second run: wall %.2fs kernel %.1f ms GCUPS %.1f" % (time.time() - t, ctx.kernel_ms, inf.cells / ctx.kernel_ms / 1e6), flush=True)
if os.environ.get("CHECK", "1") == "1":
    t = time.time(); r = O.port_align(paths[0], paths[1], flags); print("oracle port: %.1fs score %s cells %.3g" % (time.time() - t, r["score"], r["cells"]), flush=True)
    edges, sa, sb = ctx.alignment(0)
    am, sc, D = ctx.arcmatches(0, with_D=True)
    print("score match", ctx.scores()[0] == r["score"], "D match", D == r["D"], "edges match", edges == r["edges"], "struct match", sa == r["strA"] and sb == r["strB"], "cells match", inf.cells == r["cells"])
