"""BASELINE config 4 (second half, inside pass): LocARNA-P partition function of two synthetic RNAs on the GPU vs the compiled
reference's AlignerP<double> (ref_harness --pf) / the oracle port. Usage: run_cfg4_p.py [length] [check: ref|port|none]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from locarna_b200 import capi, synth
from oracle import oracle as O

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
check = sys.argv[2] if len(sys.argv) > 2 else "none"
paths = synth.make_family("/tmp/lb200_cfg4p_%d" % n, 4, 2, n)
flags = {"pf-double": True, "min-trace-probability": 1e-5}
ctx = capi.Context(0, flags)
a, b = ctx.add_pp(paths[0]), ctx.add_pp(paths[1])
ctx.add_pair(a, b)
t = time.time(); ctx.upload(); print("band + device build %.2fs" % (time.time() - t), flush=True)
for it in range(2):
    t = time.time(); ctx.run_pf(1.0); w = time.time() - t
    inf = ctx.info(0)
    print("GPU inside: wall %.3fs kernel %.1f ms launches %d K=%d tasks=%d cells=%.3g (%.1f G cell updates/s) Z=%r" % (
        w, ctx.kernel_ms, ctx.launches, inf.n_arcmatches, inf.n_tasks, inf.cells, inf.cells / ctx.kernel_ms / 1e6, ctx.partition_function(0)), flush=True)
if len(sys.argv) > 3 and sys.argv[3] == "probs":
    for it in range(2):
        t = time.time(); ctx.run_pf_probs(1.0, 0.001); w = time.time() - t
        print("GPU inside + outside + probabilities: wall %.3fs kernel %.1f ms launches %d" % (w, ctx.kernel_ms, ctx.launches), flush=True)
    if check == "ref":
        t = time.time(); r = O.ref_inside_p(paths[0], paths[1], flags, probs=True, timing=True); w = time.time() - t
        amp = ctx.arcmatch_probs(0); am, _ = ctx.arcmatches(0)
        got = {tuple(am[k]): amp[k] for k in range(len(am)) if amp[k] >= 0.001}
        want = {tuple(x[:4]): x[4] for x in r["am_probs"]}
        bm = ctx.basematch_probs(0)
        wantb = {(x[0], x[1]): x[2] for x in r["bm_probs"]}
        rel = lambda x, y: abs(x - y) / max(abs(x), abs(y), 1e-300)
        print("reference: %.1fs (inside %s ms, outside+probs %s ms); am probs >= 0.001: %d (GPU %d), max rel dev %.2e; bm probs >= 0.001: %d, max rel dev %.2e" % (
            w, r["time_pf_ms"].get("inside"), r["time_pf_ms"].get("outside"), len(want), len(got), max([rel(got.get(k, 0), want[k]) for k in want] or [0]),
            len(wantb), max([rel(bm[i][j], p) for (i, j), p in wantb.items()] or [0])), flush=True)
    sys.exit(0)
if check != "none":
    t = time.time()
    r = O.ref_inside_p(paths[0], paths[1], flags, timing=True) if check == "ref" else O.port_inside_p(paths[0], paths[1], flags)
    w = time.time() - t
    Zr = r["Z"]; Dr = r["pfD"] if check == "ref" else r["D"]
    D = ctx.arcmatch_pf(0)
    rel = lambda x, y: abs(x - y) / max(abs(x), abs(y), 1e-300)
    print("%s: %.1fs (inside %s ms) Z=%r; rel dev Z %.2e, max rel dev D %.2e over %d arc matches" % (
        check, w, r.get("time_pf_ms", {}).get("inside"), Zr, rel(Zr, ctx.partition_function(0)), max([rel(x, y) for x, y in zip(D, Dr)] or [0]), len(D)), flush=True)
