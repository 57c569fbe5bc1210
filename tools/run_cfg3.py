"""BASELINE config 3: Bralibase-shaped pairwise batch (pairs of related ~100-nt RNAs, global mode, locarna defaults) on one GPU,
scores checked against the compiled reference on a sample. Usage: run_cfg3.py [n_pairs] [trace:0|1]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from locarna_b200 import capi, synth
from oracle import oracle as O

n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
trace = len(sys.argv) > 2 and sys.argv[2] == "1"
t = time.time()
paths = synth.make_family("/tmp/lb200_cfg3_%d" % n_pairs, 3, 2 * n_pairs, lambda rng: int(np.clip(round(rng.normal(100, 15)), 60, 140)), related=True)
print("synth %.1fs" % (time.time() - t), flush=True)
ctx = capi.Context(0, {})
t = time.time(); ids = [ctx.add_pp(p) for p in paths]; print("read %d pp files %.1fs" % (len(paths), time.time() - t), flush=True)
for k in range(n_pairs):
    ctx.add_pair(ids[2 * k], ids[2 * k + 1])
t = time.time(); ctx.upload(); up = time.time() - t
dev, host = ctx.envelope_stats()
print("band derivation + device build %.2fs (envelope: %d on GPU, %d re-checked on host)" % (up, dev, host), flush=True)
for it in range(2):
    t = time.time(); ctx.run(capi.RUN_TRACE if trace else capi.RUN_SCORE_ONLY); w = time.time() - t
    cells = sum(ctx.info(k).cells for k in range(n_pairs))
    print("run wall %.3fs kernel %.1f ms launches %d cells %.3g GCUPS %.1f pairs/s (kernel) %.0f" % (w, ctx.kernel_ms, ctx.launches, cells, cells / ctx.kernel_ms / 1e6, n_pairs / (ctx.kernel_ms / 1e3)), flush=True)
sc = ctx.scores()
sample = list(range(0, n_pairs, max(1, n_pairs // 32)))[:32]
t = time.time()
ref = O.ref_batch([(paths[2 * k], paths[2 * k + 1]) for k in sample], {}, do_trace=trace)
w = time.time() - t
bad = sum(1 for k, r in zip(sample, ref) if r["score"] != sc[k])
print("reference (1 core): %d pairs in %.2fs = %.1f pairs/s; score mismatches %d" % (len(sample), w, len(sample) / w, bad), flush=True)
