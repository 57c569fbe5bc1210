"""Quick GPU sanity + timing run used during development (not a benchmark)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from locarna_b200 import capi, synth
from oracle import oracle as O

n_seq = int(sys.argv[1]) if len(sys.argv) > 1 else 12
length = int(sys.argv[2]) if len(sys.argv) > 2 else 300
flags = {"noLP": True, "max-diff-am": 30}
t = time.time()
paths = synth.make_family("/tmp/quick_%d_%d" % (n_seq, length), 5, n_seq, length)
print("synth %.1fs" % (time.time() - t))
ctx = capi.Context(0, flags)
ids = [ctx.add_pp(p) for p in paths]
pairs = [(a, b) for a in range(n_seq) for b in range(a)]
for a, b in pairs:
    ctx.add_pair(ids[a], ids[b])
t = time.time(); ctx.prepare(); print("prepare %.2fs for %d pairs" % (time.time() - t, len(pairs)))
for it in range(int(os.environ.get("ITERS", "3"))):
    t = time.time(); ctx.run(); w = time.time() - t
    cells = sum(ctx.info(k).cells for k in range(len(pairs)))
    print("run wall %.3fs kernel %.1f ms launches %d cells %.3g GCUPS %.1f pairs/s %.1f" % (w, ctx.kernel_ms, ctx.launches, cells, cells / ctx.kernel_ms / 1e6, len(pairs) / (ctx.kernel_ms / 1e3)))
sc = ctx.scores()
bad = 0
for k in range(min(6, len(pairs))):
    a, b = pairs[k]
    r = O.port_align(paths[a], paths[b], flags, do_trace=False)
    if r["score"] != sc[k]: bad += 1; print("MISMATCH", k, r["score"], sc[k])
print("checked", min(6, len(pairs)), "bad", bad)
