"""BASELINE config 5, the whole job on one GPU: all-vs-all of 512 synthetic RNAs x 300 nt (130,816 pairs) through one
lb200_run call (streamed in chunks), score matrix in mlocarna's formats, spot check against the oracle."""
import os, sys, time, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from locarna_b200 import capi, allpairs
from oracle import oracle as O

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
cores = os.cpu_count() or 1
t = time.time(); paths = bench.make_inputs(n, 300, cores); seqs = [bench.read_pp(p) for p in paths]; print("inputs %.1fs" % (time.time() - t), flush=True)
pairs = allpairs.all_vs_all(n)
ctx = capi.Context(0, bench.FLAGS)
t0 = time.time()
ids = [ctx.add_seq(*s) for s in seqs]
for a, b in pairs:
    ctx.add_pair(ids[a], ids[b])
t1 = time.time()
ctx.run()
t2 = time.time()
scores = ctx.scores()
dev, host = ctx.envelope_stats()
print("pairs %d: add %.1fs, run %.1fs (kernel %.1fs, %d launches), %.0f alignments/s end to end; envelope: %d on GPU, %d re-checked on host" %
      (len(pairs), t1 - t0, t2 - t1, ctx.kernel_ms / 1e3, ctx.launches, len(pairs) / (t2 - t0), dev, host), flush=True)
m = allpairs.assemble_matrix(n, pairs, scores)
os.makedirs("gpurun_out", exist_ok=True)
open("gpurun_out/cfg5_result.matrix", "w").write(allpairs.format_matrix(m))
open("gpurun_out/cfg5_scores-0", "w").write(allpairs.format_score_list(pairs, scores))
random.seed(1)
sample = random.sample(range(len(pairs)), 24)
bad = 0
for k in sample:
    a, b = pairs[k]
    r = O.port_align(paths[a], paths[b], bench.FLAGS, do_trace=False)
    bad += r["score"] != scores[k]
print("oracle spot check: %d pairs, %d mismatches" % (len(sample), bad))
