"""configs 3 and 5 once more with an integer-only band (--min-trace-probability 0 --max-diff 60, no probability
envelope), kernel-only and wall numbers on one GPU, scores checked against the compiled reference on a sample.
Usage: run_int_band.py [cfg3_pairs] [cfg5_seqs]   (development / reporting tool, not the benchmark)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from locarna_b200 import capi, synth
from oracle import oracle as O

BAND = {"max-diff": 60, "min-trace-probability": 0}


def run(label, paths, pairs, flags, check=16):
    ctx = capi.Context(0, flags)
    ids = [ctx.add_pp(p) for p in paths]
    for a, b in pairs:
        ctx.add_pair(ids[a], ids[b])
    t = time.time(); ctx.upload(); up = time.time() - t
    best = None
    for it in range(3):
        t = time.time(); ctx.run(capi.RUN_SCORE_ONLY); w = time.time() - t
        best = ctx.kernel_ms if best is None else min(best, ctx.kernel_ms)
    cells = sum(ctx.info(k).cells for k in range(len(pairs)))
    sc = ctx.scores()
    sample = list(range(0, len(pairs), max(1, len(pairs) // check)))[:check]
    t = time.time()
    ref = O.ref_batch([(paths[pairs[k][0]], paths[pairs[k][1]]) for k in sample], flags, do_trace=False)
    rw = time.time() - t
    bad = sum(1 for k, r in zip(sample, ref) if r["score"] != sc[k])
    print("%s flags %s: %d pairs, band+build %.2fs, kernel %.1f ms (%d launches) = %.0f alignments/s kernel-only, %.3g cells, %.1f GCUPS; "
          "reference 1 core %.1f alignments/s on %d pairs; score mismatches %d"
          % (label, flags, len(pairs), up, best, ctx.launches, len(pairs) / (best / 1e3), cells, cells / best / 1e6, len(sample) / rw, len(sample), bad), flush=True)
    ctx.close()
    return bad


def main():
    n3 = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    n5 = int(sys.argv[2]) if len(sys.argv) > 2 else 46
    bad = 0
    p3 = synth.make_family("/tmp/lb200_cfg3_%d" % n3, 3, 2 * n3, lambda rng: int(np.clip(round(rng.normal(100, 15)), 60, 140)), related=True)
    bad += run("cfg3", p3, [(2 * k, 2 * k + 1) for k in range(n3)], dict(BAND))
    p5 = synth.make_family("/tmp/lb200_cfg5_%d" % n5, 5, n5, 300)
    f5 = {"noLP": True, "max-diff-am": 30, "struct-weight": 200, "min-prob": 0.001}
    f5.update(BAND)
    bad += run("cfg5", p5, [(a, b) for a in range(n5) for b in range(a)], f5)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
