"""Phase times of one end-to-end job as bench.py's e2e leg runs it (development aid, not a benchmark).
usage: [LB200_TIMING=1] python tools/e2e_phases.py [world]   -- times rank 0's share of a `world`-way split on one GPU"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import bench
from locarna_b200 import capi

world = int(sys.argv[1]) if len(sys.argv) > 1 else 1
args = bench.parse_args(["--job-pairs", "16384"])
pairs = bench.plan(args, world)["pairs"]
paths = bench.make_inputs(512, 300, 16)
pa = np.array([p[0] for p in pairs], dtype=np.int32)
pb = np.array([p[1] for p in pairs], dtype=np.int32)
for it in range(4):
    t0 = time.time(); ctx = capi.Context(0, bench.FLAGS); t1 = time.time()
    first = ctx.add_pps(paths); t2 = time.time()
    mine = ctx.shard_job(pa + first, pb + first, world)[0]; t3 = time.time()
    ctx.add_pairs_np(pa[mine] + first, pb[mine] + first); t4 = time.time()
    ctx.run(); t5 = time.time()
    sc = ctx.scores_np(); t6 = time.time()
    ctx.close(); t7 = time.time()
    print("pairs %d: ctx %.4f parse %.4f shard %.4f addpairs %.4f run %.4f scores %.4f close %.4f total %.4f" % (
        len(mine), t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5, t7 - t6, t7 - t0))
