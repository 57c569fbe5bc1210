import sys, time, os
sys.path.insert(0,'/root/repo')
import bench
from locarna_b200 import capi
args=bench.parse_args(["--job-pairs","16384"])
pl=bench.plan(args,1); pairs=pl["pairs"]
paths=bench.make_inputs(512,300,16)
def T(): return time.time()
for it in range(3):
    t0=T(); ctx=capi.Context(0,bench.FLAGS); t1=T()
    first=ctx.add_pps(paths); t2=T()
    n_arcs=[ctx.seq_num_arcs(first+s) for s in range(512)]; lengths=[ctx.seq_length(first+s) for s in range(512)]
    mine=bench.shard_job(pairs,n_arcs,lengths,1)[0]; t3=T()
    ctx.add_pairs([(first+pairs[k][0],first+pairs[k][1]) for k in mine]); t4=T()
    ctx.run(); t5=T()
    sc=ctx.scores(); t6=T()
    ctx.close(); t7=T()
    print("ctx %.3f parse %.3f shard %.3f addpairs %.3f run %.3f scores %.3f close %.3f total %.3f kernel_ms %.1f"%(t1-t0,t2-t1,t3-t2,t4-t3,t5-t4,t6-t5,t7-t6,t7-t0,ctx_k if False else 0))
