"""Multi-rank path of the all-vs-all stage on CPU: world_size-2 gloo, shards computed with the oracle port in place of
the GPU, one gather to rank 0, identical score matrix as the single-process run."""
import os
import subprocess
import sys
import textwrap

from locarna_b200 import allpairs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_pairs_partition_and_balance():
    pairs = allpairs.all_vs_all(12)
    costs = [float((a + 1) * (b + 2)) for a, b in pairs]
    for world in (1, 2, 4, 8):
        shards = allpairs.shard_pairs(pairs, costs, world)
        flat = sorted(k for s in shards for k in s)
        assert flat == list(range(len(pairs)))
        loads = [sum(costs[k] for k in s) for s in shards]
        assert max(loads) - min(loads) <= max(costs)
    uneven = allpairs.shard_pairs([(1, 0)] * 7, [100, 1, 1, 1, 1, 1, 1], 2)
    assert uneven == [[0], [1, 2, 3, 4, 5, 6]]
    m = allpairs.assemble_matrix(3, [(1, 0), (2, 0), (2, 1)], [5, None, -7])
    assert m == [[0, 5, -100000000], [5, 0, -7], [-100000000, -7, 0]]
    assert allpairs.format_score_list([(1, 0), (2, 0)], [5, None]) == "1 0 5\n2 0 -inf\n"


WORKER = textwrap.dedent("""
    import os, sys, json
    sys.path.insert(0, %r)
    import torch.distributed as dist
    from locarna_b200 import allpairs
    from oracle import oracle as O
    paths = json.loads(os.environ["LB_PATHS"])
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    pairs = allpairs.all_vs_all(len(paths))
    # very uneven costs: the shares differ in size (one rank holds a single expensive pair), the gather must cope
    costs = [1000.0 if k == 0 else allpairs.pair_cost(10 + a, 10 + b, 40, 40) / 1e6 for k, (a, b) in enumerate(pairs)]
    mine = allpairs.shard_pairs(pairs, costs, world)[rank]
    flags = {"noLP": True, "max-diff-am": 30}
    sc = [O.port_align(paths[pairs[k][0]], paths[pairs[k][1]], flags, do_trace=False)["score"] for k in mine]
    full = allpairs.gather_scores(dist, mine, sc, len(pairs))
    # the vectorised variant bench.py uses: share sizes are known to every rank, so the gather is the only collective
    import numpy as np
    shares = allpairs.shard_pairs(pairs, costs, world)
    full_np = allpairs.gather_scores_np(dist, np.array(mine, dtype=np.int64), np.array(sc, dtype=np.int64), len(pairs), max(len(x) for x in shares))
    if rank == 0:
        assert full_np.tolist() == full
        print("SCORES " + json.dumps(full))
    dist.destroy_process_group()
""")


def test_two_rank_gather_matches_single_process(synth_dir, tmp_path):
    import json
    from oracle import oracle as O
    paths = synth_dir["short"][:5]
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, LB_PATHS=json.dumps(paths), MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29617", str(script)], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("SCORES ")][0]
    got = json.loads(line[len("SCORES "):])
    pairs = allpairs.all_vs_all(len(paths))
    flags = {"noLP": True, "max-diff-am": 30}
    want = [O.port_align(paths[a], paths[b], flags, do_trace=False)["score"] for a, b in pairs]
    assert got == want


def test_shard_job_from_context_equals_cost_then_shard(synth_dir):
    """lb200_shard_job (costs from the context's sequences + LPT in one C call) = pair_cost + shard_pairs."""
    import numpy as np
    from locarna_b200 import capi
    paths = synth_dir["short"][:6]
    ctx = capi.Context(capi.DEVICE_NONE, {"noLP": True})
    first = ctx.add_pps(paths)
    pairs = allpairs.all_vs_all(len(paths))
    a = np.array([p[0] for p in pairs], dtype=np.int32) + first
    b = np.array([p[1] for p in pairs], dtype=np.int32) + first
    costs = [allpairs.pair_cost(ctx.seq_num_arcs(x), ctx.seq_num_arcs(y), ctx.seq_length(x), ctx.seq_length(y)) for x, y in zip(a, b)]
    for world in (1, 2, 3):
        got = [list(map(int, x)) for x in ctx.shard_job(a, b, world)]
        assert got == allpairs.shard_pairs(pairs, costs, world)
    ctx.close()
