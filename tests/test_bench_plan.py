"""bench.py on CPU: argument parsing at the driver's step counts, the job's pair list, cost-balanced sharding for every N,
the score checksum. (The timed legs need a GPU; the driver's exact command line is exercised by tests/test_gpu_bench.py.)"""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_plan_at_driver_step_counts():
    b = _bench()
    args = b.parse_args(["--gpus", "8", "--steps", "20", "--warmup", "5"])
    pl = b.plan(args, 8)
    assert pl["steps"] == 20 and pl["warmup"] == 5
    pairs = pl["pairs"]
    assert len(pairs) == 16384 and pl["n_full"] == 512 * 511 // 2
    assert len(set(pairs)) == len(pairs)                      # distinct pairs
    assert all(0 <= y < x < 512 for x, y in pairs)            # mlocarna's orientation: A = the later sequence
    assert len({s for p in pairs for s in p}) == 512          # every sequence takes part
    # the job does not depend on N, steps or warmup
    assert b.plan(b.parse_args([]), 1)["pairs"] == pairs


def test_shard_job_partitions_and_balances():
    b = _bench()
    pl = b.plan(b.parse_args(["--seqs", "64", "--job-pairs", "500"]), 1)
    pairs = pl["pairs"]
    n_arcs = [300 + 7 * (k % 13) for k in range(64)]
    lengths = [300] * 64
    from locarna_b200 import allpairs
    costs = [allpairs.pair_cost(n_arcs[a], n_arcs[c], 300, 300) for a, c in pairs]
    for world in (1, 2, 4, 8):
        shards = b.shard_job(pairs, n_arcs, lengths, world)
        assert sorted(k for s in shards for k in s) == list(range(len(pairs)))
        loads = [sum(costs[k] for k in s) for s in shards]
        assert max(loads) - min(loads) <= max(costs)
        for s in shards:
            assert [costs[k] for k in s] == sorted((costs[k] for k in s), reverse=True)


def test_score_checksum_is_order_sensitive_and_handles_neg_inf():
    b = _bench()
    assert b.score_checksum([1, 2, None]) == b.score_checksum([1, 2, None])
    assert b.score_checksum([1, 2, None]) != b.score_checksum([2, 1, None])
    assert len(b.score_checksum([])) == 16
