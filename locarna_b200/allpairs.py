"""All-vs-all stage of mlocarna on top of the C ABI: pair order, sharding over ranks, score matrix / score-list formats.

Reference: src/Utils/mlocarna:3547-3643 (compute_all_pairwise_alignments: pair (a, b) for a in 0..n-1, b in 0..a-1, A = the
later sequence), :2321-2344 (--compute-pairwise-scores k/N partial score lists), lib/perl/MLocarna/SparseMatrix.pm:264-276
(score list lines "<a> <b> <score>"), lib/perl/MLocarna.pm:2034-2052 (result.matrix, "%6d", diagonal 0).
The pairwise alignments are independent, so ranks take disjoint subsets (no data-path collective); the only exchange is the
gather of the score slices to rank 0.
"""
from __future__ import annotations


def all_vs_all(n: int):
    """Pair list in mlocarna's order (lb200_all_vs_all)."""
    return [(a, b) for a in range(n) for b in range(a)]


def pair_cost(n_arcs_a: int, n_arcs_b: int, len_a: int, len_b: int) -> float:
    """Cheap proxy of the D-fill work of a pair: candidate arc matches x band length (lb200_pair_cost)."""
    from . import capi
    return capi.load().lb200_pair_cost(n_arcs_a, n_arcs_b, len_a, len_b)


def shard_pairs(pairs, costs, world: int):
    """Longest-processing-time-first assignment of pairs to ranks (host C++: lb200_shard_pairs, the same split the C++ front
    end uses). Returns a list (per rank) of pair indices, each sorted by descending cost. Deterministic (ties by pair index)."""
    import ctypes as C
    from . import capi
    n = len(pairs)
    cost = (C.c_double * max(n, 1))(*[float(c) for c in costs])
    rank_of = (C.c_int * max(n, 1))()
    order = (C.c_int64 * max(n, 1))()
    begin = (C.c_int64 * (world + 1))()
    rc = capi.load().lb200_shard_pairs(n, cost, world, rank_of, order, begin)
    if rc != capi.OK:
        raise capi.Error("lb200_shard_pairs failed with code %d" % rc)
    return [[order[k] for k in range(begin[r], begin[r + 1])] for r in range(world)]


def gather_scores_np(dist, local_idx, local_scores, n_pairs: int, cap: int, device=None, fill=None):
    """gather_scores for numpy arrays: `cap` = the largest share (every rank knows all share sizes: the split is deterministic),
    so the exchange is the ONE gather and nothing else. Returns an int64 array of n_pairs scores on rank 0 (`fill` where no rank
    reported), None elsewhere."""
    import numpy as np
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    rows = np.full((max(cap, 1), 2), -1, dtype=np.int64)
    rows[:len(local_idx), 0] = local_idx
    rows[:len(local_idx), 1] = local_scores
    buf = torch.from_numpy(rows)
    buf = buf.to(device) if device is not None else buf
    out = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, out, dst=0)
    if rank != 0:
        return None
    allrows = torch.stack(out).reshape(-1, 2).cpu().numpy()
    keep = allrows[:, 0] >= 0
    full = np.full(n_pairs, np.iinfo(np.int64).min if fill is None else fill, dtype=np.int64)
    full[allrows[keep, 0]] = allrows[keep, 1]
    return full


def align_share(paths, pair_a, pair_b, flags, device: int = 0, world: int = 1, rank: int = 0, lanes: int = 2, run_flags: int = 0):
    """This rank's share of an all-vs-all job, from the PP files to the scores: parse, cost-balanced split over `world` ranks
    (lb200_shard_job, the same on every rank), alignment on `device`. The share is worked off by `lanes` contexts on their own
    host threads and CUDA streams: while one context's D fill runs (and tails off), the other one parses, derives bands and builds
    its tables - the same scheme mlocarna_tree_b200 uses across devices. pair_a / pair_b: int32 arrays of sequence indices.
    Returns (pair indices of the share, int64 scores (capi.SCORE_NEG_INF = -inf), largest share of any rank, stats dict)."""
    import threading
    import numpy as np
    from . import capi
    lanes = max(1, lanes)
    out = [None] * lanes
    err = []
    ctxs = [None] * lanes
    parsed = threading.Event()

    def work(lane):
        try:
            ctx = ctxs[lane] = capi.Context(device, flags)
            if lane == 0:
                first = ctx.add_pps(paths)            # every file is parsed once per rank; the other lanes copy the result
                parsed.set()
            else:
                parsed.wait()
                if err:
                    return
                first = ctx.copy_seqs_from(ctxs[0])
            shares = ctx.shard_job(pair_a + first, pair_b + first, world)
            mine = shares[rank][lane::lanes]          # descending cost, dealt out in turn: equal work per lane
            ctx.add_pairs_np(pair_a[mine] + first, pair_b[mine] + first)
            ctx.run(run_flags)
            dev, host = ctx.envelope_stats()
            out[lane] = (mine, ctx.scores_np().copy(), max(len(x) for x in shares),
                         {"h2d": ctx.h2d_bytes, "d2h": ctx.d2h_bytes, "launches": ctx.launches, "env_device": dev, "env_host": host})
        except Exception as e:  # noqa: BLE001 - reported by the caller's thread
            err.append(e)
            parsed.set()

    threads = [threading.Thread(target=work, args=(l,)) for l in range(1, lanes)]
    for t in threads:
        t.start()
    work(0)
    for t in threads:
        t.join()
    for ctx in ctxs:
        if ctx is not None:
            ctx.close()
    if err:
        raise err[0]
    stats = {k: sum(o[3][k] for o in out) for k in out[0][3]}
    return np.concatenate([o[0] for o in out]), np.concatenate([o[1] for o in out]), out[0][2], stats


def assemble_matrix(n: int, pairs, scores, neg_inf_value: int = -100000000):
    """Symmetric score matrix with zero diagonal (mlocarna:2353-2373); '-inf' maps to -1e8 (mlocarna:3523)."""
    m = [[0] * n for _ in range(n)]
    for (a, b), s in zip(pairs, scores):
        v = neg_inf_value if s is None else int(s)
        m[a][b] = v
        m[b][a] = v
    return m


def format_matrix(m) -> str:
    return "".join(" ".join("%6d" % v for v in row) + "\n" for row in m)


def format_score_list(pairs, scores) -> str:
    """scores/scores-<k> format consumed by `mlocarna --score-lists`."""
    return "".join("%d %d %s\n" % (a, b, "-inf" if s is None else str(int(s))) for (a, b), s in zip(pairs, scores))


def gather_scores(dist, local_idx, local_scores, n_pairs: int, device=None):
    """Gather per-rank (pair index, score) slices on rank 0 with one collective. `dist` is torch.distributed (NCCL on the
    GPU box, gloo in the CPU tests). Returns the full score list on rank 0, None elsewhere."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    # shares are balanced by cost, not by count: pad every rank's rows to the largest share (one small all-reduce)
    cnt = torch.tensor([len(local_idx)], dtype=torch.int64)
    cnt = cnt.to(device) if device is not None else cnt
    dist.all_reduce(cnt, op=dist.ReduceOp.MAX)
    cap = max(1, int(cnt.item()))
    NEG = -(2 ** 62)
    rows = [[int(i), NEG if s is None else int(s)] for i, s in zip(local_idx, local_scores)]
    rows += [[-1, -1]] * (cap - len(rows))
    buf = torch.tensor(rows, dtype=torch.int64).to(device) if device is not None else torch.tensor(rows, dtype=torch.int64)   # one H2D copy
    out = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, out, dst=0)
    if rank != 0:
        return None
    scores = [None] * n_pairs
    for t in out:
        for i, s in t.tolist():
            if i >= 0:
                scores[i] = None if s == NEG else s
    return scores
