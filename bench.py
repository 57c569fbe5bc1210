#!/usr/bin/env python
"""Benchmark of the B200 pairwise sequence-structure alignment path.

Workload (BASELINE.json configs[4], the configuration the metric is quoted on): the all-vs-all guide-tree stage of
mlocarna for 512 synthetic RNAs x 300 nt, flags as mlocarna passes them (--noLP --max-diff-am 30 --struct-weight 200
--min-prob 0.001), score-only mode (the caller consumes only the score, mlocarna:3516-3527).

One "step" is one JOB: a fixed list of --job-pairs (default 16,384) pairs taken evenly from the 130,816-pair all-vs-all
list (every pair list index k * 130816 // 16384; all 512 sequences take part), identical for every N. The job's pairs
are split over the N ranks by estimated arc-match cost (lb200_shard_pairs, longest-processing-time-first), so N GPUs
share a fixed amount of work: STRONG scaling. There is no data-path collective; the ranks' score slices go to rank 0
with one NCCL gather per job, rank 0 assembles the score vector and prints its checksum (identical for every N).

  value  pairs/s over K jobs with the job's tables (arc matches, tasks, bands) already resident in HBM
  e2e    pairs/s for the whole job through the C ABI, inside the timed region of every step: context creation, parse of the
         512 PP 2.0 files (host), cost estimate + sharding, sequences and pair list H2D, band derivation (GPU envelope
         screening + exact host re-check of uncertain pairs), device build, D fill, top level, scores D2H, NCCL gather,
         score vector on rank 0
  --impl reference   the reference's own CPU aligner (oracle/_ref/ref_harness, compiled from the unmodified
         reference sources; the oracle port if that binary is absent), one process per host core

Synthetic data: locarna_b200/synth.py (ViennaRNA is not available in this image), seed = 1000*5 + index.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import hashlib
import json
import os
import statistics
import struct
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLAGS = {"noLP": True, "max-diff-am": 30, "struct-weight": 200, "min-prob": 0.001}
FLAGS_STR = "--noLP --max-diff-am 30 --struct-weight 200 --min-prob 0.001"
METRIC = "all-vs-all pairwise alignments/s (300 nt)"
UNIT = "alignments/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------------------------- data
def _gen_one(args):
    from locarna_b200 import synth
    path, name, n, seed = args
    if not os.path.exists(path):
        tmp = path + ".tmp%d" % os.getpid()
        synth.make_pp(tmp, name, synth.random_sequence(n, seed), seed=seed)
        os.replace(tmp, path)
    return path


def make_inputs(n_seq, length, workers):
    d = "/tmp/lb200_bench_cfg5_%d_%d" % (n_seq, length)
    os.makedirs(d, exist_ok=True)
    jobs = [(os.path.join(d, "s%d.pp" % k), "s%d" % k, length, 1000 * 5 + k) for k in range(n_seq)]
    if all(os.path.exists(j[0]) for j in jobs):
        return [j[0] for j in jobs]
    with cf.ProcessPoolExecutor(max_workers=workers) as ex:
        return list(ex.map(_gen_one, jobs, chunksize=4))


def job_pairs(n_seq, n_job):
    """The job's pair list: n_job pairs taken evenly from mlocarna's all-vs-all order (mlocarna:3577-3604)."""
    from locarna_b200 import allpairs
    full = allpairs.all_vs_all(n_seq)
    n_job = min(n_job, len(full))
    return [full[(k * len(full)) // n_job] for k in range(n_job)], len(full)


def score_checksum(scores):
    """sha256 prefix over the int64 scores in pair order; -inf (None, or the library's LB200_SCORE_NEG_INF = -2^63) as -2^63."""
    h = hashlib.sha256()
    for s in scores:
        h.update(struct.pack("<q", -(2 ** 63) if s is None else int(s)))
    return h.hexdigest()[:16]


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for line in self.lines:
            t = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(t[0])); mx = float(t[1])
            except (ValueError, IndexError):
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], t[3:7]):
                if v == "Active":
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- CPU reference arm
def _ref_worker(args):
    harness, flags, pairs = args
    from oracle import oracle as O
    t = time.time()
    if harness:
        res = O.ref_batch(pairs, flags, dump="", do_trace=True, timing=True)
        scores = [r["score"] for r in res]
    else:
        scores = [O.port_align(a, b, flags, do_trace=True)["score"] for a, b in pairs]
    return scores, time.time() - t


class CpuReference:
    """The reference's CPU aligner on the host cores: one worker process per core, kept alive across steps (mlocarna keeps its
    worker threads too, mlocarna:3607-3643); every worker runs the reference path per pair: PP parse, envelope, arc matches,
    DP, traceback."""

    def __init__(self, cores):
        from oracle import oracle as O
        self.cores = cores
        self.kind = "reference" if O.have_ref() else "port"
        if self.kind == "port":
            O.lib()
        self.pool = cf.ProcessPoolExecutor(max_workers=cores)
        list(self.pool.map(abs, range(cores * 4)))   # start the workers before anything is timed

    def align(self, paths, pairs):
        """-> (scores by pair, wall seconds)"""
        file_pairs = [(paths[a], paths[b]) for a, b in pairs]
        chunks = [c for c in (file_pairs[k::self.cores] for k in range(self.cores)) if c]
        t = time.time()
        out = list(self.pool.map(_ref_worker, [(self.kind == "reference", FLAGS, c) for c in chunks]))
        wall = time.time() - t
        scores = {}
        for k, (sc, _) in enumerate(out):
            for idx, s in enumerate(sc):
                scores[pairs[k + idx * self.cores]] = s
        return scores, wall

    def close(self):
        self.pool.shutdown()


def plan(args, world):
    """Argument / sharding logic shared by both arms (covered by tests/test_bench_plan.py on CPU)."""
    pairs, n_full = job_pairs(args.seqs, args.job_pairs)
    return {"pairs": pairs, "n_full": n_full, "steps": args.steps, "warmup": max(args.warmup, 0), "world": world,
            "sub_batch": max(1, args.sub_batch)}


def shard_job(pairs, n_arcs, lengths, world):
    """Cost-balanced split of the job's pair list (indices into `pairs`, per rank, descending cost)."""
    from locarna_b200 import allpairs
    costs = [allpairs.pair_cost(n_arcs[a], n_arcs[b], lengths[a], lengths[b]) for a, b in pairs]
    return allpairs.shard_pairs(pairs, costs, world)


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--job-pairs", type=int, default=16384, help="pairs of one job (= one step), split over the GPUs")
    ap.add_argument("--sub-batch", type=int, default=2048, help="pairs per resident device batch")
    ap.add_argument("--seqs", type=int, default=512)
    ap.add_argument("--len", type=int, default=300)
    ap.add_argument("--cpu-sample", type=int, default=0, help="pairs of the CPU baseline sample (default: 2 per core)")
    ap.add_argument("--lanes", type=int, default=2, help="contexts (host thread + CUDA stream each) that work off a rank's share in the e2e leg")
    ap.add_argument("--traced-steps", type=int, default=1, help="extra e2e jobs with traceback (reported as e2e_traced)")
    return ap.parse_args(argv)


# ----------------------------------------------------------------------------------------------- main
def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    pl = plan(args, world)
    K, W, pairs = pl["steps"], pl["warmup"], pl["pairs"]
    workload = "cfg5 mlocarna guide-tree stage, all-vs-all %d x %d nt: one step = one job of %d pairs taken evenly from the %d-pair list" % (
        args.seqs, args.len, len(pairs), pl["n_full"])

    if args.impl == "reference":
        if rank != 0:
            return 0
        paths = make_inputs(args.seqs, args.len, cores)
        per_step = args.cpu_sample or 2 * cores
        ref = CpuReference(cores)
        times = []
        for s in range(W + K):
            lo = (s * per_step) % max(1, len(pairs) - per_step)
            _, wall = ref.align(paths, pairs[lo:lo + per_step])
            if s >= W:
                times.append(wall)
        ref.close()
        total = sum(times)
        value = K * per_step / total
        line = {
            "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * total / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int64",
            "data": "synthetic",
            "config": {"workload": workload + "; reference arm: each step a sample of %d of these pairs on %d host cores" % (per_step, cores),
                       "flags": FLAGS_STR, "pairs_per_step": per_step},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": ref.kind,
                             "sample": "%d pairs per step, one persistent worker process per core, reference path per pair incl. PP parse, envelope, arc matches, DP, traceback" % per_step},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        print(json.dumps(line), flush=True)
        return 0

    # ------------------------------------------------------------------ B200 arm
    import torch
    if not torch.cuda.is_available():
        log("bench.py: no CUDA device - the B200 arm has no CPU fallback")
        return 2
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from locarna_b200 import allpairs, capi
    if world > 1:   # the ranks share the host cores: each library instance uses its part of them for parsing and table preparation
        os.environ.setdefault("LB200_HOST_THREADS", str(max(2, cores // world)))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    if rank == 0:
        paths = make_inputs(args.seqs, args.len, cores)
    if dist is not None:
        dist.barrier()
    if rank != 0:
        paths = make_inputs(args.seqs, args.len, 1)
    n_steps = W + K

    import numpy as np
    pair_a = np.array([p[0] for p in pairs], dtype=np.int32)
    pair_b = np.array([p[1] for p in pairs], dtype=np.int32)

    def load_job(ctx):
        """Sequences from the PP files, cost estimate and this rank's share of the job's pair list (lb200_shard_job: every rank
        computes the same deterministic split). Returns the first sequence id, this rank's pair indices, the largest share."""
        first = ctx.add_pps(paths)
        shares = ctx.shard_job(pair_a + first, pair_b + first, world)
        return first, shares[rank], max(len(x) for x in shares)

    # ---- resident leg: build + upload this rank's share before the timed region, in device batches of --sub-batch pairs
    t0 = time.time()
    ctxs, resident_bytes = [], 0
    probe = capi.Context(local_rank, FLAGS)
    _, mine, _ = load_job(probe)
    probe.close()
    mine = [int(k) for k in mine]
    for lo in range(0, len(mine), pl["sub_batch"]):
        part = mine[lo:lo + pl["sub_batch"]]
        ctx = capi.Context(local_rank, FLAGS)
        first = ctx.add_pps(paths)
        ctx.add_pairs([(first + pairs[k][0], first + pairs[k][1]) for k in part])
        ctx.upload()
        # S-order entry 16 B (+ 8 B packed copy) + L-order record 20 B per arc match, 20 B per D-fill task (dev_types.h)
        resident_bytes += sum(44 * ctx.info(k).n_arcmatches + 20 * ctx.info(k).n_tasks for k in range(len(part)))
        ctxs.append((ctx, part))
    prep_s = time.time() - t0

    def resident_step():
        for ctx, _ in ctxs:
            ctx.run()

    for s in range(W):
        resident_step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t0 = time.time()
    kernel_ms = dfill_ms = 0.0
    dfill_launches = launches = 0
    dfill_kinds, rows_fallbacks = set(), 0
    for s in range(K):
        resident_step()
        for ctx, _ in ctxs:
            kernel_ms += ctx.kernel_ms; dfill_ms += ctx.dfill_ms; dfill_launches += ctx.dfill_launches; launches += ctx.launches
            dfill_kinds.add(ctx.dfill_kind)
    barrier()
    elapsed = time.time() - t0
    clocks = sampler.stop()
    rows_fallbacks = sum(ctx.rows_fallbacks for ctx, _ in ctxs)
    cells = terms = am = arcs = rows = 0
    resident_scores = {}
    for ctx, part in ctxs:
        sc = ctx.scores()
        for k, idx in enumerate(part):
            inf = ctx.info(k)
            cells += inf.cells; terms += inf.terms; am += inf.n_arcmatches; arcs += inf.n_arcsA + inf.n_arcsB; rows += inf.lenA + 1
            resident_scores[idx] = sc[k]
        ctx.close()
    ctxs = []

    # ---- e2e leg: the whole job from the PP files on disk to the score vector on rank 0, everything timed
    env_stats = [0, 0]

    def e2e_step(run_flags=capi.RUN_SCORE_ONLY):
        mine, sc, cap, st = allpairs.align_share(paths, pair_a, pair_b, FLAGS, device=local_rank, world=world, rank=rank,
                                                 lanes=args.lanes, run_flags=run_flags)
        env_stats[0] += st["env_device"]; env_stats[1] += st["env_host"]
        if dist is not None:  # the one collective of the path: score slices to rank 0 (NCCL gather)
            full = allpairs.gather_scores_np(dist, mine, sc, len(pairs), cap, device="cuda", fill=capi.SCORE_NEG_INF)
        else:
            full = np.full(len(pairs), capi.SCORE_NEG_INF, dtype=np.int64)
            full[mine] = sc
        return full, st["h2d"], st["d2h"], st["launches"]

    for s in range(W):
        e2e_step()
    barrier()
    t0 = time.time()
    h2d = d2h = e2e_launches = 0
    full = None
    for s in range(K):
        full, a, b, nl = e2e_step()
        h2d += a; d2h += b; e2e_launches += nl
    barrier()
    e2e_elapsed = time.time() - t0
    if full is not None:   # rank 0: plain Python values for the checks below (outside the timed region)
        full = [None if v == capi.SCORE_NEG_INF else int(v) for v in full]
    traced_elapsed = None
    if args.traced_steps > 0:
        barrier()
        t0 = time.time()
        for s in range(args.traced_steps):
            e2e_step(capi.RUN_TRACE)
        barrier()
        traced_elapsed = time.time() - t0

    def reduce_ranks(x, op):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    MAX = dist.ReduceOp.MAX if dist is not None else None
    SUM = dist.ReduceOp.SUM if dist is not None else None
    elapsed = reduce_ranks(elapsed, MAX)
    e2e_elapsed = reduce_ranks(e2e_elapsed, MAX)
    if traced_elapsed is not None:
        traced_elapsed = reduce_ranks(traced_elapsed, MAX)
    my_dfill_ms, my_cells = dfill_ms, cells
    dfill_ms_max = reduce_ranks(dfill_ms, MAX)
    kernel_ms_max = reduce_ranks(kernel_ms, MAX)
    cells_total = reduce_ranks(cells, SUM)
    launches_total = int(reduce_ranks(launches, SUM))
    h2d_total = int(reduce_ranks(h2d, SUM)); d2h_total = int(reduce_ranks(d2h, SUM))
    resident_total = reduce_ranks(resident_bytes, SUM)
    # the resident leg's scores must equal the e2e leg's (rank 0 holds the e2e score vector)
    mismatch = 0
    if rank == 0:
        mismatch = sum(1 for idx, s in resident_scores.items() if full[idx] != s)
    total_pairs = K * len(pairs)
    value = total_pairs / elapsed
    e2e_value = total_pairs / e2e_elapsed

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        peak_src = "measured" if "hbm_gbs" in peaks else "fallback"
        # dominant kernel = the D fill; figures of rank 0 per launch (DESIGN.md "roofline"). SURVEY 8d R1 counts 9 int ops per cell
        # update + 2 per arc-match term; the kernel does not tally the entries it folds, so the term part is left out (conservative).
        dfill_s = my_dfill_ms / 1e3
        ops = 9 * my_cells * K
        sm_mhz = clocks.get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
        alu_peak = 148 * 128 * sm_mhz * 1e6
        alg_bytes = (8 * arcs + 8 * rows + 12 * am + 16 * len(resident_scores)) * K
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("dfill_dram_bytes_per_launch")
        except (OSError, ValueError):
            pass
        n_dfill = max(1, dfill_launches)
        roofline = {"bound": "int_alu", "achieved": ops / dfill_s / 1e12, "peak": alu_peak / 1e12, "unit": "Tintop/s",
                    "frac": ops / dfill_s / alu_peak, "traffic": traffic,
                    "kernel": " + ".join({0: "dfill_kernel (one launch per level group)", 1: "dfill_dep_kernel (one box per warp, dependency counters)",
                                          2: "dfill_rows_kernel (row-grouped sweep, one persistent launch per sub-batch)"}.get(k, "?") for k in sorted(dfill_kinds)),
                    "rows_fallbacks": rows_fallbacks,
                    "traffic_source": "profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of one launch on a 2048-pair sub-batch (ncu --set full)",
                    "launch_ms": my_dfill_ms / n_dfill, "launches": dfill_launches,
                    "ops": "9 int ops per cell update (SURVEY 8d R1; the 2 ops per folded arc-match term are not counted)",
                    "gcups": my_cells * K / dfill_s / 1e9, "sm_mhz": sm_mhz,
                    "peak_source": "148 SMs x 128 INT32 lanes x SM clock sampled under load",
                    "note": "max-plus DP with data-dependent gathers: not tensor-core work; the binding roof is integer-ALU issue, the HBM figure is in roofline_hbm"}
        roofline_hbm = {"bound": "hbm", "achieved": alg_bytes / dfill_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": alg_bytes / dfill_s / 1e9 / hbm_peak, "peak_source": peak_src,
                        "bytes": "8 B/arc + 8 B/band row + 12 B/arc match + 16 B/pair (SURVEY 8d R2)"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * elapsed / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic",
            "config": {"workload": workload + ", split over the GPUs by estimated arc-match cost; score only", "flags": FLAGS_STR,
                       "pairs_per_step": len(pairs), "sub_batch": pl["sub_batch"],
                       "l2": "inputs larger than L2: the job's arc-match / task tables hold %.0f MB in HBM (L2: 126 MB)" % (resident_total / 1e6),
                       "bands": "derived inside every job (value: before the timed region, e2e: inside it): GPU FP64 envelope screening decided %d pairs, %d re-checked on the host in long double" % (env_stats[0], env_stats[1]),
                       "timing": "host clock between device-wide synchronisations (+ barrier), max over ranks; device_ms_per_step = CUDA-event time of the kernels",
                       "resident_prep_s": prep_s},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_total // K, "d2h_bytes_per_step": d2h_total // K,
                    "ms_per_step": 1e3 * e2e_elapsed / K,
                    "scope": "per job: context, parse of %d PP files, sharding, H2D, envelope bands, device build, D fill, top level, D2H, gather; the share is worked off by %d contexts (host thread + stream each)" % (len(paths), args.lanes)},
            "gpu_launches": launches_total,
            "roofline": roofline, "roofline_hbm": roofline_hbm,
            "device_ms_per_step": kernel_ms_max / K, "dfill_ms_per_step": dfill_ms_max / K,
            "gcups_all_gpus": cells_total * K / (dfill_ms_max / 1e3) / 1e9,
            "score_checksum": score_checksum(full), "resident_vs_e2e_score_mismatches": mismatch,
        }
        if traced_elapsed is not None:
            line["e2e_traced"] = {"value": args.traced_steps * len(pairs) / traced_elapsed, "unit": UNIT, "steps": args.traced_steps,
                                  "scope": "as e2e, plus the device traceback of every pair and the alignment edges D2H"}
        if world == 1:
            # bounded CPU sample of the same workload on all host cores
            n_cpu = args.cpu_sample or 2 * cores
            sample = pairs[:n_cpu]
            ref = CpuReference(cores)
            cpu_scores, cpu_wall = ref.align(paths, sample)
            ref.close()
            line["cpu_baseline"] = {"value": len(sample) / cpu_wall, "unit": UNIT, "cores": cores, "kind": ref.kind,
                                    "sample": "the first %d pairs of the job, one process per core, %.1f s wall, reference path incl. PP parse, envelope, arc matches, DP, traceback" % (len(sample), cpu_wall),
                                    "scores_match_gpu": all(cpu_scores[p] == full[k] for k, p in enumerate(sample))}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
