#!/usr/bin/env python
"""Benchmark of the B200 pairwise sequence-structure alignment path.

Workload (BASELINE.json configs[4], the configuration the metric is quoted on): the all-vs-all guide-tree stage of
mlocarna for 512 synthetic RNAs x 300 nt (130,816 pairs), flags as mlocarna passes them
(--noLP --max-diff-am 30 --struct-weight 200 --min-prob 0.001), score-only mode (the caller consumes
only the score, mlocarna:3516-3527).  One "step" aligns one slice of --batch pairs of that pair list; ranks take
disjoint slices (weak scaling: the slice per GPU is fixed), there is no data-path collective, and rank 0 gathers the
score slices with a single NCCL gather at the end of the timed region.

  value  pairs/s over K steps with the step batches already resident in HBM (arc matches, tasks, bands uploaded)
  e2e    pairs/s through the C ABI from host buffers: sequences + base pairs + bands -> lb200_run (host build,
         H2D, kernels, D2H of the scores), everything inside the timed region
  --impl reference   the reference's own CPU aligner (oracle/_ref/ref_harness, compiled from the unmodified
         reference sources; the oracle port if that binary is absent), one process per host core

Synthetic data: locarna_b200/synth.py (ViennaRNA is not available in this image), seed = 1000*5 + index.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLAGS = {"noLP": True, "max-diff-am": 30, "struct-weight": 200, "min-prob": 0.001}
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
    with cf.ProcessPoolExecutor(max_workers=workers) as ex:
        return list(ex.map(_gen_one, jobs, chunksize=4))


def read_pp(path):
    """Minimal PP 2.0 reader for the benchmark's own synthetic files -> (name, seq, [(i, j, p)])."""
    name, seq, pairs, sect = None, "", [], 0
    for line in open(path):
        if not line.strip() or line[0].isspace():
            continue
        if line.startswith("#"):
            if line.startswith("#SECTION BASEPAIRS"):
                sect = 1
            continue
        t = line.split()
        if sect == 0:
            name, seq = t[0], seq + t[1]
        else:
            pairs.append((int(t[0]), int(t[1]), float(t[2])))
    return name, seq, pairs


def all_vs_all(n):
    from locarna_b200 import allpairs
    return allpairs.all_vs_all(n)  # mlocarna pair order (mlocarna:3577-3604)


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


def cpu_reference(paths, pairs, cores, per_core):
    """Align len(pairs) pairs with the reference's CPU aligner, one process per core. Returns (pairs/s, kind, scores)."""
    from oracle import oracle as O
    kind = "reference" if O.have_ref() else "port"
    if kind == "port":
        O.lib()
    file_pairs = [(paths[a], paths[b]) for a, b in pairs]
    chunks = [file_pairs[k::cores] for k in range(cores)]
    chunks = [c for c in chunks if c]
    t = time.time()
    with cf.ProcessPoolExecutor(max_workers=cores) as ex:
        out = list(ex.map(_ref_worker, [(kind == "reference", FLAGS, c) for c in chunks]))
    wall = time.time() - t
    scores = {}
    for k, (sc, _) in enumerate(out):
        for idx, s in enumerate(sc):
            scores[pairs[k + idx * cores]] = s
    return len(file_pairs) / wall, kind, scores, wall


# ----------------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="pairs per step and GPU")
    ap.add_argument("--seqs", type=int, default=512)
    ap.add_argument("--len", type=int, default=300)
    ap.add_argument("--cpu-sample", type=int, default=0, help="pairs of the CPU baseline sample (default: 2 per core)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    K, W, B = args.steps, max(args.warmup, 0), args.batch

    if args.impl == "reference":
        if rank != 0:
            return 0
        paths = make_inputs(args.seqs, args.len, cores)
        pairs = all_vs_all(args.seqs)
        per_step = args.cpu_sample or cores
        need = (W + K) * per_step
        sample = pairs[:need]
        times = []
        for s in range(W + K):
            chunk = sample[s * per_step:(s + 1) * per_step]
            v, kind, _, wall = cpu_reference(paths, chunk, cores, 1)
            if s >= W:
                times.append(wall)
        total = sum(times)
        value = K * per_step / total
        line = {
            "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64",
            "data": "synthetic",
            "config": {"workload": "cfg5 mlocarna guide-tree stage, all-vs-all %d x %d nt, step = %d pairs on %d host cores" % (args.seqs, args.len, per_step, cores),
                       "flags": "--noLP --max-diff-am 30 --struct-weight 200 --min-prob 0.001", "pairs_per_step": per_step},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": "%d pairs per step, one process per core, reference path incl. PP parse, envelope, arc matches, DP, traceback" % per_step},
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
    from locarna_b200 import capi

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    if rank == 0:
        paths = make_inputs(args.seqs, args.len, cores)
    if dist is not None:
        dist.barrier()
    paths = make_inputs(args.seqs, args.len, 1) if rank != 0 else paths
    seqs = [read_pp(p) for p in paths]
    pairs = all_vs_all(args.seqs)
    n_steps = W + K
    # rank r takes the global step slices r, r + world, ... ; a second, disjoint set of slices feeds the e2e leg
    def slice_of(g):
        lo = (g * B) % max(1, len(pairs) - B)
        return pairs[lo:lo + B]
    my_steps = [slice_of(s * world + rank) for s in range(n_steps)]
    e2e_steps = [slice_of((n_steps + s) * world + rank) for s in range(n_steps)]

    def new_ctx(step_pairs, bands=None):
        ctx = capi.Context(local_rank, FLAGS)
        used = sorted({x for p in step_pairs for x in p})
        ids = {}
        for s in used:
            name, seq, bp = seqs[s]
            ids[s] = ctx.add_seq(name, seq, bp)
        for k, (a, b) in enumerate(step_pairs):
            ctx.add_pair(ids[a], ids[b], None if bands is None else bands[k])
        return ctx

    # ---- resident leg: build + upload every step batch before the timed region
    t0 = time.time()
    ctxs = []
    resident_bytes = 0
    for sp in my_steps:
        ctx = new_ctx(sp)
        ctx.upload()
        # S-order entry 16 B + L-order record 20 B per arc match, 20 B per D-fill task (built on the device, dev_types.h)
        resident_bytes = max(resident_bytes, sum(36 * ctx.info(k).n_arcmatches + 20 * ctx.info(k).n_tasks for k in range(len(sp))))
        ctxs.append(ctx)
    prep_s = time.time() - t0
    for s in range(W):
        ctxs[s].run()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t0 = time.time()
    for s in range(W, n_steps):
        ctxs[s].run()
    barrier()
    elapsed = time.time() - t0
    clocks = sampler.stop()
    kernel_ms = sum(ctxs[s].kernel_ms for s in range(W, n_steps))
    dfill_ms = sum(ctxs[s].dfill_ms for s in range(W, n_steps))
    dfill_launches = sum(ctxs[s].dfill_launches for s in range(W, n_steps))
    launches = sum(ctxs[s].launches for s in range(W, n_steps))
    cells = terms = am = arcs = rows = 0
    for s in range(W, n_steps):
        for k in range(len(my_steps[s])):
            inf = ctxs[s].info(k)
            cells += inf.cells; terms += inf.terms; am += inf.n_arcmatches; arcs += inf.n_arcsA + inf.n_arcsB; rows += inf.lenA + 1
    my_scores = [ctxs[s].scores() for s in range(n_steps)]
    for c in ctxs:
        c.close()

    # ---- e2e leg: host buffers -> scores on the host, everything timed.
    # The job's sequences (512 RNAs: sequence + base pairs) are uploaded once per job before the steps, as a caller
    # of the all-vs-all stage would do; every step passes its pair list as host buffers through the C ABI
    # (lb200_clear_pairs, lb200_pair_add with no band, lb200_run = band derivation (GPU envelope screening + exact host
    # re-check of uncertain pairs) + device build + D fill + top level + D2H of the scores).
    env_stats = [0, 0]
    e2e_ctx = capi.Context(local_rank, FLAGS)
    e2e_ids = [e2e_ctx.add_seq(*seqs[s]) for s in range(len(seqs))]

    def e2e_step(s):
        e2e_ctx.clear_pairs()
        for k, (a, b) in enumerate(e2e_steps[s]):
            e2e_ctx.add_pair(e2e_ids[a], e2e_ids[b], None)
        e2e_ctx.run()
        dev, host = e2e_ctx.envelope_stats()
        env_stats[0] += dev; env_stats[1] += host
        return e2e_ctx.scores(), e2e_ctx.h2d_bytes, e2e_ctx.d2h_bytes

    for s in range(W):
        e2e_step(s)
    barrier()
    t0 = time.time()
    h2d = d2h = 0
    gathered = None
    for s in range(W, n_steps):
        sc, a, b = e2e_step(s)
        h2d += a; d2h += b
    if dist is not None:  # the one collective of the path: score slices of the last step to rank 0 (NCCL gather)
        from locarna_b200 import allpairs
        gathered = allpairs.gather_scores(dist, list(range(rank * B, rank * B + len(sc))), sc, world * B, device="cuda")
    barrier()
    e2e_elapsed = time.time() - t0
    e2e_ctx.close()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    elapsed = max_over_ranks(elapsed)
    e2e_elapsed = max_over_ranks(e2e_elapsed)
    total_pairs = K * B * world
    value = total_pairs / elapsed
    e2e_value = total_pairs / e2e_elapsed
    launches_total = int(sum_over_ranks(launches))

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        peak_src = "measured" if "hbm_gbs" in peaks else "fallback"
        # dominant kernel = dfill_kernel; figures per launch (rank 0), DESIGN.md "roofline"
        alg_bytes = 8 * arcs + 8 * rows + 12 * am + 16 * K * B
        # SURVEY 8d R1 counts 9 int ops per cell update + 2 per arc-match term. The kernel folds only the entries whose source lies in
        # the box (a prefix of each anti-diagonal's list) and does not tally them, so the term part is left out: a conservative count.
        ops = 9 * cells
        dfill_s = dfill_ms / 1e3
        sm_mhz = clocks.get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
        alu_peak = 148 * 128 * sm_mhz * 1e6
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("dfill_dram_bytes_per_launch")
        except (OSError, ValueError):
            pass
        roofline = {"bound": "hbm", "achieved": alg_bytes / dfill_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": alg_bytes / dfill_s / 1e9 / hbm_peak, "traffic": traffic, "kernel": "dfill_dep_kernel" if dfill_launches <= K else "dfill_kernel",
                    "launch_ms": dfill_ms / max(1, dfill_launches), "peak_source": peak_src,
                    "note": "max-plus DP: the binding roof is integer ALU issue, see roofline_alu"}
        roofline_alu = {"bound": "int_alu", "achieved": ops / dfill_s / 1e12, "peak": alu_peak / 1e12, "unit": "Tintop/s",
                        "frac": ops / dfill_s / alu_peak, "ops": "9*cells (SURVEY 8d R1; the 2 ops per folded arc-match term are not counted)",
                        "gcups": cells / dfill_s / 1e9, "sm_mhz": sm_mhz}
        # bounded CPU sample of the same workload, all host cores
        n_cpu = args.cpu_sample or 2 * cores
        sample = my_steps[W][:n_cpu]
        cpu_v, kind, cpu_scores, cpu_wall = cpu_reference(paths, sample, cores, 2)
        parity = all(cpu_scores[p] == my_scores[W][k] for k, p in enumerate(sample))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * elapsed / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic",
            "config": {"workload": "cfg5 mlocarna guide-tree stage, all-vs-all %d x %d nt (%d pairs), step = %d-pair slice per GPU, score only" % (args.seqs, args.len, len(pairs), B),
                       "flags": "--noLP --max-diff-am 30 --struct-weight 200 --min-prob 0.001", "pairs_per_step": B,
                       "l2": "inputs larger than L2: one step batch holds %.0f MB of arc-match / task tables in HBM (L2: 126 MB)" % (resident_bytes / 1e6),
                       "bands": "derived inside every run (value: before the timed region, e2e: inside it): GPU FP64 envelope screening decided %d pairs, %d re-checked on the host in long double" % (env_stats[0], env_stats[1]),
                       "resident_prep_s": prep_s},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d // K, "d2h_bytes_per_step": d2h // K},
            "gpu_launches": launches_total,
            "roofline": roofline, "roofline_alu": roofline_alu,
            "cpu_baseline": {"value": cpu_v, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": "%d pairs of the timed slice, one process per core, %.1f s wall" % (len(sample), cpu_wall),
                             "scores_match_gpu": parity},
            "kernel_ms_per_step": kernel_ms / K,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
