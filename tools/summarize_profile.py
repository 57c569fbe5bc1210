#!/usr/bin/env python
"""Summarise gpurun_out/*.ncu-rep and launch lists into small text files under profiles/ (run in the build container)."""
import collections
import csv
import io
import json
import subprocess
import sys

KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_blocks", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_st.sum",
        "lts__t_sector_hit_rate.pct", "sm__issue_active.avg.pct_of_peak_sustained_elapsed"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def source_page(rep):
    """Warp-stall sampling and instruction mix of the first kernel in the report (ncu --import-source on)."""
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3 or "Source" not in rows[1]:
        return []
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def f(r, k):
        try:
            return float(r[ix[k]])
        except (ValueError, KeyError, IndexError):
            return 0.0
    inst = sum(f(r, "Instructions Executed") for r in data)
    thr = sum(f(r, "Predicated-On Thread Instructions Executed") for r in data)
    samples = sum(f(r, "# Samples") for r in data)
    lines = ["source page: %.4g warp instructions, %.4g predicated-on thread instructions (%.1f %% of 32 lanes)" % (inst, thr, 100 * thr / inst / 32)]
    stalls = sorted(((sum(f(r, k) for r in data) / samples, k) for k in hdr if k.startswith("stall_") and "Not" not in k), reverse=True)
    lines.append("warp-stall samples: " + ", ".join("%s %.1f %%" % (k[6:], 100 * v) for v, k in stalls[:9]))
    ops = collections.Counter()
    for r in data:
        t = r[ix["Source"]].split()
        if not t:
            continue
        op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
        ops[".".join(op.split(".")[:2]) if op.startswith(("LD", "ST", "ATOM")) else op.split(".")[0]] += f(r, "Instructions Executed")
    lines.append("instruction mix: " + ", ".join("%s %.1f %%" % (k, 100 * v / inst) for k, v in ops.most_common(14)))
    lines.append("")
    return lines


def main():
    rep, launches, out_txt, out_traffic = sys.argv[1:5]
    hdr, units, rows = raw(rep)
    lines = ["ncu --set full --clock-control none, kernel regex dfill; source: %s" % rep, ""]
    traffic = []
    for r in rows:
        for k in KEYS:
            if k in hdr:
                lines.append("%-90s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
        lines.append("")
        def f(k):
            v = float(r[hdr.index(k)].replace(",", "")); u = units[hdr.index(k)]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        traffic.append(f("dram__bytes_read.sum") + f("dram__bytes_write.sum"))
    lines += source_page(rep)
    agg = collections.defaultdict(lambda: [0, 0.0])
    rows2 = [r for r in csv.reader(open(launches)) if len(r) > 10]
    h2 = rows2[0]
    for r in rows2[1:]:
        try:
            v = float(r[h2.index("Metric Value")].replace(",", ""))
        except ValueError:
            continue
        name = r[h2.index("Kernel Name")].split("(")[0]
        agg[name][0] += 1; agg[name][1] += v
    tot = sum(v for _, v in agg.values())
    lines.append("launch list (ncu --metrics gpu__time_duration.sum, cold cache / serialised: compare shares): %s" % launches)
    for k, (n, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        lines.append("  %-72s n=%5d %10.3f ms %5.1f%%" % (k[:72], n, v / 1e6, 100 * v / tot))
    open(out_txt, "w").write("\n".join(lines) + "\n")
    json.dump({"dfill_dram_bytes_per_launch": sum(traffic) / max(1, len(traffic)), "source": rep, "launches_profiled": len(traffic)}, open(out_traffic, "w"))
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main()
