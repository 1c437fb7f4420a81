#!/usr/bin/env python3
"""Summarise ncu outputs for profiles/: (1) a launch list CSV (--metrics gpu__time_duration.sum) -> per-kernel shares;
(2) an .ncu-rep (--set full) -> the metrics DESIGN.md / bench.py quote. Usage:
    tools/ncu_summary.py launches <launches.csv>
    tools/ncu_summary.py report <file.ncu-rep>"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_bytes.sum", "l1tex__t_sector_hit_rate.pct",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum",
        "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "launch__local_size" if False else "smsp__sass_average_branch_targets_threads_uniform.pct"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        name = r[ki].split("(")[0]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# per-kernel device time from {path} (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:70s} launches={v[0]:5d} total_ms={v[1] / 1e3:10.3f} share={100 * v[1] / tot:5.1f}% avg_us={v[1] / v[0]:9.1f}")


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    print(f"# selected metrics from {path} (ncu --set full --clock-control none), one column per captured launch")
    kcol = hdr.index("Kernel Name")
    print("kernel:", [r[kcol].split("(")[0] for r in rows[2:]])
    for key in KEYS:
        if key in hdr:
            i = hdr.index(key)
            print(f"{key:90s} [{rows[1][i]}] {[r[i] for r in rows[2:]]}")


def traffic(path, workload=None, note=None):
    """mean dram__bytes_read.sum + dram__bytes_write.sum per launch of every captured kernel; with a workload name the
    entry of that workload in profiles/traffic.json is replaced (bench.py reads it as roofline.traffic)"""
    import json
    import re
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    kcol, rd, wr = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    agg = collections.defaultdict(list)
    for r in rows[2:]:
        name = re.sub(r"^void\s+|<.*$|\(.*$", "", r[kcol]).split("::")[-1]
        agg[name].append(float(r[rd].replace(",", "")) * scale[units[rd]] + float(r[wr].replace(",", "")) * scale[units[wr]])
    res = {"source": f"{path}: ncu --set full --clock-control none, mean over the captured launches (steady-state waves of bench.py)",
           "kernels": {k: int(sum(v) / len(v)) for k, v in agg.items()}}
    if note:
        res["source"] = res["source"].replace(": ncu", f" ({note}): ncu", 1)
    print(json.dumps(res, indent=1))
    if workload:
        import pathlib
        f = pathlib.Path(__file__).resolve().parent.parent / "profiles" / "traffic.json"
        t = json.loads(f.read_text()) if f.exists() else {}
        t[workload] = res
        f.write_text(json.dumps(t, indent=1) + "\n")


if __name__ == "__main__":
    {"launches": launches, "report": report, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])
