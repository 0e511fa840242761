"""Turns ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

  python tools/summarize_profiles.py launches <out.md> <launches.csv> "<command that was profiled>"
  python tools/summarize_profiles.py kernel   <out.md> <report.ncu-rep> <kernel regex> [traffic.json]
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
        "smsp__sass_inst_executed_op_tmem_ldt.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_active.avg"]


def launches(out, path, cmd):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("==")) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ki].split("(")[0].replace("void ", "")
        v = float(r[vi].replace(",", ""))
        ms = v / 1e6 if r[ui].startswith("n") else (v / 1e3 if r[ui].startswith("u") else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list of `{cmd}`\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare shares, not absolutes).\n\n")
        f.write(f"{len(rows) - 1} launches, {tot:.3f} ms total device time\n\n| ms | share | launches | kernel |\n|---:|---:|---:|---|\n")
        for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {ms:.3f} | {100 * ms / tot:.1f}% | {c} | `{k[:110]}` |\n")


def kernel(out, rep, pattern, traffic_out=None):
    kre = re.compile(pattern)
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True)
    rr = list(csv.reader(raw.splitlines()))
    h, u = rr[0], rr[1]
    kcol = h.index("Kernel Name")
    sel = [r for r in rr[2:] if kre.search(r[kcol])]
    traffic = None
    with open(out, "w") as f:
        f.write(f"# `ncu --set full --clock-control none --import-source on` of kernels matching `{pattern}` ({os.path.basename(rep)})\n\n")
        for n, r in enumerate(sel):
            f.write(f"## launch {n}: `{r[kcol][:140]}`\n\n| metric | unit | value |\n|---|---|---:|\n")
            d = dict(zip(h, r))
            for w in WANT:
                if w in d:
                    f.write(f"| {w} | {u[h.index(w)]} | {d[w]} |\n")
            f.write("\n")
            try:
                def to_bytes(key):
                    v = float(d[key].replace(",", ""))
                    unit = u[h.index(key)].lower()
                    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1)
                traffic = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
            except Exception:
                pass
    if traffic_out and traffic is not None:
        json.dump({"kernel": pattern, "dram_bytes_per_launch": traffic, "source": os.path.basename(rep)}, open(traffic_out, "w"))


if __name__ == "__main__":
    os.makedirs("profiles", exist_ok=True)
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        kernel(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5] if len(sys.argv) > 5 else None)
    print("written", sys.argv[2])
