"""Turns gpurun_out/launches.csv (+ an ncu --set full report) into the tracked summaries under profiles/.
usage: python tools/summarize_profiles.py <round tag> <launches.csv> <report.ncu-rep> [kernel regex]"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

tag, launches, rep = sys.argv[1], sys.argv[2], sys.argv[3]
kre = re.compile(sys.argv[4] if len(sys.argv) > 4 else "plane_score")
os.makedirs("profiles", exist_ok=True)

rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
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
with open(f"profiles/{tag}_launches.md", "w") as f:
    f.write(f"# {tag}: ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --frames 0`\n\n")
    f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare shares).\n\n")
    f.write(f"{len(rows) - 1} launches, {tot:.3f} ms total device time\n\n| ms | share | launches | kernel |\n|---:|---:|---:|---|\n")
    for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| {ms:.3f} | {100 * ms / tot:.1f}% | {c} | `{k[:100]}` |\n")

raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True)
rr = list(csv.reader(raw.splitlines()))
h, u = rr[0], rr[1]
kcol = h.index("Kernel Name")
sel = [r for r in rr[2:] if kre.search(r[kcol])]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__cycles_active.avg"]
with open(f"profiles/{tag}_{kre.pattern}_ncu.md", "w") as f:
    f.write(f"# {tag}: `ncu --set full --clock-control none` of kernels matching `{kre.pattern}`\n\n")
    traffic = None
    for n, r in enumerate(sel):
        f.write(f"## launch {n}: `{r[kcol][:120]}`\n\n| metric | unit | value |\n|---|---|---:|\n")
        d = dict(zip(h, r))
        for w in want:
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
    if traffic is not None:
        json.dump({"kernel": kre.pattern, "dram_bytes_per_launch": traffic, "source": os.path.basename(rep)},
                  open(f"profiles/{tag}_{kre.pattern}_traffic.json", "w"))
print("profiles written")
