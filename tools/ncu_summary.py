#!/usr/bin/env python
"""Summarises ncu outputs brought back in gpurun_out/<tag>/ into profiles/<tag>_*.{csv,txt} (tracked)."""
import csv
import os
import subprocess
import sys
from collections import defaultdict

tag = sys.argv[1]
src = os.path.join("gpurun_out", tag)
dst = "profiles"
os.makedirs(dst, exist_ok=True)

# 1. launch list -> per-kernel totals and shares
rows = []
with open(os.path.join(src, "launches.csv")) as f:
    lines = [ln for ln in f if ln.startswith('"')]
rd = csv.DictReader(lines)
tot = defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    name = r["Kernel Name"].split("(")[0]
    tot[name][0] += 1
    tot[name][1] += us
allus = sum(v[1] for v in tot.values())
with open(os.path.join(dst, "%s_launches_summary.txt" % tag), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none ; one warm-up step + one timed step of bench.py\n")
    f.write("# per-launch times are cold-cache and serialised: compare SHARES\n")
    f.write("%-60s %8s %12s %8s\n" % ("kernel", "launches", "total_us", "share"))
    for name, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        f.write("%-60s %8d %12.1f %7.1f%%\n" % (name[:60], n, us, 100 * us / allus))
    f.write("%-60s %8d %12.1f\n" % ("TOTAL", sum(v[0] for v in tot.values()), allus))
print(open(os.path.join(dst, "%s_launches_summary.txt" % tag)).read())

# 2. key metrics of each full capture
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_lsu.sum", "smsp__cycles_active.avg", "launch__occupancy_limit_registers",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__sass_average_data_bytes_per_sector_mem_global_op_st.pct"]
with open(os.path.join(dst, "%s_ncu_full_summary.txt" % tag), "w") as out:
    for fn in sorted(os.listdir(src)):
        if not fn.endswith(".ncu-rep"):
            continue
        p = subprocess.run(["ncu", "-i", os.path.join(src, fn), "--page", "raw", "--csv"], capture_output=True, text=True)
        lines = p.stdout.splitlines()
        if len(lines) < 3:
            continue
        rd = list(csv.reader(lines))
        hdr, units, vals = rd[0], rd[1], rd[2]
        out.write("== %s : %s\n" % (fn, vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""))
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                out.write("   %-90s %s %s\n" % (k, vals[i], units[i]))
        out.write("\n")
print(open(os.path.join(dst, "%s_ncu_full_summary.txt" % tag)).read())

# 3. DRAM traffic per launch of every captured kernel -> profiles/ncu_traffic.json (bench.py's roofline.traffic reads it;
#    valid for the default bench workload the captures are taken on: C2, k=31)
import json
import re
tj = os.path.join(dst, "ncu_traffic.json")
traffic = json.load(open(tj)) if os.path.exists(tj) else {}
cur = None
MUL = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for ln in open(os.path.join(dst, "%s_ncu_full_summary.txt" % tag)):
    m = re.match(r"== prof_(\w+)\.ncu-rep", ln)
    if m:
        cur = m.group(1); traffic[cur] = {"tag": tag, "workload": "bench.py default (C2, k=31)"}
        continue
    f = ln.split()
    if cur and len(f) >= 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"):
        if f[0].startswith("dram"):
            traffic[cur][f[0]] = float(f[1]) * MUL.get(f[2], 1)
        else:
            traffic[cur]["ncu_duration"] = "%s %s" % (f[1], f[2])
json.dump(traffic, open(tj, "w"), indent=1, sort_keys=True)
