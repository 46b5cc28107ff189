#!/usr/bin/env python
"""profiles/ncu_kernels.json + profiles/<tag>_ncu_kernels.txt from ONE gpurun call at HEAD (tools/run_gpu.sh <tag> launches ncuk):
   gpurun_out/<tag>/launches.csv      ncu --metrics gpu__time_duration.sum over one warm-up + one timed bench step
   gpurun_out/<tag>/ncu_<kernel>.ncu-rep   ncu --set full of one launch of every hot kernel (same command, same HEAD)
bench.py reads the JSON: per kernel and step, measured DRAM bytes (dram__bytes_read + write of the captured launch x launches
per step), its share of the step, and the counter that binds it."""
import csv, json, os, subprocess, sys
from collections import defaultdict

tag = sys.argv[1]
src = os.path.join("gpurun_out", tag)
lines = [ln for ln in open(os.path.join(src, "launches.csv")) if ln.startswith('"')]
tot = defaultdict(lambda: [0, 0.0])
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", "")); unit = r["Metric Unit"]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    name = r["Kernel Name"].split("(")[0].replace("void ", "").replace("dsk::", "").split("<")[0]
    tot[name][0] += 1; tot[name][1] += us
steps = 3                                               # bench.py --steps 2 --warmup 1 under the launch list (tools/run_gpu.sh launches)
allus = sum(v[1] for v in tot.values())
KEYS = {"time_us": "gpu__time_duration.sum", "dram_read": "dram__bytes_read.sum", "dram_write": "dram__bytes_write.sum",
        "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "issue_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active", "alu_pct": "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex_pct": "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "occupancy_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
        "barrier_stall_per_issue": "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "long_sb_stall_per_issue": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "short_sb_stall_per_issue": "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "math_throttle_per_issue": "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "lg_throttle_per_issue": "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"}


def unit_scale(u):
    return {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3}.get(u, 1)


kern = {}
for fn in sorted(os.listdir(src)):
    if not (fn.startswith("ncu_") and fn.endswith(".ncu-rep")):
        continue
    p = subprocess.run(["ncu", "-i", os.path.join(src, fn), "--page", "raw", "--csv"], capture_output=True, text=True)
    rd = list(csv.reader(p.stdout.splitlines()))
    if len(rd) < 3:
        continue
    hdr, units, vals = rd[0], rd[1], rd[2]
    name = vals[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("dsk::", "").split("<")[0]
    m = {}
    for k, col in KEYS.items():
        if col in hdr:
            i = hdr.index(col)
            try:
                m[k] = float(vals[i].replace(",", "")) * (unit_scale(units[i]) if k in ("time_us", "dram_read", "dram_write") else 1)
            except ValueError:
                pass
    n, us = tot.get(name, [0, 0.0])
    per_step = n / steps
    dram = m.get("dram_read", 0) + m.get("dram_write", 0)
    bind = max((("SM issue", m.get("issue_pct", 0)), ("ALU pipe", m.get("alu_pct", 0)), ("DRAM", m.get("dram_pct", 0)), ("L1/shared", m.get("l1tex_pct", 0)),
                ("L2", m.get("lts_pct", 0))), key=lambda x: x[1])
    kern[name] = {"launches_per_step": per_step, "share_of_step_pct": 100 * us / allus if allus else None,
                  "captured_launch_us": m.get("time_us"), "dram_bytes_per_launch": dram, "dram_bytes_per_step": dram * per_step,
                  "dram_pct_of_peak": m.get("dram_pct"), "sm_throughput_pct": m.get("sm_pct"), "issue_active_pct": m.get("issue_pct"),
                  "alu_pipe_pct": m.get("alu_pct"), "l1tex_pct": m.get("l1tex_pct"), "l2_pct": m.get("lts_pct"), "occupancy_pct": m.get("occupancy_pct"),
                  "barrier_stall_per_issue": m.get("barrier_stall_per_issue"), "long_scoreboard_per_issue": m.get("long_sb_stall_per_issue"),
                  "short_scoreboard_per_issue": m.get("short_sb_stall_per_issue"), "math_throttle_per_issue": m.get("math_throttle_per_issue"),
                  "binding": {"resource": bind[0], "pct_of_peak": bind[1]}}
head = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
out = {"tag": "%s_ncu_kernels.txt" % tag, "head_at_summary": head,
       "workload": {"kmer_size": 31, "genome": 5000000, "coverage": 100},
       "note": "one gpurun call: launch list (gpu__time_duration.sum, --clock-control none) + one `ncu --set full` launch per hot kernel; per-launch "
               "times are cold-cache and serialised, compare shares; dram bytes are measured, per launch",
       "launch_shares": {k: {"launches": v[0], "total_us": v[1], "share_pct": 100 * v[1] / allus} for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])},
       "kernels": kern}
os.makedirs("profiles", exist_ok=True)
json.dump(out, open("profiles/ncu_kernels.json", "w"), indent=1)
with open("profiles/%s_ncu_kernels.txt" % tag, "w") as f:
    f.write("# %s\n" % out["note"])
    f.write("%-22s %8s %10s %7s | full capture of one launch: %9s %10s %6s %6s %6s %6s %6s %6s | binding\n" % (
        "kernel", "launches", "total_us", "share", "time_us", "dram_MB", "dram%", "sm%", "issue%", "alu%", "l1%", "occ%"))
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        m = kern.get(k)
        f.write("%-22s %8d %10.1f %6.1f%% | " % (k[:22], v[0], v[1], 100 * v[1] / allus))
        if m:
            f.write("%37.1f %10.1f %6.1f %6.1f %6.1f %6.1f %6.1f %6.1f | %s %.0f%%" % (m["captured_launch_us"] or 0, m["dram_bytes_per_launch"] / 1e6, m["dram_pct_of_peak"] or 0,
                    m["sm_throughput_pct"] or 0, m["issue_active_pct"] or 0, m["alu_pipe_pct"] or 0, m["l1tex_pct"] or 0, m["occupancy_pct"] or 0,
                    m["binding"]["resource"], m["binding"]["pct_of_peak"]))
        f.write("\n")
    f.write("%-22s %8d %10.1f\n" % ("TOTAL", sum(v[0] for v in tot.values()), allus))
print(open("profiles/%s_ncu_kernels.txt" % tag).read())
