#!/usr/bin/env python
"""top SASS instructions by stall samples from an .ncu-rep (source page)"""
import csv, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
p = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True)
rows = list(csv.reader(p.stdout.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]; data = rows[hi + 1:]
ci = {n: h.index(n) for n in ("Source", "# Samples", "Instructions Executed", "Avg. Threads Executed")}
stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
tot = sum(int(r[ci["# Samples"]] or 0) for r in data)
print("total samples", tot, "instructions", len(data))
idx = sorted(range(len(data)), key=lambda i: -int(data[i][ci["# Samples"]] or 0))[:topn]
for i in sorted(idx):
    r = data[i]
    st = sorted(((int(r[c] or 0), h[c]) for c in stall_cols), reverse=True)[:2]
    print("%5d %6.2f%% exec=%-9s thr=%-5s %-70s %s" % (i, 100.0 * int(r[ci["# Samples"]]) / max(1, tot), r[ci["Instructions Executed"]],
          r[ci["Avg. Threads Executed"]], r[ci["Source"]].strip()[:70], " ".join("%s=%d" % (n, v) for v, n in st)))
