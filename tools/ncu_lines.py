#!/usr/bin/env python
"""per CUDA source line: executed warp instructions and stall samples from an .ncu-rep (source page, cuda,sass view)"""
import csv, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
p = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True)
rows = list(csv.reader(p.stdout.splitlines()))
out = []; fname = ""; h = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No":
        h = r; ci = {n: i for i, n in enumerate(h)}; continue
    if h is None or len(r) < len(h) or r[0] == "":
        continue
    try:
        out.append((fname, int(r[0]), r[1].strip(), int(r[ci["# Samples"]] or 0), int(r[ci["Instructions Executed"]] or 0),
                    int(r[ci["stall_barrier"]] or 0), int(r[ci["stall_short_sb"]] or 0), int(r[ci["stall_wait"]] or 0), int(r[ci["stall_long_sb"]] or 0),
                    int(r[ci["stall_math"]] or 0), int(r[ci["stall_branch_resolving"]] or 0)))
    except ValueError:
        pass
ts = sum(o[3] for o in out); ti = sum(o[4] for o in out)
print("total samples %d, warp instructions %d" % (ts, ti))
print("%-16s %5s %7s %7s | bar shortsb wait longsb math branch | source" % ("file", "line", "samp%", "inst%"))
for o in sorted(out, key=lambda o: -o[3])[:topn]:
    print("%-16s %5d %6.2f%% %6.2f%% | %d %d %d %d %d %d | %s" % (o[0], o[1], 100.0 * o[3] / max(1, ts), 100.0 * o[4] / max(1, ti), o[5], o[6], o[7], o[8], o[9], o[10], o[2][:110]))
