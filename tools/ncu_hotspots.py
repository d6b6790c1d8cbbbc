#!/usr/bin/env python3
"""Hottest SASS lines (warp-stall samples) of one kernel of an .ncu-rep captured with
--import-source on:  tools/ncu_hotspots.py report.ncu-rep [kernel-index] [top-n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
kid = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
sec = rows[starts[kid]:starts[kid + 1]]
print(sec[0][1][:100])
h = sec[1]
iN, iE = h.index("# Samples"), h.index("Instructions Executed")
body = [r for r in sec[2:] if len(r) > max(iN, iE)]
data = [(int(r[iN] or 0), int(r[iE] or 0), k, r[1].strip()) for k, r in enumerate(body)]
tot = sum(d[0] for d in data)
print("total samples", tot, "warp instructions", sum(d[1] for d in data))
for n, e, k, src in sorted(data, reverse=True)[:top]:
    print("%6d %5.1f%% exec=%9d  #%4d %s" % (n, 100.0 * n / max(tot, 1), e, k, src[:100]))
