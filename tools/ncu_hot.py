#!/usr/bin/env python3
"""Top SASS instructions by stall samples for one kernel of an .ncu-rep.
usage: tools/ncu_hot.py report.ncu-rep <launch-index> [topN]"""
import csv, subprocess, sys, io
rep, kid = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", kid, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
print(lines[start - 1][:120])
rows = list(csv.reader(lines[start:]))
h = rows[0]
ix = {n: i for i, n in enumerate(h)}
stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
data = []
for k, r in enumerate(rows[1:]):
    if len(r) < len(h):
        continue
    data.append((int(r[ix["# Samples"]] or 0), k, r))
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
for smp, k, r in sorted(data, reverse=True)[:top]:
    st = sorted(((int(r[i] or 0), h[i]) for i in stall_cols), reverse=True)[:3]
    print("%5.1f%% #%4d %-70s exec=%s %s" % (100.0 * smp / max(tot, 1), k, r[ix["Source"]].strip()[:70],
          r[ix["Instructions Executed"]], " ".join("%s=%d" % (n[6:], v) for v, n in st if v)))
