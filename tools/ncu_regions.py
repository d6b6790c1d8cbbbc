#!/usr/bin/env python3
"""Stall-reason totals per code region (loader / mma / epilogue) of a conv3x3_tc launch."""
import csv, subprocess, sys
rep, kid = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", kid, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.reader(lines[start:]))
h = rows[0]; ix = {n: i for i, n in enumerate(h)}
stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
data = [r for r in rows[1:] if len(r) >= len(h)]
src = [r[ix["Source"]] for r in data]
first_mma = next(i for i, t in enumerate(src) if "UTCHMMA" in t)
last_mma = max(i for i, t in enumerate(src) if "UTCHMMA" in t)
# loader = from first LDG.E.128 after setup to first_mma-ish; crude split using TRYWAIT positions
tw = [i for i, t in enumerate(src) if "TRYWAIT" in t]
print("trywait idx", tw, "first/last mma", first_mma, last_mma, "n", len(src))
bounds = {"setup+loader": (0, tw[1] if len(tw) > 1 else first_mma), "mma": (tw[1] if len(tw) > 1 else first_mma, last_mma + 40), "epilogue": (last_mma + 40, len(src))}
for name, (a, b) in bounds.items():
    tot = {}
    ninst = 0; smp = 0
    for r in data[a:b]:
        smp += int(r[ix["# Samples"]] or 0)
        ninst += int(r[ix["Instructions Executed"]] or 0)
        for i in stall_cols:
            v = int(r[i] or 0)
            if v: tot[h[i]] = tot.get(h[i], 0) + v
    print("%-14s sass[%d:%d] samples=%d warp-instr=%d  %s" % (name, a, b, smp, ninst,
          " ".join("%s=%d" % (k[6:], v) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:7])))
