#!/usr/bin/env python3
"""Per-kernel durations of one training step from an ncu launch list (gpu__time_duration)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
i0 = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
h = rows[i0]; ix = {n: i for i, n in enumerate(h)}
seq = []
for r in rows[i0 + 1:]:
    if len(r) < len(h): continue
    v = float(r[ix['Metric Value']].replace(',', '')); unit = r[ix['Metric Unit']]
    v = v / 1000.0 if unit == 'ns' else (v * 1000.0 if unit == 'ms' else v)
    seq.append((r[ix['Kernel Name']], v))
idx = [i for i, (n, _) in enumerate(seq) if 'build_batch' in n]
k = int(sys.argv[2]) if len(sys.argv) > 2 else 4
tot = 0; agg = {}
for n, v in seq[idx[k]:idx[k + 1]]:
    short = n.split('(')[0].replace('void ', '').replace('dincae::', '')
    if len(sys.argv) <= 3: print("%8.1f us  %s" % (v, short[:70]))
    tot += v; a = agg.setdefault(short, [0, 0.0]); a[0] += 1; a[1] += v
print("step total us %.1f" % tot)
for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%8.1f us  x%-2d %s" % (v, c, n[:70]))
