#!/usr/bin/env python3
"""Markdown table of the per-launch metrics that profiles/ quotes, from `ncu --page raw --csv`
output (full set) or from a `--metrics ... --csv` log:  tools/ncu_table.py raw.csv [raw2.csv ...]"""
import csv
import sys

COLS = [("gpu__time_duration.sum", "µs", 1.0),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %", 1.0),
        ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem pipe: TC %", 1.0),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem pipe: LSU %", 1.0),
        ("dram__bytes_read.sum", "DRAM rd MB", 1.0),
        ("dram__bytes_write.sum", "DRAM wr MB", 1.0),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %", 1.0),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %", 1.0),
        ("launch__grid_size", "grid", 1.0)]


def short(name):
    name = name.replace("void ", "").replace("dincae::", "")
    return name.split("(")[0][:44]


def rows_of(path):
    rows = list(csv.reader(open(path, errors="replace")))
    # full-set raw page: header row 0, units row 1; --metrics log: long format (one metric per row)
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hdr]
    if "Metric Name" in h:                                  # long format
        iK, iM, iV, iU, iID = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit"), h.index("ID")
        out = {}
        for r in rows[hdr + 1:]:
            if len(r) <= iV:
                continue
            d = out.setdefault(r[iID], {"Kernel Name": r[iK]})
            v = r[iV].replace(",", "")
            try:
                v = float(v)
            except ValueError:
                continue
            u = r[iU]
            if u in ("ns", "nsecond"):
                v /= 1e3
            elif u in ("ms", "msecond"):
                v *= 1e3
            if u in ("byte",):
                v /= 1e6
            elif u in ("Kbyte",):
                v /= 1e3
            elif u in ("Gbyte",):
                v *= 1e3
            d[r[iM]] = v
        return list(out.values())
    units = rows[hdr + 1]
    out = []
    for r in rows[hdr + 2:]:
        if len(r) < len(h):
            continue
        d = {"Kernel Name": r[h.index("Kernel Name")]}
        for i, n in enumerate(h):
            for key, _, _ in COLS:
                if n == key or n.endswith("." + key) or n.endswith(key):
                    try:
                        v = float(r[i].replace(",", ""))
                    except ValueError:
                        continue
                    u = units[i]
                    if key.startswith("gpu__time"):
                        v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
                    if key.startswith("dram__bytes"):
                        v = {"byte": v / 1e6, "Kbyte": v / 1e3, "Mbyte": v, "Gbyte": v * 1e3}.get(u, v)
                    d.setdefault(key, v)
        out.append(d)
    return out


for path in sys.argv[1:]:
    print("\n`%s`\n" % path)
    print("| # | kernel | " + " | ".join(c[1] for c in COLS) + " |")
    print("|---|---|" + "---:|" * len(COLS))
    for k, d in enumerate(rows_of(path)):
        cells = []
        for key, _, _ in COLS:
            v = d.get(key)
            cells.append("" if v is None else ("%.0f" % v if key == "launch__grid_size" else "%.1f" % v))
        print("| %d | %s | %s |" % (k, short(d["Kernel Name"]), " | ".join(cells)))
