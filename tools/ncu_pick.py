#!/usr/bin/env python3
"""Print a few named metrics per kernel launch from an .ncu-rep (ncu -i ... --page raw --csv).
usage: tools/ncu_pick.py report.ncu-rep [extra_metric_substring ...]"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    extra = sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    for row in rows[2:]:
        d = dict(zip(h, row))
        print("=== %s  id=%s" % (d.get("Kernel Name", "?")[:90], d.get("ID")))
        for i, name in enumerate(h):
            short = name.split(".Triage")[-1]
            if any(name == w or name.endswith("." + w) for w in WANT) or any(e in name for e in extra):
                print("  %-80s %s %s" % (name[-80:], row[i], units[i]))
        st = [(float(row[i] or 0), name) for i, name in enumerate(h)
              if "warp_issue_stalled" in name and name.endswith("per_warp_active.pct")]
        st.sort(reverse=True)
        for v, name in st[:6]:
            print("  stall %-60s %.1f" % (name.replace("smsp__warp_issue_stalled_", "").replace("_per_warp_active.pct", ""), v))


if __name__ == "__main__":
    main()
