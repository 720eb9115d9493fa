#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the few numbers DESIGN.md / profiles/ quote.
usage: scripts/ncu_summary.py gpurun_out/prof_x.ncu-rep"""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__warps_eligible.avg.per_cycle_active", "launch__grid_size", "launch__block_size",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed_pipe_fp64.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    print("==", d.get("Kernel Name"), "id", d.get("ID"))
    for w in want:
        if w in d:
            print(f"  {w} = {d[w]} {units[hdr.index(w)]}")
    st = sorted(((float(v), h) for h, v in d.items()
                 if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and v),
                reverse=True)[:6]
    for v, h in st:
        print(f"  stall {h.split('stalled_')[1].split('_per_issue')[0]} = {v:.2f}")
