#!/usr/bin/env python
"""Extracts the judged metrics of one kernel from an `ncu --page raw --csv` export into a small metric,unit,value CSV
(the format of profiles/*_ncu_*.csv).  usage: ncu_extract.py raw.csv > out.csv"""
import csv
import sys

KEEP = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__occupancy_limit_warps",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second")
rows = list(csv.reader(open(sys.argv[1])))
h, u, v = rows[0], rows[1], rows[2]
w = csv.writer(sys.stdout)
w.writerow(["metric", "unit", "value"])
w.writerow(["Kernel Name", "", v[h.index("Kernel Name")]])
for name in KEEP:
    if name in h:
        w.writerow([name, u[h.index(name)], v[h.index(name)]])
for i, name in enumerate(h):
    if name.startswith("smsp__average_warps_issue_stalled") and name.endswith("_per_issue_active.ratio"):
        w.writerow([name, u[i], v[i]])
