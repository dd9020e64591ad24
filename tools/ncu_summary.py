#!/usr/bin/env python
"""Prints selected metrics from `ncu -i X.ncu-rep --page raw --csv` (reads the CSV on stdin)."""
import csv
import sys

KEYS = sys.argv[1:] or [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
    "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak", "smsp__issue_active.avg.pct",
    "sm__inst_executed_pipe_fp64", "sm__pipe_fp64_cycles_active", "smsp__inst_executed.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared",
    "smsp__average_warps_issue_stalled", "smsp__average_warp_latency_issue_stalled",
    "lts__t_bytes.sum", "sm__throughput.avg.pct", "l1tex__throughput.avg.pct",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__pcsamp_warps_issue_stalled",
]
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:80])
    for h, u, v in zip(hdr, units, r):
        if any(k in h for k in KEYS):
            print(f"  {h:100s} {v} {u}")
