#!/usr/bin/env python
"""Compact per-launch summary of an `ncu --set full` report:  ncu -i X.ncu-rep --page raw --csv | this"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
H, U = rows[0], rows[1]
KEYS = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe%_active"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("lts__t_sector_hit_rate.pct", "l2_hit%"),
        ("smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "stall_long_sb%"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts")]
for r in rows[2:]:
    name = r[H.index("Kernel Name")].split("(")[0]
    parts = []
    for k, label in KEYS:
        if k in H:
            i = H.index(k)
            parts.append(f"{label}={r[i]}{U[i] if U[i] not in ('', '%') else ''}")
    print(f"{name:28s} " + "  ".join(parts))
