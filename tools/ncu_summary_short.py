#!/usr/bin/env python3
"""Print the handful of ncu metrics that decide what bounds the kernel.  usage: ncu_summary_short.py <rep> [...]"""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_read.sum.per_second',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'sm__cycles_elapsed.avg.per_second', 'launch__registers_per_thread', 'launch__grid_size',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed_op_ldgsts.sum']
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h = rows[0]
    for r in rows[2:]:
        print("==", rep, r[h.index('Kernel Name')][:70])
        for k in WANT:
            if k in h:
                print(f"  {k:75s} {r[h.index(k)]} {rows[1][h.index(k)]}")
        for i, k in enumerate(h):
            if 'issue_stalled' in k and 'per_issue_active' in k and float(r[i] or 0) > 0.05:
                print(f"  stall {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):40s} {r[i]}")
