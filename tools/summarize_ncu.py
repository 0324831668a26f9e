#!/usr/bin/env python3
"""Summarise an ncu report + launch list into profiles/ (run in the dev container).

    python tools/summarize_ncu.py gpurun_out/r1a r1a
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src, tag = sys.argv[1], sys.argv[2]
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
]

rep = os.path.join(src, "prof.ncu-rep")
lines = []
traffic = None
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    lines.append(f"# ncu --set full summary ({tag})\n")
    lines.append("Command: `ncu --set full --clock-control none --import-source on -k regex:flagstat_kernel "
                 "-s 3 -c 2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline` on one B200 "
                 "(workload: 824,541,892 HiSeqX-shaped records, 1,649,083,784 B).\n")
    for r in rows[2:]:
        lines.append(f"\n## {r[idx['Kernel Name']]}  (launch id {r[idx['ID']]})\n")
        lines.append("| metric | value | unit |\n|---|---|---|")
        for k in KEEP:
            if k in idx:
                lines.append(f"| `{k}` | {r[idx[k]]} | {units[idx[k]]} |")
        if traffic is None:
            def num(k):
                v = float(r[idx[k]].replace(",", ""))
                u = units[idx[k]].lower()
                mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
                return v * mult
            traffic = {
                "dram_bytes_per_launch": num("dram__bytes_read.sum") + num("dram__bytes_write.sum"),
                "dram_bytes_read": num("dram__bytes_read.sum"), "dram_bytes_write": num("dram__bytes_write.sum"),
                "algorithmic_bytes": 1649083784,
                "source": f"profiles/{tag}_ncu_summary.md (ncu --set full, one launch of {r[idx['Kernel Name']].split('(')[0].replace('void ', '')})",
                "kernel": r[idx['Kernel Name']].split('(')[0].replace('void ', ''),
            }
    with open(os.path.join(out_dir, f"{tag}_ncu_summary.md"), "w") as fh:
        fh.write("\n".join(lines) + "\n")
    if traffic:
        with open(os.path.join(out_dir, "ncu_traffic.json"), "w") as fh:
            json.dump(traffic, fh, indent=1)
            fh.write("\n")

ll = os.path.join(src, "launches.csv")
if os.path.exists(ll):
    rows = list(csv.reader(l for l in open(ll) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        t = float(r[vi].replace(",", ""))
        name = r[ki]
        # the counting kernel runs on two very different inputs inside bench.py: the resident
        # 1.65 GB shard (the timed steps, ~240 us) and 8-block groups of the stream leg (~12 us)
        if "flagstat_kernel" in name:
            name = ("[shard 1.65 GB] " if t > 100e3 else "[stream group 8 x 1,024,000 B] ") + name
        if "hbm_read_probe" in name:
            name = "[bench.py's read-only roofline probe, outside the timed steps] " + name
        key = (name, r[gi])
        agg.setdefault(key, []).append(t)
    tot = sum(sum(v) for v in agg.values())
    with open(os.path.join(out_dir, f"{tag}_launches.md"), "w") as fh:
        fh.write(f"# ncu launch list ({tag})\n\n`ncu --metrics gpu__time_duration.sum --clock-control none -c 200 "
                 "python bench.py --steps 3 --warmup 3 --no-cpu-baseline` (per-launch times are cold-cache and "
                 "serialised: compare shares, not absolutes).  First 200 launches of the process.\n\n"
                 "| kernel | grid | launches | total us | share | avg us |\n|---|---|---|---|---|---|\n")
        for (k, g), v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            fh.write(f"| `{k[:120]}` | {g} | {len(v)} | {sum(v)/1e3:.1f} | {100*sum(v)/tot:.1f}% | {sum(v)/len(v)/1e3:.1f} |\n")
    os.system(f"cp {ll} {os.path.join(out_dir, tag + '_launches.csv')}")
for f in ("pipe_microbench.txt", "bench.json", "bench_ref.json", "sweep.jsonl", "cpu.txt", "gpu.txt"):
    p = os.path.join(src, f)
    if os.path.exists(p):
        os.system(f"cp {p} {os.path.join(out_dir, tag + '_' + f)}")
print("ok")
