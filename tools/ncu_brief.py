#!/usr/bin/env python3
"""Brief per-kernel summary of an ncu report: python tools/ncu_brief.py report.ncu-rep"""
import csv
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_lsu.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d.get("Kernel Name", "?")[:90])
    for k in keys:
        if k in d:
            print(f"   {k} = {d[k]} {rows[1][hdr.index(k)]}")
    st = sorted(((float(v), k) for k, v in d.items() if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and v), reverse=True)
    print("   stalls/issue:", ", ".join(f"{k.split('issue_stalled_')[1].split('_per_issue')[0]}={v:.2f}" for v, k in st[:7]))
