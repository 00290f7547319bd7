#!/usr/bin/env python3
"""The hot loops of a kernel as SASS, from the source page of an ncu report (compiled with -lineinfo, captured with
--import-source on): contiguous instruction runs with the same execution count, hottest first, with their instruction mix.
    python tools/sass_loops.py report.ncu-rep <kernel regex> [launch-skip] [number of runs] > profiles/rNN_sass_<kernel>.txt"""
import csv
import re
import subprocess
import sys
from collections import Counter

rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
nruns = int(sys.argv[4]) if len(sys.argv) > 4 else 3
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx, "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
isrc, iall, iex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
ins = [(int(r[iex] or 0), int(r[iall] or 0), r[isrc]) for r in rows[hi + 1:] if len(r) > max(iall, iex, isrc) and r[iex].isdigit()]
total_ex = sum(i[0] for i in ins)
total_smp = sum(i[1] for i in ins)
runs, start = [], 0
for n in range(1, len(ins) + 1):
    if n == len(ins) or ins[n][0] != ins[start][0]:
        if n - start >= 8 and ins[start][0] > 0:
            runs.append((ins[start][0] * (n - start), start, n))
        start = n


def opcode(src):
    t = src.split()
    return (t[1] if t[0].startswith("@") else t[0]).split(".")[0]


print(f"# {rows[0][1] if len(rows[0]) > 1 else rx}")
print(f"# {rep}: {len(ins)} SASS instructions, {total_ex} warp-instructions executed, {total_smp} stall samples")
for w, a, b in sorted(runs, reverse=True)[:nruns]:
    mix = Counter(opcode(i[2]) for i in ins[a:b])
    fp64 = sum(v for k, v in mix.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
    smp = sum(i[1] for i in ins[a:b])
    print(f"\n## instructions #{a}..#{b - 1}: {b - a} instructions x {ins[a][0]} executions = {100.0 * w / total_ex:.1f} % of the executed instructions, "
          f"{100.0 * smp / max(1, total_smp):.1f} % of the stall samples")
    print("## mix: " + ", ".join(f"{k} {v}" for k, v in mix.most_common()) + f"   [FP64 pipe (DFMA/DMUL/DADD/DSETP): {fp64}; DMMA: {mix.get('DMMA', 0)}; MUFU: {mix.get('MUFU', 0)}]")
    for n in range(a, b):
        print(f"{ins[n][1]:7d}  {ins[n][2]}")
