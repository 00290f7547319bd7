#!/usr/bin/env python3
"""Hottest SASS instructions (warp-stall samples) of one kernel of an ncu report:
    python tools/ncu_hot.py report.ncu-rep <kernel regex> [launch-skip] [top]"""
import csv
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx, "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
isrc, iall, iex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
data = []
for n, r in enumerate(rows[hi + 1:]):
    if len(r) > iall and r[iall].isdigit():
        data.append((int(r[iall]), int(r[iex] or 0), n, r[isrc]))
tot = sum(d[0] for d in data)
print(rows[0][1][:100] if len(rows[0]) > 1 else "", "total samples", tot, "instructions", len(data))
for s, ex, n, src in sorted(data, reverse=True)[:top]:
    print(f"{100.0 * s / tot:6.2f}%  exec {ex:>9}  #{n:<5} {src[:120]}")
