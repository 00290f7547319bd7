#!/usr/bin/env python3
"""One line per captured launch: time, DRAM bytes, instructions, issue-slot and DRAM utilisation.  usage: ncu_brief2.py report.ncu-rep"""
import csv, subprocess, sys
M = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
     'smsp__issue_active.avg.pct_of_peak_sustained_active', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
     'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active']
raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv', '--metrics', ','.join(M)], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
ix = [h.index(c) for c in ['Kernel Name'] + M]
print('kernel | ' + ' | '.join(f'{m.split(".")[0][-14:]} [{rows[1][h.index(m)]}]' for m in M))
for r in rows[2:]:
    print(' | '.join(r[i][:48] if n == 0 else f'{float(r[i]):.3f}' for n, i in enumerate(ix)))
