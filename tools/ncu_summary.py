#!/usr/bin/env python3
"""Turn gpurun_out/*.ncu-rep (+ the launch-list csv) into the small tracked summaries under profiles/.
usage: tools/ncu_summary.py <report.ncu-rep> <launches.csv> <out_prefix> [units of the captured slab]"""
import csv
import subprocess
import sys
from collections import OrderedDict

METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
           "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def main():
    rep, launches, out = sys.argv[1:4]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(out + "_kernels.csv", "w") as f:
        w = csv.writer(f)
        cols = [hdr.index("Kernel Name")] + [hdr.index(m) for m in METRICS if m in hdr]
        w.writerow([hdr[c] for c in cols])
        w.writerow([units[c] for c in cols])
        for r in data:
            w.writerow([r[c].split("(")[0] if c == cols[0] else r[c] for c in cols])
    # per-kernel figures of the captured slab for bench.py (roofline.traffic, roofline.kernels_ncu): kernels of the same
    # name (template arguments kept) are summed over the capture
    if len(sys.argv) > 4:
        import json
        import re
        units_slab = float(sys.argv[4])
        g = lambda r, m: float(r[hdr.index(m)].replace(",", "")) if m in hdr and r[hdr.index(m)] else 0.0
        ub = lambda m: {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(units[hdr.index(m)], 1.0)
        ut = {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(units[hdr.index("gpu__time_duration.sum")], 1.0)
        ks = OrderedDict()
        for r in data:
            name = re.sub(r"^void |<unnamed>::|\(.*$", "", r[hdr.index("Kernel Name")])
            k = ks.setdefault(name, {"launches": 0, "ms": 0.0, "dram_bytes": 0.0, "fp64_pipe_pct": 0.0, "units": units_slab})
            ms = g(r, "gpu__time_duration.sum") * ut
            k["launches"] += 1
            k["fp64_pipe_pct"] += g(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active") * ms   # time-weighted
            k["ms"] += ms
            k["dram_bytes"] += g(r, "dram__bytes_read.sum") * ub("dram__bytes_read.sum") + g(r, "dram__bytes_write.sum") * ub("dram__bytes_write.sum")
        for k in ks.values():
            k["fp64_pipe_pct"] = k["fp64_pipe_pct"] / k["ms"] if k["ms"] > 0 else 0.0
        json.dump({"source": rep, "units_of_the_captured_slab": units_slab,
                   "note": "ncu --set full --clock-control none of one mid-mesh k-point slab (tools/ncu_slab.py), both schemes once; ms and "
                           "dram_bytes are sums over the launches of a kernel in that capture (cold-cache, serialised)", "kernels": ks},
                  open(out + "_kernels.json", "w"), indent=1)
    rows = list(csv.reader(open(launches)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot = OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0]
        tot.setdefault(name, [0, 0.0])
        tot[name][0] += 1
        tot[name][1] += float(r[vi].replace(",", ""))
    s = sum(v[1] for v in tot.values())
    with open(out + "_launches.md", "w") as f:
        f.write("| kernel | launches | ms / launch | share of all kernel time |\n|---|---|---|---|\n")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {v[0]} | {v[1] / 1e6 / v[0]:.3f} | {100 * v[1] / s:.1f} % |\n")
    print("wrote", out + "_kernels.csv", out + "_launches.md")


if __name__ == "__main__":
    main()
