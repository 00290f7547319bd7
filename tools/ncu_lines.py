#!/usr/bin/env python3
"""Per source line (CUDA-C view) executed warp-instructions and stall samples of one kernel of an ncu report:
    python tools/ncu_lines.py report.ncu-rep <kernel regex> [launch-skip] [top]"""
import csv
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + rx,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
out = {}
fname = ""
hdr = None
cur = None
for r in rows:
    if not r:
        continue
    if r[0] in ("File Path", "File Name"):
        fname = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].isdigit():
        cur = (fname, int(r[0]), r[1].strip())
    elif hdr and r[0] == "" and cur and len(r) > 4:
        d = dict(zip(hdr[2:], r[2:]))
        num = lambda v: int(v) if v and v.isdigit() else 0
        ex = num(d.get("Instructions Executed"))
        st = num(d.get("Warp Stall Sampling (All Samples)"))
        o = out.setdefault(cur, [0, 0, {}])
        o[0] += ex
        o[1] += st
        for k, v in d.items():
            if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v) > 0:
                o[2][k[6:]] = o[2].get(k[6:], 0) + int(v)
out = [(v[0], v[1], k[0], k[1], k[2], v[2]) for k, v in out.items() if v[0] or v[1]]
tex, tst = sum(o[0] for o in out), sum(o[1] for o in out)
print(f"total warp-instructions {tex}, stall samples {tst}")
for ex, st, f, ln, src, rs in sorted(out, key=lambda o: -o[1])[:top]:
    rtxt = " ".join(f"{k}:{v}" for k, v in sorted(rs.items(), key=lambda kv: -kv[1])[:3])
    print(f"{100.0 * st / max(tst, 1):5.1f}% stalls {100.0 * ex / max(tex, 1):5.1f}% instr  {f}:{ln:<5} {src[:70]:70s} {rtxt}")
