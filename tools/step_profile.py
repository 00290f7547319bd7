#!/usr/bin/env python3
"""Per-scheme device times of the config-2 step (prepare / accumulate, lane efficiency) without a profiler:
    python tools/step_profile.py [nk] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import optados_b200 as ob

nk = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = ob.Context(0)
ctx.synth_bands((100, 100, 100), 0, nk, 2, 256, seed=2)
mn, mx = ctx.band_minmax()
emin, nb = mn - 5.0, 20000
delta = (mx + 5.0 - emin) / (nb - 1)
for t in ("a", "l"):
    ctx.calculate_dos(t, emin, delta, nb)
    ctx.profile_reset()
    ctx.timer_start()
    for _ in range(reps):
        ctx.calculate_dos(t, emin, delta, nb)
    ms = ctx.timer_stop() / reps
    p = ctx.profile_get()
    c = ctx.count_contributions(t, emin, delta, nb)
    print(f"{t}: total {ms:.2f} ms  prepare {p['prepare_ms'] / reps:.2f}  accumulate {p['accumulate_ms'] / reps:.2f}  "
          f"lane efficiency {p['contributions'] / p['lane_steps']:.3f}  in-window (engine) {p['contributions'] / reps:.4g}  "
          f"in-window (metric) {c:.4g}  accumulate rate {p['contributions'] / reps / (p['accumulate_ms'] / reps * 1e-3):.4g} /s")
ctx.close()

# fixed cost per call: a 2000-k-point problem through the driver (energy scale, both schemes, merge, Fermi analysis)
import time
ctx = ob.Context(0)
ctx.synth_bands((100, 100, 100), 0, 2000, 2, 256, seed=2)
p = ctx.dos_params(["adaptive", "linear"], dos_nbins=20000, num_electrons=(128.0, 128.0))
for _ in range(3):
    p.dos_nbins = 20000
    ctx.dos_utils_calculate(p)
ctx.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    p.dos_nbins = 20000
    ctx.dos_utils_calculate(p)
ctx.synchronize()
print(f"2000 k-points, adaptive + linear through od_dos_utils_calculate: {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms per step (wall)")
ctx.close()
