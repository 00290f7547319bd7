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
