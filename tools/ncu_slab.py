#!/usr/bin/env python3
"""One mid-mesh k-point slab of config 2 (adaptive, then linear), twice: the workload behind the per-kernel ncu captures
    ncu --set full --clock-control none --import-source on -k regex:<kernels> -s <skip> -c <count> -o out python tools/ncu_slab.py [nk]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import optados_b200 as ob

nk = int(sys.argv[1]) if len(sys.argv) > 1 else 62500      # x 256 bands x 2 spins = 32M units
ctx = ob.Context(0)
ctx.synth_bands((100, 100, 100), 240000, nk, 2, 256, seed=2)
emin, nb = -14.9, 20000
delta = 44.5 / (nb - 1)
for rep in range(2):
    for t in ("a", "l"):
        ctx.calculate_dos(t, emin, delta, nb)
ctx.synchronize()
ctx.close()
