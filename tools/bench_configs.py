#!/usr/bin/env python3
"""The other BASELINE.json configurations (3: PDOS, 4: JDOS + optics, 5: core-loss) on B200: timing, roofline
fractions and a k-slab parity figure against the oracle.  Used by bench.py for the `configs` block of its JSON line
(config 2 stays the headline), and stand-alone:

    python tools/bench_configs.py [cfg3 cfg4 cfg5 post] [--scale f]

Inputs are generated in HBM (bands by od_synth_bands, weights / matrix elements by torch on the device, passed as
OD_MEM_DEVICE pointers).  With `world` > 1 every rank takes a contiguous block of k-points (algor_dist_array) and
the library merges the partial spectra (merge=True: the NCCL allreduce that replaces comms_reduce).

Algorithmic work: one contribution = one in-window (state or pair, bin); FP64 flops per
contribution: 76 (Gaussian DOS + erf integrated DOS) + 2 per weight column for PDOS / core-loss; 38 + 2 per geometry
component for JDOS / optics; HBM bytes per state: 32 + 8 per weight column, per pair 48 (the three complex matrix
elements)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import optados_b200 as ob  # noqa: E402

BOHR2ANG, H2EV = 0.52917720859, 27.21138342902473


def _block(nk_total, rank, world):
    per = ob.dist_array(nk_total, world)
    return sum(per[:rank]), per[rank]


def _timed(ctx, fn, reps, sync_all=None):
    fn()
    ctx.synchronize()
    if sync_all:
        sync_all()
    ctx.profile_reset()
    ctx.timer_start()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    ms = ctx.timer_stop() / reps
    ctx.synchronize()
    wall = (time.perf_counter() - t0) / reps
    p = ctx.profile_get()
    return ms, wall, {k: (v / reps if isinstance(v, float) else v) for k, v in p.items()}, out


def _roof(flops, nbytes, ms, fp64_peak, hbm_peak):
    tf = flops / (ms * 1e-3) / 1e12
    gbs = nbytes / (ms * 1e-3) / 1e9
    f_fp, f_hbm = tf / fp64_peak, gbs / hbm_peak
    return {"fp64_tflops": tf, "fp64_frac": f_fp, "hbm_gbs": gbs, "hbm_frac": f_hbm, "bound": "fp64" if f_fp >= f_hbm else "hbm",
            "frac": max(f_fp, f_hbm)}


def _rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300))


class Env:
    """what bench.py hands over: device index, rank layout, reductions over ranks, peaks"""

    def __init__(self, local=0, rank=0, world=1, make_ctx=None, max_over_ranks=None, sum_over_ranks=None, barrier=None,
                 fp64_peak=36.7, hbm_peak=6545.0, reps=2, scale=1.0, parity=True):
        self.local, self.rank, self.world = local, rank, world
        self.make_ctx = make_ctx or (lambda: ob.Context(local))
        self.max_over_ranks = max_over_ranks or (lambda x: x)
        self.sum_over_ranks = sum_over_ranks or (lambda x: x)
        self.barrier = barrier or (lambda: None)
        self.fp64_peak, self.hbm_peak, self.reps, self.scale, self.parity = fp64_peak, hbm_peak, reps, scale, parity


def cfg3(env):
    """PDOS: 500k k-points x 200 bands x 64 projectors, one spin, adaptive, 20 000 bins"""
    import torch
    grid, nb, ns, norb, nbins = (100, 100, 50), 200, 1, 64, 20000
    nk_total = int(grid[0] * grid[1] * grid[2] * env.scale)
    k0, nk = _block(nk_total, env.rank, env.world)
    ctx = env.make_ctx()
    ctx.synth_bands(grid, k0, nk, ns, nb, seed=3)
    g = torch.Generator(device=f"cuda:{env.local}").manual_seed(3 + env.rank)
    mw = torch.rand((nk * nb * ns, norb), dtype=torch.float64, device=f"cuda:{env.local}", generator=g)
    mw /= mw.sum(dim=1, keepdim=True)      # sum over projectors = 1
    mn, mx = ctx.band_minmax()             # (MIN / MAX-reduced over the ranks)
    emin, delta = mn - 5.0, (mx - mn + 10.0) / (nbins - 1)
    merge = env.world > 1
    ms, wall, prof, out = _timed(ctx, lambda: ctx.calculate_dos_ptr("a", emin, delta, nbins, mw.data_ptr(), norb, nb, nk, merge=merge),
                                 env.reps, env.barrier)
    dos, intdos, wdos = out
    ms = env.max_over_ranks(ms)
    contrib = env.sum_over_ranks(ctx.count_contributions("a", emin, delta, nbins))
    states = float(nk_total) * nb * ns
    res = {"workload": "cfg3: PDOS, adaptive, 64 projectors (matrix_weights contraction on the FP64 tensor cores)",
           "k_points_total": nk_total, "bands": nb, "spins": ns, "projectors": norb, "energy_bins": nbins, "ms": ms,
           "contributions": contrib, "contributions_per_s": contrib / (ms * 1e-3), "weighted_contributions_per_s": contrib * norb / (ms * 1e-3),
           "sum_rule_rel_err": _rel(wdos.sum(axis=2), dos),       # projector weights sum to one -> sum of PDOS = DOS
           "matrix_weights_GB_per_gpu": mw.numel() * 8 / 1e9,
           "roofline": _roof(contrib * (76.0 + 2.0 * norb), states * (32.0 + 8.0 * norb), ms, env.fp64_peak * env.world, env.hbm_peak * env.world),
           "device_ms": {"prepare": prof["prepare_ms"], "accumulate": prof["accumulate_ms"]}}
    del mw
    ctx.close()
    if env.parity and env.rank == 0:
        res["parity_vs_oracle"] = _parity_weighted(env, grid, nb, ns, norb, nbins, seed=3, nk=16, k_first=251300)
    return res


def cfg5(env):
    """core-loss: 300k k-points x 300 bands x 16 core orbitals, adaptive, absorption, polycrystalline"""
    import torch
    grid, nb, ns, norb, nbins = (60, 50, 100), 300, 1, 16, 20000
    nk_total = int(grid[0] * grid[1] * grid[2] * env.scale)
    k0, nk = _block(nk_total, env.rank, env.world)
    ctx = env.make_ctx()
    ctx.synth_bands(grid, k0, nk, ns, nb, seed=5)
    dev = f"cuda:{env.local}"
    g = torch.Generator(device=dev).manual_seed(5 + env.rank)
    elnes = torch.randn((nk * ns * 3 * nb * norb, 2), dtype=torch.float64, device=dev, generator=g) * 0.01
    mw = torch.empty((nk * ns * nb * norb,), dtype=torch.float64, device=dev)
    p = ctx.dos_params("adaptive", dos_nbins=nbins, num_electrons=(300.0,), efermi="insulator")
    ef = ctx.dos_utils_set_efermi(p)
    ctx.core_prepare_matrix_elements_ptr(elnes.data_ptr(), norb, "absorption", ef, mw.data_ptr())
    ctx.synchronize()
    env.barrier()
    ctx.timer_start()
    ctx.core_prepare_matrix_elements_ptr(elnes.data_ptr(), norb, "absorption", ef, mw.data_ptr())
    ms_w = env.max_over_ranks(ctx.timer_stop())
    mn, mx = ctx.band_minmax()
    emin, delta = mn - 5.0, (mx - mn + 10.0) / (nbins - 1)
    merge = env.world > 1
    ms, wall, prof, out = _timed(ctx, lambda: ctx.calculate_dos_ptr("a", emin, delta, nbins, mw.data_ptr(), norb, nb, nk, merge=merge),
                                 env.reps, env.barrier)
    ms = env.max_over_ranks(ms)
    contrib = env.sum_over_ranks(ctx.count_contributions("a", emin, delta, nbins))
    states = float(nk_total) * nb * ns
    # absorption: the weights of the occupied bands (150 of 300 here) are zero without reading their matrix elements
    # (core.f90:126-133), so the map reads 3 complex numbers for half of the (orbital, band) pairs and writes one weight each
    map_bytes = states * norb * (0.5 * 48.0 + 8.0)
    res = {"workload": "cfg5: core-loss ELNES, adaptive, 16 core orbitals (core_prepare_matrix_elements + weighted DOS)",
           "k_points_total": nk_total, "bands": nb, "spins": ns, "core_orbitals": norb, "energy_bins": nbins, "efermi": ef,
           "ms": ms, "contributions": contrib, "contributions_per_s": contrib / (ms * 1e-3),
           "weighted_contributions_per_s": contrib * norb / (ms * 1e-3), "elnes_GB_per_gpu": elnes.numel() * 8 / 1e9,
           "roofline": _roof(contrib * (76.0 + 2.0 * norb), states * (32.0 + 8.0 * norb), ms, env.fp64_peak * env.world, env.hbm_peak * env.world),
           "core_prepare_matrix_elements": {"ms": ms_w, "roofline": _roof(states * norb * 20.0, map_bytes, ms_w, env.fp64_peak * env.world,
                                                                           env.hbm_peak * env.world)},
           "device_ms": {"prepare": prof["prepare_ms"], "accumulate": prof["accumulate_ms"]}}
    del elnes, mw
    ctx.close()
    if env.parity and env.rank == 0:
        res["parity_vs_oracle"] = _parity_core(env, grid, nb, ns, norb, nbins, seed=5, nk=16, k_first=151200)
    return res


def cfg4(env, k_per_gpu=25000):
    """JDOS + optics eps_2: 200k k-points x 128 bands (64 occupied), all occupied -> unoccupied pairs, make_weights fused.
    optical_mat is 157 GB per spin for the whole mesh, i.e. it only exists sharded: every GPU holds 25 000 k-points
    (19.7 GB), so that the full 200k-point problem is what N = 8 runs; fewer GPUs run the first N/8 of it."""
    import torch
    grid, nb, ns = (50, 50, 80), 128, 1
    nk = int(k_per_gpu * env.scale)
    nk_total = nk * env.world
    k0 = nk * env.rank
    ctx = env.make_ctx()
    ctx.synth_bands(grid, k0, nk, ns, nb, seed=4)
    dev = f"cuda:{env.local}"
    g = torch.Generator(device=dev).manual_seed(4 + env.rank)
    ome = torch.randn((nk * ns * 3 * nb * nb, 2), dtype=torch.float64, device=dev, generator=g) * (0.1 * BOHR2ANG * H2EV)
    p = ctx.dos_params("adaptive", dos_nbins=20000, num_electrons=(128.0,), efermi="insulator")
    ef = ctx.dos_utils_set_efermi(p)     # vbm / cbm are MIN / MAX-reduced over the ranks (dos_utils.f90:353-356)
    jmax, spacing = 30.0, 0.01
    nbins = int(abs(-(-jmax // spacing)))
    delta = nbins * spacing / (nbins - 1)
    merge = env.world > 1
    res = {"workload": "cfg4: JDOS + eps_2 weights (make_weights fused), adaptive, all occupied -> unoccupied pairs",
           "k_points_total": nk_total, "k_points_per_gpu": nk, "bands": nb, "occupied": 64, "energy_bins": nbins, "efermi": ef,
           "optical_mat_GB_per_gpu": ome.numel() * 8 / 1e9}
    pairs = float(nk_total) * 64 * 64 * ns
    for geom, ng in (("polycrys", 1), ("tensor", 6)):
        ms, wall, prof, out = _timed(ctx, lambda: ctx.calculate_jdos_optics_ptr("a", delta, nbins, ef, ome.data_ptr(), geom, merge=merge),
                                     env.reps, env.barrier)
        ms = env.max_over_ranks(ms)
        evals = env.sum_over_ranks(prof["contributions"])       # in-window (pair, bin) evaluations of the engine, ALL channel passes
        lane = env.sum_over_ranks(prof["lane_steps"])
        npass = (1 + ng + 3) // 4 if ng >= 2 else 1 + ng // 2   # channels per pass: 4 from three channels on (tensor: 2 passes), else 2
        contrib = evals / npass                                 # the metric's unit: one in-window (pair, bin) contribution, counted once
        res[geom] = {"ms": ms, "contributions": contrib, "contributions_per_s": contrib / (ms * 1e-3), "lane_efficiency": evals / max(lane, 1.0),
                     "channel_passes": npass, "engine_evaluations": evals,
                     "roofline": _roof(contrib * (38.0 + 2.0 * ng), pairs * 48.0, ms, env.fp64_peak * env.world, env.hbm_peak * env.world),
                     "device_ms": {"prepare": prof["prepare_ms"], "accumulate": prof["accumulate_ms"]}}
    del ome
    ctx.close()
    if env.parity and env.rank == 0:
        res["parity_vs_oracle"] = _parity_optics(env, grid, nb, ns, nbins, delta, seed=4, nk=8, k_first=99900)
    return res


# ---------------------------------------------------------------- k-slab parity against the oracle (rank 0, one GPU)
def _slab(env, grid, nb, ns, seed, nk, k_first):
    from oracle import odo
    ctx = ob.Context(env.local)
    ctx.synth_bands(grid, k_first, nk, ns, nb, seed=seed)
    be, bg, kw = ctx.get_bands(), ctx.get_gradients(), ctx.get_kweights()
    s = odo.Sys(be, kw, np.eye(3) * (2 * np.pi / 5.43), grid, bg, electrons_per_state=2.0 if ns == 1 else 1.0)
    return ctx, s, odo


def _parity_weighted(env, grid, nb, ns, norb, nbins, seed, nk, k_first):
    ctx, s, odo = _slab(env, grid, nb, ns, seed, nk, k_first)
    rng = np.random.default_rng(seed)
    mw = rng.random((norb, nb, nk, ns))
    mw = np.asfortranarray(mw / mw.sum(axis=0, keepdims=True))
    emin, emax = s.band_energy.min() - 5.0, s.band_energy.max() + 5.0
    delta = (emax - emin) / (nbins - 1)
    E = emin + delta * np.arange(nbins)
    ref = odo.calculate_dos("a", s, odo.make_broadening(), E, delta, mw, nthreads=os.cpu_count() or 1)
    dos, intdos, wdos = ctx.calculate_dos("a", emin, delta, nbins, mw=mw)
    ctx.close()
    return {"k_points": nk, "max_rel_dos": _rel(dos, ref[0]), "max_rel_intdos": _rel(intdos, ref[1]), "max_rel_weighted_dos": _rel(wdos, ref[2])}


def _parity_core(env, grid, nb, ns, norb, nbins, seed, nk, k_first):
    ctx, s, odo = _slab(env, grid, nb, ns, seed, nk, k_first)
    rng = np.random.default_rng(seed)
    em = np.asfortranarray(rng.normal(0, 0.01, (norb, nb, 3, nk, ns)) + 1j * rng.normal(0, 0.01, (norb, nb, 3, nk, ns)))
    s.num_electrons = np.array([300.0])
    ef = odo.efermi_insulator(s)
    mw_ref = odo.core_prepare_matrix_elements(s, em, "absorption", ef)
    mw = ctx.core_prepare_matrix_elements(em, "absorption", ef)
    emin, emax = s.band_energy.min() - 5.0, s.band_energy.max() + 5.0
    delta = (emax - emin) / (nbins - 1)
    E = emin + delta * np.arange(nbins)
    ref = odo.calculate_dos("a", s, odo.make_broadening(), E, delta, mw_ref, nthreads=os.cpu_count() or 1)
    dos, intdos, wdos = ctx.calculate_dos("a", emin, delta, nbins, mw=mw)
    ctx.close()
    return {"k_points": nk, "max_rel_weights": _rel(mw, mw_ref), "max_rel_dos": _rel(dos, ref[0]), "max_rel_intdos": _rel(intdos, ref[1]),
            "max_rel_weighted_dos": _rel(wdos, ref[2])}


def _parity_optics(env, grid, nb, ns, nbins, delta, seed, nk, k_first):
    ctx, s, odo = _slab(env, grid, nb, ns, seed, nk, k_first)
    rng = np.random.default_rng(seed)
    om = np.asfortranarray((rng.normal(0, 0.1, (nb, nb, 3, nk, ns)) + 1j * rng.normal(0, 0.1, (nb, nb, 3, nk, ns))) * (BOHR2ANG * H2EV))
    s.num_electrons = np.array([128.0])
    ef = odo.efermi_insulator(s)
    E = delta * np.arange(nbins)
    out = {"k_points": nk}
    for geom in ("polycrys", "tensor"):
        mw = odo.make_weights(s, om, geom, ef)
        rj, rw = odo.calculate_jdos("a", s, odo.make_broadening(), E, delta, ef, mw=mw, nthreads=os.cpu_count() or 1)
        jd, wj = ctx.calculate_jdos_optics("a", delta, nbins, ef, om, geom)
        out[geom] = {"max_rel_jdos": _rel(jd, rj), "max_rel_weighted_jdos": _rel(wj, rw)}
    ctx.close()
    return out


def post(env):
    """the reference's root-only O(B^2) passes on the grids of configs 4 and 5 (one GPU)"""
    ctx = ob.Context(env.local)
    out = {"workload": "post-processing"}
    rng = np.random.default_rng(7)
    for name, nbins, ncols in (("kk_eps1_cfg4_tensor", 3000, 6), ("kk_eps1_20k", 20000, 6), ("lai_cfg5", 20000, 16), ("loss_fn_20k", 20000, 1)):
        E = 0.01 * np.arange(nbins)
        x = np.asfortranarray(rng.random((nbins, ncols)))
        if name.startswith("kk"):
            fn = lambda: ctx.calc_epsilon_1(E, x)
            pairs = nbins * nbins * ncols
        elif name.startswith("lai"):
            fn = lambda: ctx.core_lai_broadening(E, x, 50.0, gaussian_width=0.8, lorentzian_width=0.2, lorentzian_scale=0.1)
            pairs = 2 * nbins * nbins * ncols
        else:
            fn = lambda: ctx.calc_loss_fn(E, 1.0 + x[:, 0], x[:, 0], lossfn_gaussian=1.0)
            pairs = nbins * nbins
        fn()
        t0 = time.perf_counter()
        for _ in range(3):
            fn()
        dt = (time.perf_counter() - t0) / 3
        out[name] = {"bins": nbins, "columns": ncols, "ms": dt * 1e3, "pair_terms_per_s": pairs / dt}
    ctx.close()
    return out


RUNNERS = {"cfg3": cfg3, "cfg4": cfg4, "cfg5": cfg5, "post": post}

if __name__ == "__main__":
    import torch
    scale = 1.0
    which = [a for a in sys.argv[1:] if a in RUNNERS] or ["cfg3", "cfg4", "cfg5", "post"]
    if "--scale" in sys.argv:
        scale = float(sys.argv[sys.argv.index("--scale") + 1])
    env = Env(scale=scale, parity="--no-parity" not in sys.argv)
    c = ob.Context(0)
    env.fp64_peak = c.measure_fp64_peak(0.4)
    c.close()
    for name in which:
        print(json.dumps(RUNNERS[name](env)), flush=True)
        torch.cuda.empty_cache()
