#!/usr/bin/env python3
"""Timing of the other BASELINE.json configurations (3: PDOS, 4: JDOS + optics, 5: core-loss) on one B200.
These are parity-test shapes, not bench.py lines; this script records what the same engines do on them.
Inputs are generated in HBM (bands by od_synth_bands, weights / matrix elements by torch on the device and
passed as OD_MEM_DEVICE pointers).  usage: python tools/bench_configs.py [cfg3 cfg4 cfg5] [--scale f]"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import optados_b200 as ob  # noqa: E402


def timed(ctx, fn, reps=2):
    fn()
    ctx.synchronize()
    ctx.profile_reset()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    ctx.synchronize()
    dt = (time.perf_counter() - t0) / reps
    p = ctx.profile_get()
    return dt, {k: (v / reps if isinstance(v, float) else v) for k, v in p.items()}, out


def cfg3(scale):
    # 500k k-points x 200 bands x 64 projectors, ns = 1, adaptive, 20k bins
    grid = (100, 100, 50)
    nk, nb, ns, norb, nbins = int(grid[0] * grid[1] * grid[2] * scale), 200, 1, 64, 20000   # a k-slab of the full mesh
    ctx = ob.Context(0)
    ctx.synth_bands(grid, 0, nk, ns, nb, seed=3)
    g = torch.Generator(device="cuda").manual_seed(3)
    mw = torch.rand((nk * nb * ns, norb), dtype=torch.float64, device="cuda", generator=g)
    mw /= mw.sum(dim=1, keepdim=True)      # sum over projectors = 1
    mn, mx = ctx.band_minmax()
    emin, delta = mn - 5.0, (mx - mn + 10.0) / (nbins - 1)
    dt, prof, out = timed(ctx, lambda: ctx.calculate_dos_ptr("a", emin, delta, nbins, mw.data_ptr(), norb, nb, nk))
    dos, intdos, wdos = out
    err = abs(wdos.sum(axis=2) - dos).max() / abs(dos).max()      # projector weights sum to one -> sum of PDOS = DOS
    contrib = ctx.count_contributions("a", emin, delta, nbins)
    return {"config": "cfg3 PDOS", "k_points": nk, "bands": nb, "projectors": norb, "bins": nbins, "s_per_pass": dt,
            "contributions": contrib, "weighted_contributions_per_s": contrib * norb / dt, "sum_rule_rel_err": float(err),
            "profile": prof, "mw_GB": mw.numel() * 8 / 1e9}


def cfg4(scale):
    # 200k k-points x 128 bands (64 occupied), all occ -> unocc pairs, polycrystalline and tensor eps2; the OME array
    # of the full config is 157 GB per spin, i.e. it only exists sharded: this is the 1/8 shard of an 8-GPU run
    grid = (50, 50, 80)
    nk, nb, ns = int(grid[0] * grid[1] * grid[2] * scale / 8), 128, 1   # the k-shard one of 8 GPUs holds
    ctx = ob.Context(0)
    ctx.synth_bands(grid, 0, nk, ns, nb, seed=4)
    g = torch.Generator(device="cuda").manual_seed(4)
    ome = torch.randn((nk * ns * 3 * nb * nb, 2), dtype=torch.float64, device="cuda", generator=g) * 0.1 * 0.52917720859 * 27.21138342902473
    p = ctx.dos_params("adaptive", dos_nbins=20000, num_electrons=(128.0,), efermi="insulator")
    ef = ctx.dos_utils_set_efermi(p)
    jmax, spacing = 30.0, 0.01
    nbins = int(abs(-(-jmax // spacing)))
    delta = nbins * spacing / (nbins - 1)
    res = {"config": "cfg4 JDOS + optics (1/8 shard)", "k_points": nk, "bands": nb, "bins": nbins, "efermi": ef, "ome_GB": ome.numel() * 8 / 1e9}
    for geom in ("polycrys", "tensor"):
        dt, prof, out = timed(ctx, lambda: ctx.calculate_jdos_optics_ptr("a", delta, nbins, ef, ome.data_ptr(), geom), reps=1)
        res[geom] = {"s_per_pass": dt, "pairs_x_bins_in_window": prof["contributions"], "lane_steps": prof["lane_steps"],
                     "contributions_per_s": prof["contributions"] / dt, "profile": prof}
    return res


def cfg5(scale):
    # 300k k-points x 300 bands x 16 core orbitals, adaptive, absorption, polycrystalline
    grid = (60, 50, 100)
    nk, nb, ns, norb, nbins = int(grid[0] * grid[1] * grid[2] * scale), 300, 1, 16, 20000
    ctx = ob.Context(0)
    ctx.synth_bands(grid, 0, nk, ns, nb, seed=5)
    g = torch.Generator(device="cuda").manual_seed(5)
    elnes = torch.randn((nk * ns * 3 * nb * norb, 2), dtype=torch.float64, device="cuda", generator=g) * 0.01
    mw = torch.empty((nk * ns * nb * norb,), dtype=torch.float64, device="cuda")
    p = ctx.dos_params("adaptive", dos_nbins=nbins, num_electrons=(300.0,), efermi="insulator")
    ef = ctx.dos_utils_set_efermi(p)
    t0 = time.perf_counter()
    ctx.core_prepare_matrix_elements_ptr(elnes.data_ptr(), norb, "absorption", ef, mw.data_ptr())
    ctx.synchronize()
    t_w = time.perf_counter() - t0
    mn, mx = ctx.band_minmax()
    emin, delta = mn - 5.0, (mx - mn + 10.0) / (nbins - 1)
    dt, prof, out = timed(ctx, lambda: ctx.calculate_dos_ptr("a", emin, delta, nbins, mw.data_ptr(), norb, nb, nk))
    contrib = ctx.count_contributions("a", emin, delta, nbins)
    return {"config": "cfg5 core-loss", "k_points": nk, "bands": nb, "core_orbitals": norb, "bins": nbins, "efermi": ef,
            "core_prepare_matrix_elements_s": t_w, "elnes_GB": elnes.numel() * 8 / 1e9,
            "weight_map_GBps": (elnes.numel() * 8 + mw.numel() * 8) / t_w / 1e9, "s_per_pass": dt, "contributions": contrib,
            "weighted_contributions_per_s": contrib * norb / dt, "profile": prof}


def post(scale):
    """the reference's root-only O(B^2) passes on the grids of configs 4 and 5 (one GPU)"""
    import numpy as np
    ctx = ob.Context(0)
    out = {"config": "post-processing"}
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


if __name__ == "__main__":
    scale = 1.0
    which = [a for a in sys.argv[1:] if a.startswith("cfg") or a == "post"] or ["cfg3", "cfg4", "cfg5", "post"]
    if "--scale" in sys.argv:
        scale = float(sys.argv[sys.argv.index("--scale") + 1])
    for name in which:
        r = {"cfg3": cfg3, "cfg4": cfg4, "cfg5": cfg5, "post": post}[name](scale)
        print(json.dumps(r), flush=True)
        torch.cuda.empty_cache()
