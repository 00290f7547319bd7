#!/usr/bin/env python3
"""bench.py -- throughput of the Brillouin-zone spectral integration hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): synthetic 1M k-points x 256 bands x 2 spins, adaptive + linear
DOS (with integrated DOS) on a 20 000-bin energy grid.  One "step" = one dos_utils_calculate pass
(energy scale, adaptive calculate_dos + merge, linear calculate_dos + merge, Fermi analysis) over
the rank's k-point shard.  Multi-GPU: one process per GPU, k-points sharded in contiguous blocks
(algor_dist_array), partial spectra combined with ONE NCCL allreduce per broadening inside the
library (replaces comms_reduce); STRONG scaling by default -- the named 1M-k-point problem is split
over the N GPUs (`--scaling weak` instead gives every GPU its own 1M k-points of a 100x100x(100 N) mesh).

Metric: in-window broadened (k,band,E) contributions per second, whole
job.  `value` is measured with the inputs resident in HBM; `e2e` goes through the same C-ABI
calls with HOST (pinned) input buffers, the host->device copies inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GRID = (100, 100, 100)
NB, NS, NBINS = 256, 2, 20000
SEED = 2
LATTICE_A = 5.43
FLOPS_ADAPTIVE = 76.0   # per in-window contribution: gaussian 38 + erf intDOS 38
FLOPS_LINEAR = 30.0     # doslin, no transcendental
BYTES_PER_STATE = 32.0  # E (8) + gradient (24)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nk", type=int, default=0, help="override k-points per GPU (debug only; invalidates the number)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--partition", default="cyclic", choices=["cyclic", "block"],
                    help="k-point assignment over ranks: cyclic (balanced on the k1-ordered synthetic mesh) or contiguous blocks")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the configs 3-5 block")
    ap.add_argument("--no-parity", action="store_true", help="skip the N-rank vs 1-rank spectrum comparison (N > 1)")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples if len(s) >= 7 for i in range(4) if s[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


def dist_setup(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = dist_mod
    return rank, world, local, dist


def barrier(dist):
    if dist is not None:
        dist.barrier()


def max_over_ranks(dist, x, local):
    if dist is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64, device=torch.device("cuda", local))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(dist, x, local):
    if dist is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64, device=torch.device("cuda", local))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def cpu_sample(ob, ctx_factory, nk_sample, threads, emin, delta, contrib_per_k):
    """Time the restated reference loop (oracle, dense like the reference) on a bounded k-point sample of
    the SAME workload: same generator, seed, nb, ns and energy grid.  The dense loop costs the same for
    every k-point, so time extrapolates linearly in nk (exactly); the rate is quoted in the metric's unit
    with the FULL workload's in-window contributions per k-point."""
    from oracle import odo
    ctx = ctx_factory()
    grid = (GRID[0], GRID[1], GRID[2])
    ctx.synth_bands(grid, 0, nk_sample, NS, NB, lattice_a=LATTICE_A, seed=SEED)
    be, bg, kw = ctx.get_bands(), ctx.get_gradients(), ctx.get_kweights()
    contrib = contrib_per_k * nk_sample
    ctx.close()
    s = odo.Sys(be, kw, np.eye(3) * (2 * np.pi / LATTICE_A), grid, bg, electrons_per_state=1.0)
    E = emin + np.arange(NBINS, dtype=np.float64) * delta
    t0 = time.perf_counter()
    odo.calculate_dos("a", s, odo.make_broadening(), E, delta, nthreads=threads)
    odo.calculate_dos("l", s, odo.make_broadening(), E, delta, nthreads=threads)
    dt = time.perf_counter() - t0
    dense = 2.0 * nk_sample * NB * NS * NBINS
    return contrib / dt, dense / dt, dt


def run_reference(args, rank, world, local, dist):
    """--impl reference: the reference's own CPU implementation of the path.  The reference is Fortran and
    no Fortran compiler exists in this image, so this arm times the C oracle that restates its
    dense loops and is pinned to its golden outputs, threaded over k-points like the MPI build."""
    if rank != 0:
        return
    import optados_b200 as ob
    threads = os.cpu_count() or 1
    nk_sample = 96 * threads
    # the energy grid of the full workload (min/max of the cosine bands are analytic enough for a sample run:
    # take them from a full-size device pass so that the grid is EXACTLY the one the GPU arm uses)
    ctx = ob.Context(local)
    nk_full = GRID[0] * GRID[1] * GRID[2] if not args.nk else args.nk
    ctx.synth_bands(GRID, 0, nk_full, NS, NB, lattice_a=LATTICE_A, seed=SEED)
    mn, mx = ctx.band_minmax()
    emin, emax = mn - 5.0, mx + 5.0
    delta = (emax - emin) / (NBINS - 1)
    cpk = (ctx.count_contributions("a", emin, delta, NBINS) + ctx.count_contributions("l", emin, delta, NBINS)) / nk_full
    ctx.close()
    vals, times = [], []
    for i in range(args.warmup + args.steps):
        v, dense, dt = cpu_sample(ob, lambda: ob.Context(local), nk_sample, threads, emin, delta, cpk)
        if i >= args.warmup:
            vals.append(v)
            times.append(dt)
    value = float(np.mean(vals))
    sample = f"{nk_sample} k-points x {NB} bands x {NS} spins x {NBINS} bins, adaptive+linear, dense loop"
    line = {"impl": "reference", "metric": "in-window broadened (k,band,E) contributions/s", "value": value,
            "unit": "contributions/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": float(np.mean(times) * 1e3), "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args, args.gpus, nk_full * (args.gpus if args.scaling == "weak" else 1)),
            "cpu_baseline": {"value": value, "unit": "contributions/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "contributions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def workload_config(args, n_gpus, nk_total):
    return {"workload": "cfg2: synthetic cosine bands, adaptive + linear DOS with integrated DOS",
            "k_points_total": nk_total, "k_points_per_gpu": nk_total // n_gpus, "bands": NB, "spins": NS,
            "energy_bins": NBINS, "adaptive_smearing": 0.4, "finite_bin_correction": True, "seed": SEED,
            "sharding": (f"k-points over {n_gpus} GPU(s), " +
                         ("cyclic (k = rank + i N): the synthetic mesh is ordered by k1, contiguous blocks would differ by +-12 % in work"
                          if args.partition == "cyclic" else "contiguous blocks as algor_dist_array") +
                         ", one NCCL allreduce per broadening"),
            "l2": "inputs (16.4 GB / N per GPU) are larger than the 126 MB L2; no flush needed"}


_REAL_STDOUT = None


def emit(line):
    """the ONE JSON line of the contract, on the process's original stdout"""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # Libraries print to stdout on their own (torch's NCCL process group announces "NCCL version ..." there): everything
    # but the JSON line goes to stderr, so that stdout carries exactly one line.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    args = parse()
    if args.impl == "reference":
        # rank 0 alone runs the CPU arm; the other ranks leave at once WITHOUT joining a process group (ranks spinning in a
        # barrier would take host cores away from the loop being timed)
        rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
        if rank == 0:
            run_reference(args, rank, int(os.environ.get("WORLD_SIZE", "1")), local, None)
        return
    rank, world, local, dist = dist_setup(args)
    import optados_b200 as ob

    nk_named = args.nk if args.nk else GRID[0] * GRID[1] * GRID[2]
    if args.scaling == "weak":
        grid_total, nk_total = (GRID[0], GRID[1], GRID[2] * world), nk_named * world
    else:
        grid_total, nk_total = GRID, nk_named
    per = ob.dist_array(nk_total, world)            # algor_dist_array (algorithms.f90:309)
    nk_gpu, k_first = per[rank], sum(per[:rank])
    def new_ctx():
        """one context per rank; a communicator id serves ONE ncclCommInitRank round, so every new set of rank contexts gets
        a fresh id from rank 0 (collective: every rank calls this)"""
        if world == 1:
            return ob.Context(local)
        import torch
        buf = torch.zeros(128, dtype=torch.uint8, device=torch.device("cuda", local))
        if rank == 0:
            buf.copy_(torch.frombuffer(bytearray(ob.get_unique_id()), dtype=torch.uint8))
        dist.broadcast(buf, 0)
        return ob.Context(local, rank, world, bytes(buf.cpu().numpy().tobytes()))

    ctx = new_ctx()
    if args.partition == "cyclic":     # same counts as algor_dist_array, k = rank + i * world
        ctx.synth_bands(grid_total, rank, nk_gpu, NS, NB, lattice_a=LATTICE_A, seed=SEED, k_stride=world)
    else:
        ctx.synth_bands(grid_total, k_first, nk_gpu, NS, NB, lattice_a=LATTICE_A, seed=SEED)

    p = ctx.dos_params(["adaptive", "linear"], dos_nbins=NBINS, num_electrons=(NB / 2.0, NB / 2.0))

    def step():
        p.dos_nbins = NBINS
        return ctx.dos_utils_calculate(p)

    # ---- warm-up, then K timed steps (device stopwatch on the library's stream + barrier/sync on both sides)
    for _ in range(args.warmup):
        r = step()
    emin, delta = r["emin"], r["delta"]
    contrib_local = ctx.count_contributions("a", emin, delta, NBINS) + ctx.count_contributions("l", emin, delta, NBINS)
    contrib_a_local = ctx.count_contributions("a", emin, delta, NBINS)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ctx.kernel_launches()
    ctx.profile_reset()
    ctx.synchronize()
    barrier(dist)
    ctx.timer_start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = step()
    ms_dev = ctx.timer_stop()
    ctx.synchronize()
    barrier(dist)
    wall = time.perf_counter() - t0
    sampler.stop_flag = True
    launches = ctx.kernel_launches() - launches0
    prof = ctx.profile_get()
    ms_step = max_over_ranks(dist, ms_dev / args.steps, local)
    contrib_total = sum_over_ranks(dist, contrib_local, local)
    value = contrib_total / (ms_step * 1e-3)

    # ---- roofline, FP64-bound.  `frac`: the accumulation kernels alone (CUDA events around their launches);
    #      `frac_step`: the WHOLE timed step (prepare passes, merges, Fermi analysis included) against the same peak.
    fp64_peak = ctx.measure_fp64_peak(0.6)          # FP64 pipe peak: the larger of a DFMA and a DMMA loop
    fp64_dfma = ctx.measure_fp64_peak(0.3, mode=1)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    acc_ms = prof["accumulate_ms"] / max(1, args.steps)                       # per step, all accumulate launches
    prep_ms = prof["prepare_ms"] / max(1, args.steps)
    flops_step = contrib_a_local * FLOPS_ADAPTIVE + (contrib_local - contrib_a_local) * FLOPS_LINEAR
    achieved_tf = flops_step / (acc_ms * 1e-3) / 1e12 if acc_ms > 0 else 0.0
    step_tf = sum_over_ranks(dist, flops_step, local) / (ms_step * 1e-3) / 1e12 / world     # per GPU
    states = float(nk_gpu) * NB * NS
    hbm_gbs = 2.0 * states * BYTES_PER_STATE / (acc_ms * 1e-3) / 1e9 if acc_ms > 0 else 0.0
    traffic, kernels, pipe_pct = None, None, None
    try:   # per-kernel ncu figures of one k-point slab (tools/ncu_summary.py -> profiles/r02_kernels.json), scaled to this step's units
        kj = json.load(open(os.path.join(ROOT, "profiles", "r02_kernels.json")))
        kernels = {k: {"fp64_pipe_pct": v["fp64_pipe_pct"], "ms_per_launch": v["ms"] / max(1, v["launches"]),
                       "dram_bytes_per_unit": v["dram_bytes"] / v["launches"] / v["units"]}
                   for k, v in kj["kernels"].items()}
        # every kernel of the capture runs once per scheme and unit (prep_hist twice: once per scheme)
        traffic = sum(v["dram_bytes"] / v["units"] for v in kj["kernels"].values()) / kj.get("repeats", 1) * states
        acc = [v for k, v in kj["kernels"].items() if k.startswith("narrow_kernel")]
        pipe_pct = sum(v["fp64_pipe_pct"] * v["ms"] for v in acc) / sum(v["ms"] for v in acc)
    except Exception:
        pass
    roofline = {"bound": "fp64", "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": achieved_tf / fp64_peak if fp64_peak > 0 else None,
                "frac_note": "ALGORITHMIC flops (SURVEY.md 8(d): 76 per adaptive contribution with its integrated DOS, 30 per linear one) / "
                             "duration / peak.  It exceeds 1 since round 2: the integrated DOS of the wide Gaussians is no longer evaluated per "
                             "(state, bin) -- it follows from the summed DOS through a 40-tap Euler-Maclaurin filter (tools/em_fir.py) -- so the "
                             "kernels execute about a third of the reference's arithmetic.  Hardware utilisation: fp64_pipe_pct.",
                "fp64_pipe_pct": pipe_pct,
                "frac_step": step_tf / fp64_peak if fp64_peak > 0 else None, "achieved_step": step_tf,
                "traffic": traffic,
                "traffic_note": "DRAM bytes per step of ALL kernels of the step, prepare passes included (ncu dram__bytes_read+write per unit, "
                                "profiles/r02_kernels.json, x this rank's units); algorithmic input: 32 B per state and scheme",
                "peak_source": "measured in this run: od_measure_fp64_peak = max(DFMA loop, DMMA m16n8k8 loop) on all SMs; "
                               "MEASURED_PEAKS.json has no FP64 entry",
                "peak_dfma_loop": fp64_dfma,
                "kernel": "accumulation kernels (narrow_kernel<0,0> adaptive + narrow_kernel<1,0> linear), CUDA events on the library stream",
                "kernel_ms_per_step": acc_ms, "kernel_share_of_step": acc_ms / ms_step if ms_step > 0 else None,
                "prepare_ms_per_step": prep_ms, "kernels_ncu": kernels,
                "algorithmic_flops_per_contribution": {"adaptive+intdos": FLOPS_ADAPTIVE, "linear": FLOPS_LINEAR},
                "lane_efficiency": prof["contributions"] / prof["lane_steps"] if prof.get("lane_steps") else None,
                "hbm": {"achieved_gbs": hbm_gbs, "peak_gbs": hbm_peak, "frac": hbm_gbs / hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s"}}

    # ---- end to end: host (pinned) inputs through the same C-ABI calls, H2D inside the timed region
    e2e = None
    if not args.no_e2e:
        nbe, nbg = nk_gpu * NB * NS * 8, nk_gpu * NB * NS * 3 * 8
        h_be, h_bg = ob.host_alloc_pinned(nbe), ob.host_alloc_pinned(nbg)
        ctx.get_bands_ptr(h_be)
        ctx.get_gradients_ptr(h_bg)

        def e2e_step():
            ctx.set_bands_ptr(h_be)                       # 4.1 GB / N, asynchronous on the library stream
            ctx.set_gradients_ptr(h_bg, deferred=True)    # 12.3 GB / N, streamed slab by slab under the kernels
            return step()

        e2e_step()
        ctx.synchronize()
        barrier(dist)
        t0 = time.perf_counter()
        n_e2e = max(1, min(args.steps, 3))
        for _ in range(n_e2e):
            r = e2e_step()
        ctx.synchronize()
        barrier(dist)
        dt = max_over_ranks(dist, (time.perf_counter() - t0) / n_e2e, local)
        d2h = 2 * 2 * NBINS * NS * 8
        e2e = {"value": contrib_total / dt, "unit": "contributions/s", "h2d_bytes_per_step": nbe + nbg,
               "d2h_bytes_per_step": d2h, "ms_per_step": dt * 1e3, "steps": n_e2e,
               "api": "od_set_bands + od_set_gradients(OD_MEM_HOST_DEFERRED) from pinned host memory + od_dos_utils_calculate + od_dos_get"}
        # OD_MEM_HOST_DEFERRED leaves the library reading h_bg on every later call: make the gradients resident again
        # (one plain upload) before the host buffers go away.  That upload doubles as the measurement of what the box can
        # deliver: all N ranks copy their 12.3 GB / N of pinned memory at the same time and nothing else runs -- the
        # concurrent host-to-device ceiling the end-to-end number is to be read against.
        ctx.synchronize()
        barrier(dist)
        t0 = time.perf_counter()
        ctx.set_gradients_ptr(h_bg, deferred=False)
        ctx.synchronize()
        dt_copy = time.perf_counter() - t0
        barrier(dist)
        dt_copy_max = max_over_ranks(dist, dt_copy, local)
        e2e["h2d_ceiling_gbs"] = sum_over_ranks(dist, float(nbg), local) / dt_copy_max / 1e9      # all ranks together
        e2e["h2d_ceiling_gbs_per_gpu"] = nbg / dt_copy / 1e9
        e2e["h2d_floor_ms_per_step"] = sum_over_ranks(dist, float(nbe + nbg), local) / (e2e["h2d_ceiling_gbs"] * 1e9) * 1e3
        ob.host_free_pinned(h_be)
        ob.host_free_pinned(h_bg)

    # ---- the reference's own decomposition beside the cyclic deal: contiguous k-point blocks (algor_dist_array), which on the
    #      k1-ordered synthetic mesh carry unequal work (the windows follow sin(2 pi k1))
    block = None
    if world > 1 and args.partition == "cyclic":
        ctx.synth_bands(grid_total, k_first, nk_gpu, NS, NB, lattice_a=LATTICE_A, seed=SEED)
        step()
        ctx.synchronize()
        barrier(dist)
        ctx.timer_start()
        for _ in range(args.steps):
            step()
        ms_blk = max_over_ranks(dist, ctx.timer_stop() / args.steps, local)
        barrier(dist)
        block = {"partition": "contiguous blocks (algor_dist_array)", "ms_per_step": ms_blk, "value": contrib_total / (ms_blk * 1e-3)}
        ctx.synth_bands(grid_total, rank, nk_gpu, NS, NB, lattice_a=LATTICE_A, seed=SEED, k_stride=world)

    # ---- multi-rank parity of the LIBRARY (not the oracle): rank 0 runs the whole mesh on one communicator-less ctx and
    #      compares the spectra the N ranks produced through od_dos_utils_calculate + the NCCL merge with it
    parity = None
    if world > 1 and not args.no_parity:
        r_multi = step()                       # collective: every rank calls it
        if rank == 0:
            ctx1 = ob.Context(local)
            ctx1.synth_bands(grid_total, 0, nk_total, NS, NB, lattice_a=LATTICE_A, seed=SEED)
            p1 = ctx1.dos_params(["adaptive", "linear"], dos_nbins=NBINS, num_electrons=(NB / 2.0, NB / 2.0))
            r_one = ctx1.dos_utils_calculate(p1)
            ctx1.close()
            parity = {"reference": "the same library on ONE GPU over all k-points, no communicator",
                      "nbins_equal": bool(r_one["nbins"] == r_multi["nbins"] and r_one["emin"] == r_multi["emin"])}
            for k in ("dos_adaptive", "dos_linear"):
                parity["max_rel_" + k] = float(np.abs(r_multi[k] - r_one[k]).max() / np.abs(r_one[k]).max())
            for k in ("intdos_adaptive", "intdos_linear"):
                parity["max_rel_" + k] = float(np.abs(r_multi[k] - r_one[k]).max() / np.abs(r_one[k]).max())
            parity["efermi_abs_diff"] = float(max(abs(r_multi["efermi_adaptive"] - r_one["efermi_adaptive"]),
                                                  abs(r_multi["efermi_linear"] - r_one["efermi_linear"])))
        barrier(dist)

    # ---- CPU baseline beside it (rank 0, N = 1 only): restated reference loop on a bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        nk_sample = 96 * threads
        v, dense, dt = cpu_sample(ob, lambda: ob.Context(local), nk_sample, threads, emin, delta, contrib_local / nk_gpu)
        cpu = {"value": v, "unit": "contributions/s", "cores": threads, "kind": "port",
               "sample": f"{nk_sample} of {nk_gpu} k-points (same nb, ns, bins, seed), dense reference loop, {dt:.1f} s",
               "dense_evaluations_per_s": dense}

    # ---- the other BASELINE configurations (3: PDOS, 4: JDOS + optics, 5: core-loss): not bench lines, but driver-visible
    #      numbers with their own roofline fractions and a k-slab parity figure against the oracle
    configs = None
    ctx.close()
    ctx = None
    if not args.no_configs:
        import gc
        gc.collect()
        import importlib.util
        spec = importlib.util.spec_from_file_location("od_bench_configs", os.path.join(ROOT, "tools", "bench_configs.py"))
        bc = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(bc)
        env = bc.Env(local=local, rank=rank, world=world,
                     make_ctx=new_ctx,
                     max_over_ranks=lambda x: max_over_ranks(dist, x, local), sum_over_ranks=lambda x: sum_over_ranks(dist, x, local),
                     barrier=lambda: barrier(dist), fp64_peak=fp64_peak, hbm_peak=hbm_peak, reps=2, parity=True)
        configs = {}
        for name in ("cfg3", "cfg4", "cfg5"):
            try:
                configs[name] = bc.RUNNERS[name](env)
            except Exception as e:     # a config that does not fit / fails must not take the headline down with it
                configs[name] = {"error": repr(e)}
            try:
                import torch
                torch.cuda.empty_cache()
            except Exception:
                pass
            barrier(dist)

    if rank == 0:
        line = {"metric": "in-window broadened (k,band,E) contributions/s", "value": value, "unit": "contributions/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(args, world, nk_total), "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
                "configs": configs,
                "multi_gpu_parity": parity, "block_partition": block, "gpu_launches": int(launches), "clocks": sampler.summary(),
                "dense_equivalent_evaluations_per_s": 2.0 * float(nk_total) * NB * NS * NBINS / (ms_step * 1e-3),
                "contributions_per_step": contrib_total, "wall_ms_per_step": wall / args.steps * 1e3,
                "efermi_adaptive": r.get("efermi_adaptive"), "efermi_linear": r.get("efermi_linear")}
        emit(line)
    if ctx is not None:
        ctx.close()
    barrier(dist)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
