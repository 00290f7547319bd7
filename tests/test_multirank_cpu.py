"""N > 1 host logic on CPU (gloo, world_size 2): k-points partitioned exactly like algor_dist_array
(algorithms.f90:309), every rank accumulates its shard's partial spectra on the energy scale derived
from the MIN/MAX-reduced band extrema (dos_utils.f90:964-977), one SUM allreduce replaces comms_reduce
(comms.F90:556) -- and the result equals the single-rank spectrum.  The per-shard arithmetic is the
oracle's here (no GPU in this container); the same schedule runs on the GPUs in bench.py / the library."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import odo
    import optados_b200 as ob
    fx = dict(np.load(os.path.join(ROOT, "tests", "golden", "fx_si2_core.npz")))
    nk = int(fx["nk"])
    per = ob.dist_array(nk, world)
    k0 = sum(per[:rank])
    full = odo.Sys(fx["band_energy"], fx["kpoint_weight"], fx["recip_lattice"], fx["kpoint_grid_dim"], fx["band_gradient"],
                   electrons_per_state=float(fx["electrons_per_state"]), num_electrons=fx["num_electrons"])
    shard = full.kslice(k0, k0 + per[rank])
    # setup_energy_scale: local extrema, MIN / MAX reduce + bcast
    mm = torch.tensor([shard.band_energy.min(), -shard.band_energy.max()], dtype=torch.float64)
    dist.all_reduce(mm, op=dist.ReduceOp.MIN)
    E, delta, n = odo.dos_energy_scale(shard, 0.05, minmax=(float(mm[0]), float(-mm[1])))
    res = {}
    for t in ("a", "l"):
        dos, intdos, _ = odo.calculate_dos(t, shard, odo.make_broadening(), E, delta)
        buf = torch.from_numpy(np.concatenate([dos.ravel(order="F"), intdos.ravel(order="F")]))
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)            # dos_utils_merge
        res[t] = buf.numpy().copy()
    if rank == 0:
        E1, delta1, n1 = odo.dos_energy_scale(full, 0.05)
        assert n1 == n and abs(delta1 - delta) < 1e-15
        for t in ("a", "l"):
            dos, intdos, _ = odo.calculate_dos(t, full, odo.make_broadening(), E1, delta1)
            ref = np.concatenate([dos.ravel(order="F"), intdos.ravel(order="F")])
            out[t] = float(np.abs(res[t] - ref).max() / np.abs(ref).max())
            if t == "a":
                ef_one = odo.calc_efermi_from_intdos(intdos, E1, full.num_electrons)
        ef_multi = odo.calc_efermi_from_intdos(res["a"][n * full.ns:].reshape((n, full.ns), order="F"), E, full.num_electrons)
        out["efermi"] = abs(ef_multi - ef_one)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_shard_and_allreduce():
    from oracle import odo
    odo.build()
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29600 + (os.getpid() % 200)
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert out["a"] < 1e-13 and out["l"] < 1e-13
    assert out["efermi"] < 1e-12


def test_dist_array_partitions_cover_all_kpoints():
    import optados_b200 as ob
    for nk, n in ((14, 2), (14, 4), (1000000, 8), (7, 7)):
        per = ob.dist_array(nk, n)
        assert sum(per) == nk and max(per) - min(per) <= 1 and per == sorted(per, reverse=True)
