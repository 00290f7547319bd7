"""Worker of tests/test_multigpu.py: one process per GPU (torch.distributed.run), k-points in the contiguous blocks of
algor_dist_array (algorithms.f90:309); every merged result of the LIBRARY (the NCCL allreduce that replaces comms_reduce,
comms.F90:556-602) is compared on rank 0 with the same library run on ONE communicator-less ctx over all k-points."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import optados_b200 as ob
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    buf = torch.zeros(128, dtype=torch.uint8, device=torch.device("cuda", local))
    if rank == 0:
        buf.copy_(torch.frombuffer(bytearray(ob.get_unique_id()), dtype=torch.uint8))
    dist.broadcast(buf, 0)
    ctx = ob.Context(local, rank, world, bytes(buf.cpu().numpy().tobytes()))

    grid, ns, nb, nbins = (24, 20, 18), 2, 24, 6000
    nk = grid[0] * grid[1] * grid[2] + 0
    per = ob.dist_array(nk, world)
    nk_r, k0 = per[rank], sum(per[:rank])
    rng = np.random.default_rng(7)
    om_full = None

    def setup(c, k_first, n):
        c.synth_bands(grid, k_first, n, ns, nb, seed=5)

    def run(c, k_first, n, nk_global):
        out = {}
        p = c.dos_params(["fixed", "adaptive", "linear"], dos_nbins=nbins, num_electrons=(nb / 2.0, nb / 2.0),
                         br=ob.broadening(fixed_smearing=0.05))
        r = c.dos_utils_calculate(p)
        for k in ("dos_fixed", "dos_adaptive", "dos_linear", "intdos_fixed", "intdos_adaptive", "intdos_linear"):
            out[k] = r[k]
        out["efermi"] = np.array([r["efermi_fixed"], r["efermi_adaptive"], r["efermi_linear"]])
        ef = r["efermi_adaptive"]
        out["dos_at_e"] = c.dos_utils_calculate_at_e(p, ef, r["delta"])[0]
        bg = c.compute_bandgap(ef, (nb / 2.0, nb / 2.0), nk_global=nk_global, k_first=k_first)
        out["bandgap"] = np.array([bg["thermal_bandgap"], bg["thermal_vbm_k"], bg["thermal_cbm_k"]] + bg["optical_bandgap"] +
                                  bg["average_bandgap"] + bg["vbm_multiplicity"] + bg["cbm_multiplicity"], dtype=np.float64)
        be, bec = c.compute_band_energies(ef)
        out["band_energies"] = np.array([be["fixed"], be["adaptive"], be["linear"], bec])
        # JDOS + fused optics (tensor: 6 components, the case jdos_utils_merge gets wrong under MPI, quirk Q18)
        rs = np.random.default_rng(11)
        om = rs.normal(0, 0.1, (nb, nb, 3, nk_global, ns)) + 1j * rs.normal(0, 0.1, (nb, nb, 3, nk_global, ns))
        om = np.asfortranarray(om[:, :, :, k_first:k_first + n, :])
        for t in ("a", "l"):
            jd, wj = c.calculate_jdos_optics(t, 0.004, 3500, ef, om, "tensor", merge=True)
            out["jdos_" + t], out["wjdos_" + t] = jd, wj
        mw = np.asfortranarray(np.random.default_rng(13).random((5, nb, nk_global, ns))[:, :, k_first:k_first + n, :])
        d, i, w = c.calculate_dos("a", r["emin"], r["delta"], nbins, mw=mw, merge=True)
        out["wdos"] = w
        return out

    setup(ctx, k0, nk_r)
    multi = run(ctx, k0, nk_r, nk)
    res = None
    if rank == 0:
        c1 = ob.Context(local)
        setup(c1, 0, nk)
        one = run(c1, 0, nk, nk)
        c1.close()
        res = {k: float(np.abs(multi[k] - one[k]).max() / max(np.abs(one[k]).max(), 1e-300)) for k in one}
        res["world"] = world
        print("MGPU_RESULT " + json.dumps(res), flush=True)
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
