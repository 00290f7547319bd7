"""Helpers shared by the tests: fixtures -> oracle systems, synthetic band structures."""
import numpy as np

from conftest import load_npz


def oracle_sys(odo, fx, gradient="auto"):
    """Build an oracle Sys from a tests/golden/fx_*.npz fixture dict."""
    grad = None
    if gradient == "auto":
        if "band_gradient" in fx:
            grad = fx["band_gradient"]
        elif "optical_mat" in fx:
            grad = odo.gradient_from_ome(fx["optical_mat"])  # electronic.f90:280-291
    elif gradient is not None:
        grad = gradient
    return odo.Sys(fx["band_energy"], fx["kpoint_weight"], fx["recip_lattice"], fx["kpoint_grid_dim"], grad,
                   electrons_per_state=float(fx["electrons_per_state"]), num_electrons=fx["num_electrons"],
                   cell_volume=float(fx["cell_volume"]))


def fixture(name):
    return load_npz("fx_" + name)


def pdos_mask(fx, projectors):
    """projectors: list of dicts(species=?, rank=?, am=?) with None = wildcard
    -> (norb, nproj) 0/1 mask = projection_array(species(orb), rank(orb), am(orb)+1, proj)."""
    sp, rk, am = fx["pdos_species"], fx["pdos_rank"], fx["pdos_am"]
    m = np.zeros((sp.size, len(projectors)), dtype=np.int32, order="F")
    for p, sel in enumerate(projectors):
        for o in range(sp.size):
            ok = all(sel.get(k) is None or sel[k] == v for k, v in (("species", sp[o]), ("rank", rk[o]), ("am", am[o])))
            m[o, p] = 1 if ok else 0
    return m


def splitmix64(seed, n):
    """n uniform doubles in [0,1) from SplitMix64(seed) (SURVEY.md section 8(d) synthetic inputs)."""
    out = np.empty(n, dtype=np.float64)
    x = np.uint64(seed)
    with np.errstate(over="ignore"):
        for i in range(n):
            x = x + np.uint64(0x9E3779B97F4A7C15)
            z = x
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z = z ^ (z >> np.uint64(31))
            out[i] = float(z >> np.uint64(11)) * (1.0 / 9007199254740992.0)
    return out
