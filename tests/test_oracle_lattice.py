"""The oracle's lattice index order, pinned WITHOUT the reference's fixtures.

Every fixture of the reference (Si2, NaGe) has a symmetric reciprocal-lattice matrix, so a row/column swap of
`recip_lattice` in `doslin_sub_cell_corners` (dos_utils.f90:1355-1387) or in the adaptive width
(dos_utils.f90:1210-1216) cannot fail a golden file.  This test restates the two formulas literally in numpy
-- `matmul(recip_lattice, stepp)` sums over the SECOND index, `sub_cell_length(i)` is the norm of ROW i --
on a non-symmetric, non-orthogonal lattice with an anisotropic k-grid and compares the C oracle with it.
The GPU tests then compare the library with the oracle on the same lattice (tests/test_gpu_parity.py).
"""
import numpy as np
from scipy.special import erf

RECIP = np.array([[1.2, -0.6, 0.6], [0.0, 1.2, -1.2], [0.0, 0.0, 1.2]])   # recip_lattice(i, j), rows = b_i
KGRID = (7, 5, 3)


def corners_numpy(grad, energy):
    """dos_utils.f90:1320-1388, literally: stepp = matmul(recip_lattice, step * sigma); DE = dot(grad, stepp)."""
    step = 1.0 / np.asarray(KGRID, dtype=np.float64) / 2.0
    DE = []
    for m in (-1, 1):
        for n in (-1, 1):
            for o in (-1, 1):
                stepp = step * np.array([m, n, o], dtype=np.float64)
                stepp = RECIP @ stepp                       # stepp(i) = sum_j recip_lattice(i,j) stepp(j)
                DE.append(float(np.dot(grad, stepp)))
    DE = np.sort(np.array(DE))[::-1]                        # descending
    return np.array([energy, energy + DE[4], energy + DE[5], energy + DE[6], energy + DE[7]])


def doslin_numpy(EV, e):
    """dos_utils.f90:1391-1528 for one state at the energies e (vector)."""
    e0, e1, e2, e3, e4 = EV
    alpha = 1.0 / ((e1 - e3) * (e1 - e4) + (e3 - e4) ** 2 / 3.0 - (e1 - e2) ** 2 / 3.0 + (e1 + e2 - e3 - e4) * (e0 - e1))
    d = np.zeros_like(e)
    c = np.zeros_like(e)
    for i, en in enumerate(e):
        if en <= e4:
            continue
        if en >= 2.0 * e0 - e4:
            c[i] = 1.0
            continue
        shift = en > e0
        u = 2.0 * e0 - en if shift else en
        if u <= e4:
            dl, il = 0.0, 0.0
        elif u < e3:
            dl = (u - e4) ** 2 / (e3 - e4) / 2.0
            il = (u - e4) ** 3 / (e3 - e4) / 6.0
        elif u < e2:
            dl = u - (e3 + e4) / 2.0
            il = ((e3 - e4) ** 2 / 3.0 + (u - e3) * (u - e4)) / 2.0
        elif u < e1:
            dl = (e1 + e2 - e3 - e4) / 2.0 - (e1 - u) ** 2 / (e1 - e2) / 2.0
            il = ((e2 - e4) * (u - e3) + (e1 - e3) * (u - e2) + (e3 - e4) ** 2 / 3.0 + ((e1 - u) ** 3 - (e1 - e2) ** 3) / (e1 - e2) / 3.0) / 2.0
        else:
            dl = (e1 + e2 - e3 - e4) / 2.0
            il = ((e1 - e3) * (e1 - e4) + (e3 - e4) ** 2 / 3.0 - (e1 - e2) ** 2 / 3.0 + (e1 + e2 - e3 - e4) * (u - e1)) / 2.0
        d[i] = dl * alpha
        c[i] = 1.0 - il * alpha if shift else il * alpha
    return d, c


def make_sys(odo, nk=6, nb=5, ns=1, seed=3):
    rng = np.random.default_rng(seed)
    be = np.sort(rng.uniform(-3.0, 3.0, (nb, ns, nk)), axis=0)
    bg = rng.normal(0.0, 2.0, (nb, 3, nk, ns))
    kw = rng.uniform(0.5, 1.5, nk)
    kw /= kw.sum()
    return odo.Sys(be, kw, RECIP, KGRID, bg, electrons_per_state=2.0)


def test_oracle_corners_index_order(odo):
    s = make_sys(odo)
    E = np.linspace(-6.0, 6.0, 1201)
    delta = E[1] - E[0]
    dos, intdos, _ = odo.calculate_dos("l", s, odo.make_broadening(), E, delta)
    ref_d = np.zeros_like(E)
    ref_i = np.zeros_like(E)
    for ik in range(s.nk):
        for ib in range(s.nb):
            EV = corners_numpy(s.band_gradient[ib, :, ik, 0], s.band_energy[ib, 0, ik])
            d, c = doslin_numpy(EV, E)
            ref_d += d * s.kpoint_weight[ik] * 2.0
            ref_i += c * s.kpoint_weight[ik] * 2.0
    assert np.abs(dos[:, 0] - ref_d).max() < 1e-12 * np.abs(ref_d).max()
    assert np.abs(intdos[:, 0] - ref_i).max() < 1e-12 * np.abs(ref_i).max()
    # the transposed lattice gives a different answer on this cell (the test can fail)
    st = odo.Sys(s.band_energy, s.kpoint_weight, RECIP.T, KGRID, s.band_gradient, electrons_per_state=2.0)
    dos_t, _, _ = odo.calculate_dos("l", st, odo.make_broadening(), E, delta)
    assert np.abs(dos_t[:, 0] - ref_d).max() > 1e-3 * np.abs(ref_d).max()


def test_oracle_adaptive_width_index_order(odo):
    s = make_sys(odo, seed=4)
    E = np.linspace(-6.0, 6.0, 601)
    delta = E[1] - E[0]
    smear = 0.9
    dos, intdos, _ = odo.calculate_dos("a", s, odo.make_broadening(adaptive_smearing=smear, finite_bin_correction=False), E, delta)
    step = 1.0 / np.asarray(KGRID, dtype=np.float64) / 2.0
    sub = np.sqrt((RECIP ** 2).sum(axis=1)) * step          # norm of ROW i (dos_utils.f90:1213)
    temp = smear * sub.sum() / 3.0
    ref_d = np.zeros_like(E)
    ref_i = np.zeros_like(E)
    for ik in range(s.nk):
        for ib in range(s.nb):
            g = s.band_gradient[ib, :, ik, 0]
            w = np.sqrt(np.dot(g, g)) * temp
            z = (E - s.band_energy[ib, 0, ik]) / w
            gs = np.where(0.5 * z * z > 30.0, 0.0, np.exp(-0.5 * z * z) / (np.sqrt(2 * np.pi) * w))
            ref_d += gs * s.kpoint_weight[ik] * 2.0
            ref_i += 0.5 * (1.0 + erf(z / np.sqrt(2.0))) * s.kpoint_weight[ik] * 2.0
    assert np.abs(dos[:, 0] - ref_d).max() < 1e-12 * np.abs(ref_d).max()
    assert np.abs(intdos[:, 0] - ref_i).max() < 1e-12 * np.abs(ref_i).max()
    col = np.sqrt((RECIP ** 2).sum(axis=0)) * step          # column norms differ on this lattice
    assert abs(col.sum() - sub.sum()) > 1e-2
