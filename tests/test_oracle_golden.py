"""Pin the CPU oracle against the reference's own golden outputs
(optados/test-suite/tests/<t>/benchmark.out.default.inp=<seed>.odi, extracted by
tests/golden/make_golden.py).  CPU only.  If these pass the oracle IS the reference's
arithmetic for this path to the print precision of the golden files.
"""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_npz, rel_to_peak
from fxutil import fixture, oracle_sys, pdos_mask

E21_13 = 2e-12     # golden printed with E21.13 -> 13 significant digits
LIST17 = 5e-13     # list-directed, 17 digits
ES14_7 = 2e-7      # es14.7


def gold(test):
    return load_npz("gold_" + test)


def test_dos_adaptive_dat_spin(odo):
    """testopt_dos_adaptive_dat = BASELINE config 1 (Si2 adaptive total DOS)."""
    fx = fixture("si2_optics")
    s = oracle_sys(odo, fx)
    assert tuple(fx["kpoint_grid_dim"]) == (3, 3, 3)
    r = odo.dos_utils_calculate(s, "adaptive", dos_spacing=0.1)
    g = gold("testopt_dos_adaptive_dat")["block0"]
    assert r["nbins"] == g.shape[0] == 482
    assert np.abs((r["E"] - r["efermi"]) - g[:, 0]).max() < 1e-10
    assert rel_to_peak(r["dos"][:, 0], g[:, 1]) < E21_13
    assert rel_to_peak(-r["dos"][:, 1], g[:, 2]) < E21_13
    assert rel_to_peak(r["intdos"][:, 0], g[:, 3]) < E21_13
    assert rel_to_peak(-r["intdos"][:, 1], g[:, 4]) < E21_13


def test_dos_adaptive_no_spin(odo):
    fx = fixture("si2_optics_no_spin")
    s = oracle_sys(odo, fx)
    r = odo.dos_utils_calculate(s, "adaptive", dos_spacing=0.1)
    g = gold("testopt_dos_adaptive_no_spin")["block0"]
    assert r["nbins"] == g.shape[0]
    assert np.abs((r["E"] - r["efermi"]) - g[:, 0]).max() < 1e-10
    assert rel_to_peak(r["dos"][:, 0], g[:, 1]) < E21_13
    assert rel_to_peak(r["intdos"][:, 0], g[:, 2]) < E21_13


def test_dos_linear_dome(odo):
    fx = fixture("si2_dos")
    s = oracle_sys(odo, fx)
    r = odo.dos_utils_calculate(s, "linear", dos_spacing=0.1)
    g = gold("testopt_dos_linear_dome")["block0"]
    assert r["nbins"] == g.shape[0]
    assert np.abs((r["E"] - r["efermi"]) - g[:, 0]).max() < 1e-10
    assert rel_to_peak(r["dos"][:, 0], g[:, 1]) < E21_13
    assert rel_to_peak(-r["dos"][:, 1], g[:, 2]) < E21_13
    assert rel_to_peak(r["intdos"][:, 0], g[:, 3]) < E21_13
    assert rel_to_peak(-r["intdos"][:, 1], g[:, 4]) < E21_13


def test_odo_fermi_energies(odo):
    """.odo goldens: 'Fermi energy (X broadening)' printed f8.4."""
    odo_gold = json.load(open(os.path.join(GOLDEN, "gold_odo.json")))
    fx = fixture("si2_optics")
    s = oracle_sys(odo, fx)
    for test, btype in (("testopt_dos_adaptive", "adaptive"), ("testopt_dos_linear", "linear"),
                        ("testopt_jdos_fixed_odo", "fixed"), ("testopt_pdos_string4", "adaptive"),
                        ("testopt_optics_poly", "fixed")):
        r = odo.dos_utils_calculate(s, btype, dos_spacing=0.1)
        assert abs(r["efermi"] - odo_gold[test]["efermi_" + btype]) < 5.1e-5, test
    fxc = fixture("si2_core")
    sc = oracle_sys(odo, fxc)
    # core: the golden .odo prints the Fermi analysis twice -- 4.3195 from dos_utils_set_efermi's pass and
    # 4.3173 (the value the regex keeps, the last one) from the weighted pass on the re-gridded
    # energy scale (sticky dos_nbins, quirk Q4: dos_utils.f90:980-994).
    r = odo.dos_utils_calculate(sc, "adaptive", dos_spacing=0.01)
    assert abs(r["efermi"] - 4.3195) < 5.1e-5
    E2, delta2, n2 = odo.dos_energy_scale(sc, 0.01, dos_nbins=r["nbins"])
    assert n2 == r["nbins"] and abs(delta2 - r["delta"]) > 1e-9
    dos2, intdos2, _ = odo.calculate_dos("a", sc, odo.make_broadening(), E2, delta2)
    ef2 = odo.calc_efermi_from_intdos(intdos2, E2, sc.num_electrons)
    assert abs(ef2 - odo_gold["testopt_core_absorb"]["efermi_adaptive"]) < 5.1e-5


@pytest.mark.parametrize("test,fxname,btype,tag", [("testopt_dos_adaptive", "si2_optics", "adaptive", "BEA"),
                                                   ("testopt_dos_linear", "si2_optics", "linear", "BEL"),
                                                   ("testopt_dos_fixed", "si2_fixed", "fixed", "BEF")])
def test_odo_bandgap_and_band_energies(odo, test, fxname, btype, tag):
    """.odo goldens of the analysis routines: dos_utils_compute_bandgap (dos_utils.f90:457-750: TBg / OBg / ABg / wAB,
    f15.10), dos_utils_compute_band_energies (:853-924: BE?, BEC, f12.4) and the DOS at the Fermi level (:391-455, f8.4)."""
    G = json.load(open(os.path.join(GOLDEN, "gold_odo.json")))[test]
    fx = fixture(fxname)
    s = oracle_sys(odo, fx, gradient=None if fxname == "si2_fixed" else "auto")
    r = odo.dos_utils_calculate(s, btype, dos_spacing=0.1)
    ef = r["efermi"]
    assert abs(ef - G["efermi_" + btype]) < 5.1e-5
    bg = odo.compute_bandgap(s, ef)
    assert abs(bg["thermal_bandgap"] - G["TBg"]) < 5.1e-11
    for is_ in range(s.ns):
        assert abs(bg["optical_bandgap"][is_] - G["OBg"][is_]) < 5.1e-11
        assert abs(bg["average_bandgap"][is_] - G["ABg"][is_]) < 5.1e-11
        assert [bg["vbm_multiplicity"][is_], bg["cbm_multiplicity"][is_]] == G["vbm_cbm_multiplicity"][is_]
    if s.ns > 1:
        assert abs(bg["weighted_average"] - G["wAB"]) < 5.1e-11
    assert (bg["thermal_vbm_k"] == bg["thermal_cbm_k"]) == (G["gap_kind"] == "Direct Gap")
    assert abs(odo.calc_band_energies(r["dos"], r["E"], r["delta"], ef) - G[tag]) < 5.1e-5
    assert abs(odo.band_energy_from_bands(s, ef) - G["BEC"]) < 5.1e-5
    d_at_e, _ = odo.calculate_dos_at_e(btype[0], s, odo.make_broadening(), r["delta"], ef)
    for is_ in range(s.ns):
        assert abs(d_at_e[is_] - G["DE" + btype[0].upper()][is_]) < 5.1e-5


def test_jdos_adaptive_dome(odo):
    fx = fixture("si2_dos")
    s = oracle_sys(odo, fx)
    ef = odo.dos_utils_calculate(s, "adaptive", dos_spacing=0.1)["efermi"]
    r = odo.jdos_utils_calculate(s, ef, "adaptive")
    g = gold("testopt_jdos_adaptive_dome")["block0"]
    assert r["nbins"] == g.shape[0]
    assert np.abs(r["E"] - g[:, 0]).max() < 1e-11
    assert rel_to_peak(r["jdos"][:, 0], g[:, 1]) < E21_13
    assert rel_to_peak(-r["jdos"][:, 1], g[:, 2]) < E21_13


def test_jdos_linear_no_spin(odo):
    fx = fixture("si2_optics_no_spin")
    s = oracle_sys(odo, fx)
    ef = odo.dos_utils_calculate(s, "linear", dos_spacing=0.1)["efermi"]
    r = odo.jdos_utils_calculate(s, ef, "linear")
    g = gold("testopt_jdos_linear_no_spin")["block0"]
    assert r["nbins"] == g.shape[0]
    assert np.abs(r["E"] - g[:, 0]).max() < 1e-11
    assert rel_to_peak(r["jdos"][:, 0], g[:, 1]) < E21_13


@pytest.mark.parametrize("test,fxname,btype,projectors", [
    ("testopt_pdos_angular", "si2_optics", "adaptive", [dict(am=l) for l in range(4)]),
    ("testopt_pdos_string2", "si2_dos", "adaptive", [dict(species=1, rank=1, am=1)]),
    ("testopt_pdos_string3", "si2_optics_no_spin", "adaptive", [dict(species=1, rank=1), dict(species=1, rank=2, am=0)]),
])
def test_pdos(odo, test, fxname, btype, projectors):
    fx = fixture(fxname)
    s = oracle_sys(odo, fx)
    mw = odo.projection_merge(fx["pdos_weights"], pdos_mask(fx, projectors))
    r = odo.dos_utils_calculate(s, btype, dos_spacing=0.1, mw=mw)
    g = gold(test)["block0"]
    nproj = len(projectors)
    assert r["nbins"] == g.shape[0]
    assert np.abs((r["E"] - r["efermi"]) - g[:, 0]).max() < 2e-6
    for p in range(nproj):
        assert rel_to_peak(r["wdos"][:, 0, p], g[:, 1 + p]) < ES14_7
        if s.ns == 2:
            assert rel_to_peak(-r["wdos"][:, 1, p], g[:, 1 + nproj + p]) < ES14_7


def _optics(odo, geom, btype, qdir=(0.5, 0.5, 0.25)):
    fx = fixture("si2_optics")
    s = oracle_sys(odo, fx)
    ef = odo.dos_utils_calculate(s, btype, dos_spacing=0.1)["efermi"]
    mw = odo.make_weights(s, fx["optical_mat"], geom, ef, qdir=qdir, symops=fx["symops"])
    r = odo.jdos_utils_calculate(s, ef, btype, jdos_spacing=0.01, jdos_max_energy=30.0, mw=mw)
    e2 = odo.calc_epsilon_2(r["wjdos"], s.cell_volume)
    e1 = odo.calc_epsilon_1(e2, r["E"])
    return r, e1, e2


def test_optics_polar_epsilon(odo):
    r, e1, e2 = _optics(odo, "polar", "adaptive")
    g = gold("testopt_optics_polar")["block0"]
    assert r["nbins"] == g.shape[0] == 3000
    assert np.abs(r["E"] - g[:, 0]).max() < 1e-12
    assert rel_to_peak(e2[:, 0], g[:, 2]) < LIST17
    assert rel_to_peak(e1[:, 0], g[:, 1]) < 1e-12


def test_optics_unpolar_loss_fn(odo):
    """Si2_loss_fn.dat of testopt_optics_unpolar: E, loss function, Gaussian-broadened loss function (FWHM 1 eV)"""
    r, e1, e2 = _optics(odo, "unpolar", "linear")
    g = gold("testopt_optics_unpolar")["block0"]
    lf, lb = odo.calc_loss_fn(e1[:, 0], e2[:, 0], r["E"], lossfn_gaussian=1.0)
    assert np.abs(r["E"] - g[:, 0]).max() < 1e-12
    assert rel_to_peak(lf, g[:, 1]) < 1e-11
    assert rel_to_peak(lb, g[:, 2]) < 1e-11


def test_optics_sum_rules(odo):
    """the sum rules printed in the headers of Si2_epsilon.dat (polar) and Si2_loss_fn.dat (unpolar)"""
    import json
    G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "gold_odo.json")))
    fx = fixture("si2_optics")
    vol = float(fx["cell_volume"])
    r, e1, e2 = _optics(odo, "polar", "adaptive")
    lf, _ = odo.calc_loss_fn(e1[:, 0], e2[:, 0], r["E"])
    n1, _, _ = odo.optics_sum_rules(r["E"], e2[:, 0], lf, vol)
    assert abs(n1 / G["testopt_optics_polar"]["N_eff"] - 1.0) < 1e-12
    r, e1, e2 = _optics(odo, "unpolar", "linear")
    lf, _ = odo.calc_loss_fn(e1[:, 0], e2[:, 0], r["E"], lossfn_gaussian=1.0)
    _, n2, n3 = odo.optics_sum_rules(r["E"], e2[:, 0], lf, vol)
    assert abs(n2 / G["testopt_optics_unpolar"]["N_eff2"] - 1.0) < 1e-11
    assert abs(n3 / G["testopt_optics_unpolar"]["N_eff3"] - 1.0) < 1e-11


def test_optics_tensor_epsilon(odo):
    r, e1, e2 = _optics(odo, "tensor", "adaptive")
    G = gold("testopt_optics_tensor")
    for c in range(6):
        g = G[f"block{c}"]
        assert rel_to_peak(e2[:, c], g[:, 2]) < LIST17, c
        assert rel_to_peak(e1[:, c], g[:, 1]) < 1e-12, c


def _core(odo, fxname, btype, core_type, polar, dos_spacing, adaptive_smearing=0.4):
    fx = fixture(fxname)
    s = oracle_sys(odo, fx)
    br = odo.make_broadening(adaptive_smearing=adaptive_smearing)
    first = odo.dos_utils_calculate(s, btype, dos_spacing=dos_spacing, br=br)     # dos_utils_set_efermi
    ef = first["efermi"]
    nsym = fx["symops"].shape[2]
    symops = fx["symops"] if nsym > 1 else None   # core.f90:106 errors for >1; 1 op == identity
    mw = odo.core_prepare_matrix_elements(s, fx["elnes_mat"], core_type, ef, polar=polar, qdir=(0.5, 0.5, 0.25),
                                          symops=fx["symops"] if polar else symops)
    # second setup_energy_scale with the now-sticky dos_nbins (quirk Q4)
    E, delta, n = odo.dos_energy_scale(s, dos_spacing, dos_nbins=first["nbins"])
    dos, intdos, wdos = odo.calculate_dos(btype[0], s, br, E, delta, mw)
    # core.f90:317-323
    eps0, e_charge = 8.8541878176E-12, 1.602176487E-19
    c = (e_charge * odo.C.c_double(3.141592653589793238462643383279502884197).value * float(np.float32(1e-20))) / (
        s.cell_volume * float(np.float32(1e-30)) * eps0)
    spec = wdos * 0.52917720859 ** 2
    spec = spec * c
    return fx, E, spec


def _edges(fx):
    """core.f90:366-399: consecutive orbitals with equal (species, rank, shell, am) form one edge."""
    # CASTEP channel numbering 1..16 -> l  (elec_elnes_find_channel_names, electronic.f90:995-1062)
    ch2l = {1: 0, 2: 1, 3: 1, 4: 1, 5: 2, 6: 2, 7: 2, 8: 2, 9: 2}
    ch2l.update({c: 3 for c in range(10, 17)})
    am = [ch2l[int(c)] for c in fx["elnes_am"]]
    keys = list(zip(fx["elnes_species"], fx["elnes_rank"], fx["elnes_shell"], am))
    edges = []
    for o, k in enumerate(keys):
        if edges and edges[-1][0] == k:
            edges[-1][1].append(o)
        else:
            edges.append((k, [o]))
    return [e[1] for e in edges]


# LAI broadening of each test's .odi: (LAI_gaussian_width, LAI_lorentzian_width, LAI_lorentzian_scale)
LAI = {"testopt_core_polarised": (0.0, 0.2, 0.1), "testopt_core_emisson": (0.5, 0.0, 0.0), "testopt_core_all": (0.8, 0.2, 0.1)}


@pytest.mark.parametrize("test,btype,core_type,polar", [
    ("testopt_core_polarised", "adaptive", "emission", True),
    ("testopt_core_emisson", "linear", "emission", True),
    ("testopt_core_all", "linear", "all", False),
])
def test_core_si2(odo, test, btype, core_type, polar):
    fx, E, spec = _core(odo, "si2_core", btype, core_type, polar, 0.01)
    ef = odo.dos_utils_calculate(oracle_sys(odo, fx), btype, dos_spacing=0.01)["efermi"]
    gw, lw, ls = LAI[test]
    lai = odo.core_lai_broadening(spec, E, ef, gaussian_width=gw, lorentzian_width=lw, lorentzian_scale=ls)   # core.f90:70-76
    G = gold(test)
    edges = _edges(fx)
    assert len(edges) == len(G) == 6
    for i, orbs in enumerate(edges):
        g = G[f"block{i}"]
        assert g.shape[0] == E.size
        assert np.abs(E - g[:, 0]).max() < 1e-11
        col = np.zeros(E.size)
        for o in orbs:
            col = col + spec[:, 0, o] / float(len(orbs))
        assert rel_to_peak(col, g[:, 1]) < LIST17, (test, i)
        colb = np.zeros(E.size)
        for o in orbs:
            colb = colb + lai[:, 0, o] / float(len(orbs))
        assert rel_to_peak(colb, g[:, 2]) < 1e-12, (test, i, "LAI")


def test_core_chemical_shift_nage(odo):
    """testopt_core_chemical_shift: NaGe (2x2x2 mesh, a non-symmetric reciprocal matrix -- the one fixture of the reference
    whose lattice is not symmetric), adaptive_smearing 0.2, absorption, LAI 1.0 / 0.5 / 0.1 and the Mizoguchi shift.
    core_calculate never calls dos_utils_compute_bandgap, so write_core's cbm_energy is the module's initial 0
    (core.f90:38-83, :533-535; dos_utils.f90:56): E_shift = E + core_chemical_shift."""
    fx, E, spec = _core(odo, "nage_core", "adaptive", "absorption", False, 0.1, adaptive_smearing=0.2)
    rl = fx["recip_lattice"]
    assert np.abs(rl - rl.T).max() > 1e-3 * np.abs(rl).max()
    s = oracle_sys(odo, fx)
    ef = odo.dos_utils_calculate(s, "adaptive", dos_spacing=0.1, br=odo.make_broadening(adaptive_smearing=0.2))["efermi"]
    lai = odo.core_lai_broadening(spec, E, ef, gaussian_width=1.0, lorentzian_width=0.5, lorentzian_scale=0.1)
    G = gold("testopt_core_chemical_shift")
    edges = _edges(fx)
    assert len(edges) == len(G) == 17
    for i, orbs in enumerate(edges):
        g = G[f"block{i}"]
        assert g.shape[0] == E.size
        assert np.abs(E + 11142.275222 - g[:, 0]).max() < 1e-9
        col = sum(spec[:, 0, o] for o in orbs) / float(len(orbs))
        colb = sum(lai[:, 0, o] for o in orbs) / float(len(orbs))
        # 2.7e-11 of peak at worst (the Si2 core goldens: 5e-13), largest on the steep onset of each edge: the size an
        # 11-digit formatted fixture (es23.10 gradients and matrix elements) leaves if the golden was written from the
        # binary originals; with the TRANSPOSED reciprocal lattice the same comparison gives 4e-3
        assert rel_to_peak(col, g[:, 1]) < 5e-11, i
        assert rel_to_peak(colb, g[:, 2]) < 5e-11, i
    st = odo.Sys(s.band_energy, s.kpoint_weight, rl.T, s.kpoint_grid_dim, s.band_gradient, electrons_per_state=s.electrons_per_state,
                 num_electrons=s.num_electrons, cell_volume=s.cell_volume)
    mw = odo.core_prepare_matrix_elements(st, fx["elnes_mat"], "absorption", ef)
    wd = odo.calculate_dos("a", st, odo.make_broadening(adaptive_smearing=0.2), E, E[1] - E[0], mw)[2]
    wd0 = odo.calculate_dos("a", s, odo.make_broadening(adaptive_smearing=0.2), E, E[1] - E[0], mw)[2]
    assert rel_to_peak(wd, wd0) > 1e-3     # the lattice orientation is visible in this golden


def test_oracle_threads_match_serial(odo):
    """the multi-threaded helper of the oracle (k-points split like algor_dist_array) against the serial loop,
    with a k-point count that leaves a remainder > 1"""
    fx = fixture("si2_core")          # 14 k-points
    s = oracle_sys(odo, fx)
    E, delta, n = odo.dos_energy_scale(s, 0.05)
    for t in ("a", "l"):
        a = odo.calculate_dos(t, s, odo.make_broadening(), E, delta)
        b = odo.calculate_dos(t, s, odo.make_broadening(), E, delta, nthreads=4)     # 14 = 4 + 4 + 3 + 3
        assert rel_to_peak(b[0], a[0]) < 1e-13 and rel_to_peak(b[1], a[1]) < 1e-13      # summation order only
