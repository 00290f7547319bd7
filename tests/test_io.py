"""Input decoders of the C ABI (include/optados_b200.h, "input decoders").

Two kinds of checks, all on the CPU:
  * files written HERE from the committed fixtures (tests/golden/fx_*.npz) in the layouts the reference's writers
    use (CASTEP .bands text; od2od's *_fmt text; Fortran sequential unformatted *_bin and legacy files, big- and
    little-endian) are decoded by the library and compared with the arrays they were written from;
  * when the reference checkout is mounted (this container only), its own fixture files are decoded and compared
    with the committed fixtures, which tests/golden/make_golden.py extracted with independent Python readers.
"""
import bz2
import os
import struct

import numpy as np
import pytest

import optados_b200 as ob
from fxutil import fixture

H2eV = 27.21138342902473
bohr2ang = 0.52917720859
REF = "/root/reference/optados/test-suite/checkpoints"


# ------------------------------------------------------------------------------------------------ writers
def write_bands(path, fx):
    nk, ns, nb = int(fx["nk"]), int(fx["ns"]), int(fx["nb"])
    with open(path, "w") as f:
        f.write(f"Number of k-points {nk:5d}\n")
        f.write(f"Number of spin components {ns}\n")
        f.write("Number of electrons " + " ".join(f"{x:8.4g}" for x in fx["num_electrons"][:ns]) + "\n")
        f.write("Number of eigenvalues " + " ".join(f"{nb:6d}" for _ in range(ns)) + "\n")
        ef = float(fx["efermi_castep"]) / H2eV
        f.write(("Fermi energies (in atomic units) " if ns > 1 else "Fermi energy (in atomic units) ") +
                " ".join(f"{ef:12.6f}" for _ in range(ns)) + "\n")
        f.write("Unit cell vectors\n")
        for i in range(3):
            f.write(" ".join(f"{x:25.17e}" for x in fx["real_lattice_bohr"][:, i]) + "\n")
        for ik in range(nk):
            k = fx["kpoint_r"][:, ik]
            f.write(f"K-point {ik + 1:5d} {k[0]:.17e} {k[1]:.17e} {k[2]:.17e} {fx['kpoint_weight'][ik]:.17e}\n")
            for s in range(ns):
                f.write(f"Spin component {s + 1}\n")
                for ib in range(nb):
                    f.write(f"{fx['band_energy'][ib, s, ik] / H2eV:25.17e}\n")


class FortranWriter:
    """Fortran sequential unformatted records with 4-byte markers"""

    def __init__(self, path, big_endian):
        self.f = open(path, "wb")
        self.e = ">" if big_endian else "<"

    def rec(self, payload):
        m = struct.pack(self.e + "i", len(payload))
        self.f.write(m + payload + m)

    def reals(self, a):
        self.rec(np.asarray(a, dtype=np.float64).ravel(order="F").astype(self.e + "f8").tobytes())

    def ints(self, a):
        self.rec(np.asarray(a, dtype=np.int32).ravel(order="F").astype(self.e + "i4").tobytes())

    def cplx(self, a):
        a = np.asarray(a, dtype=np.complex128).ravel(order="F")
        v = np.empty(2 * a.size)
        v[0::2], v[1::2] = a.real, a.imag
        self.rec(v.astype(self.e + "f8").tobytes())

    def header(self):
        self.reals([1.0])
        self.rec(b"written by tests/test_io.py".ljust(80))

    def close(self):
        self.f.close()


def fmt_reals(v, per_line=6):
    v = np.asarray(v).ravel()
    return "\n".join(" ".join(f"{x:23.16E}" for x in v[i:i + per_line]) for i in range(0, v.size, per_line)) + "\n"


def interleave(c):
    c = np.asarray(c).ravel(order="F")
    v = np.empty(2 * c.size)
    v[0::2], v[1::2] = c.real, c.imag
    return v


# ------------------------------------------------------------------------------------------------ .bands, cell
@pytest.mark.parametrize("name", ["si2_optics", "si2_core", "nage_core"])
def test_bands_round_trip(tmp_path, name):
    fx = fixture(name)
    p = tmp_path / "x.bands"
    write_bands(p, fx)
    b = ob.read_bands(p)
    assert (b["nk"], b["ns"], b["nb"]) == (int(fx["nk"]), int(fx["ns"]), int(fx["nb"]))
    np.testing.assert_allclose(b["band_energy"], fx["band_energy"], rtol=2e-16, atol=0)
    np.testing.assert_allclose(b["kpoint_r"], fx["kpoint_r"], rtol=0, atol=0)
    np.testing.assert_allclose(b["kpoint_weight"], fx["kpoint_weight"], rtol=0, atol=0)
    np.testing.assert_allclose(b["real_lattice_bohr"], fx["real_lattice_bohr"], rtol=0, atol=0)
    assert b["electrons_per_state"] == float(fx["electrons_per_state"])
    assert abs(b["efermi_castep"] - float(fx["efermi_castep"])) < 1e-4 * H2eV   # printed with 4 decimals (f12.4)
    rl, rc, vol = ob.cell_calc_lattice(b["real_lattice_bohr"])
    np.testing.assert_array_equal(rl, fx["real_lattice"])
    np.testing.assert_array_equal(rc, fx["recip_lattice"])
    assert vol == float(fx["cell_volume"])
    np.testing.assert_array_equal(ob.cell_find_mp_grid(b["kpoint_r"]), fx["kpoint_grid_dim"])


def test_mp_grid_shapes():
    # full and symmetry-reduced Monkhorst-Pack sets, shifted and unshifted
    for n in [(1, 1, 1), (2, 2, 2), (3, 3, 3), (4, 4, 2), (5, 4, 3), (10, 10, 10)]:
        ks = np.array([[(2 * i - n[0] + 1) / (2 * n[0]), (2 * j - n[1] + 1) / (2 * n[1]), (2 * k - n[2] + 1) / (2 * n[2])]
                       for i in range(n[0]) for j in range(n[1]) for k in range(n[2])]).T
        np.testing.assert_array_equal(ob.cell_find_mp_grid(ks), n)
        half = ks[:, : max(1, ks.shape[1] // 2 + 1)]            # k and -k are folded together (time reversal)
        np.testing.assert_array_equal(ob.cell_find_mp_grid(half), n)


def test_symmetry_ops(tmp_path):
    fx = fixture("si2_optics")
    ops = fx["symops"]
    p = tmp_path / "x-out.cell"
    with open(p, "w") as f:
        f.write("%BLOCK LATTICE_CART\n 1 0 0\n 0 1 0\n 0 0 1\n%ENDBLOCK LATTICE_CART\n\n# a comment\n%BLOCK SYMMETRY_OPS\n")
        for n in range(ops.shape[2]):
            f.write(f"# Symm. op. {n + 1}\n")
            for r in range(3):
                f.write(" ".join(f"{x:25.17e}" for x in ops[:, r, n]) + "\n")
            f.write("  0.0 0.0 0.0   ! displacement\n")
        f.write("%ENDBLOCK SYMMETRY_OPS\n")
    got = ob.read_symmetry_ops(p)
    np.testing.assert_array_equal(got, ops)
    q = tmp_path / "none.cell"
    q.write_text("%block lattice_cart\n1 0 0\n0 1 0\n0 0 1\n%endblock lattice_cart\n")
    assert ob.read_symmetry_ops(q).shape[2] == 0


# ------------------------------------------------------------------------------------------------ matrix elements
@pytest.mark.parametrize("big", [True, False])
def test_ome_bin_legacy_fmt(tmp_path, big):
    fx = fixture("si2_optics")
    om = fx["optical_mat"]
    nb, _, _, nk, ns = om.shape
    raw = om / (bohr2ang * H2eV)      # what CASTEP writes (Hartree Bohr)
    w = FortranWriter(tmp_path / "x.ome_bin", big)
    w.header()
    for ik in range(nk):
        for s in range(ns):
            w.cplx(raw[:, :, :, ik, s])
    w.close()
    got = ob.read_ome(tmp_path / "x.ome_bin", "b", nb, nk, ns)
    expect = (raw.real * bohr2ang * H2eV) + 1j * (raw.imag * bohr2ang * H2eV)
    np.testing.assert_array_equal(got, expect)
    np.testing.assert_allclose(got, om, rtol=4e-16, atol=0)
    # legacy .cst_ome: one complex per record, Ha Bohr^2 / Ang (electronic.f90:370-376, :423-425)
    sub = raw[:5, :5, :, :2, :]
    w = FortranWriter(tmp_path / "x.cst_ome", big)
    for ik in range(2):
        for s in range(ns):
            for i in range(3):
                for jb in range(5):
                    for ib in range(5):
                        w.cplx([sub[ib, jb, i, ik, s]])
    w.close()
    got = ob.read_ome(tmp_path / "x.cst_ome", "l", 5, 2, ns)
    np.testing.assert_array_equal(got.real, sub.real * bohr2ang * bohr2ang * H2eV)
    np.testing.assert_array_equal(got.imag, sub.imag * bohr2ang * bohr2ang * H2eV)
    # od2od's formatted twin
    with open(tmp_path / "x.ome_fmt", "w") as f:
        f.write(f"{1.0:23.10E}\n" + "header".ljust(80) + "\n")
        for ik in range(nk):
            for s in range(ns):
                f.write(fmt_reals(interleave(raw[:, :, :, ik, s])))
    got = ob.read_ome(tmp_path / "x.ome_fmt", "t", nb, nk, ns)
    np.testing.assert_allclose(got, om, rtol=1e-15, atol=1e-300)


@pytest.mark.parametrize("big", [True, False])
def test_dome(tmp_path, big):
    fx = fixture("si2_dos")
    g = fx["band_gradient"]
    nb, _, nk, ns = g.shape
    raw = g / (bohr2ang * H2eV)
    for fmt, name, hdr in (("b", "x.dome_bin", True), ("l", "x.cst_vel", False)):
        w = FortranWriter(tmp_path / name, big)
        if hdr:
            w.header()
        for ik in range(nk):
            for s in range(ns):
                w.reals(raw[:, :, ik, s])
        w.close()
        got = ob.read_dome(tmp_path / name, fmt, nb, nk, ns)
        np.testing.assert_array_equal(got, raw * bohr2ang * H2eV)
    with open(tmp_path / "x.dome_fmt", "w") as f:
        f.write(f"{1.0:23.10E}\n" + "header".ljust(80) + "\n")
        for ik in range(nk):
            for s in range(ns):
                f.write(fmt_reals(raw[:, :, ik, s].ravel(order="F"), per_line=3 * nb))
    np.testing.assert_allclose(ob.read_dome(tmp_path / "x.dome_fmt", "t", nb, nk, ns), g, rtol=1e-15, atol=1e-300)


@pytest.mark.parametrize("big", [True, False])
def test_pdos(tmp_path, big):
    fx = fixture("si2_optics")
    wts, nocc = fx["pdos_weights"], fx["nbands_occ"]
    norb, nb, nk, ns = wts.shape
    for fmt, name, hdr in (("b", "x.pdos_bin", True), ("l", "x.pdos_weights", False)):
        w = FortranWriter(tmp_path / name, big)
        if hdr:
            w.header()
        for v in (nk, ns, norb, nb):
            w.ints([v])
        for a in (fx["pdos_species"], fx["pdos_rank"], fx["pdos_am"]):
            w.ints(a)
        for ik in range(nk):
            w.rec(struct.pack(w.e + "i3d", ik + 1, *fx["kpoint_r"][:, ik]))
            for s in range(ns):
                w.ints([s + 1])
                w.ints([nocc[ik, s]])
                for ib in range(nocc[ik, s]):
                    w.reals(wts[:, ib, ik, s])
        w.close()
        d = ob.read_pdos(tmp_path / name, fmt)
        np.testing.assert_array_equal(d["pdos_weights"], wts)
        np.testing.assert_array_equal(d["nbands_occ"], nocc)
        for k in ("pdos_species", "pdos_rank", "pdos_am"):
            np.testing.assert_array_equal(d[k], fx[k])
    with open(tmp_path / "x.pdos_fmt", "w") as f:
        f.write(f"{1.0:23.10E}\n" + "header".ljust(80) + "\n")
        for cap, v in (("Kpoints", nk), ("Spins", ns), ("Orbials", norb), ("Bands", nb)):
            f.write(f"{cap:<10s}{v:6d}\n")
        for cap, a in (("Species number", fx["pdos_species"]), ("Rank in species", fx["pdos_rank"]), ("AM channel", fx["pdos_am"])):
            f.write(cap + "\n" + "".join(f" {int(x):5d}" for x in a) + "\n")
        for ik in range(nk):
            f.write(f"{ik + 1:6d}" + "".join(f"{x:23.16E}" for x in fx["kpoint_r"][:, ik]) + "\n")
            for s in range(ns):
                f.write(f"{s + 1:6d}\n{int(nocc[ik, s]):6d}\n")
                for ib in range(nocc[ik, s]):
                    f.write("".join(f" {x:23.16E}" for x in wts[:, ib, ik, s]) + "\n")
    d = ob.read_pdos(tmp_path / "x.pdos_fmt", "t")
    np.testing.assert_allclose(d["pdos_weights"], wts, rtol=1e-15, atol=1e-300)
    np.testing.assert_array_equal(d["nbands_occ"], nocc)
    np.testing.assert_array_equal(d["pdos_am"], fx["pdos_am"])


@pytest.mark.parametrize("big", [True, False])
def test_elnes(tmp_path, big):
    fx = fixture("si2_core")
    m = fx["elnes_mat"]
    norb, nb, _, nk, ns = m.shape
    meta = [fx["elnes_species"], fx["elnes_rank"], fx["elnes_shell"], fx["elnes_am"]]
    w = FortranWriter(tmp_path / "x.elnes_bin", big)
    w.header()
    for v in (norb, nb, nk, ns):
        w.ints([v])
    for a in meta:
        w.ints(a)
    for ik in range(nk):
        for s in range(ns):
            w.cplx(m[:, :, :, ik, s])
    w.close()
    d = ob.read_elnes(tmp_path / "x.elnes_bin", "b")
    np.testing.assert_array_equal(d["elnes_mat"], m)
    for k, a in zip(("elnes_species", "elnes_rank", "elnes_shell", "elnes_am"), meta):
        np.testing.assert_array_equal(d[k], a)
    # legacy .eels_mat: one record of the three cartesian components per (orbital, band)
    w = FortranWriter(tmp_path / "x.eels_mat", big)
    for v in (norb, nb, nk, ns):
        w.ints([v])
    for a in meta:
        w.ints(a)
    for ik in range(nk):
        for s in range(ns):
            for o in range(norb):
                for ib in range(nb):
                    w.cplx(m[o, ib, :, ik, s])
    w.close()
    np.testing.assert_array_equal(ob.read_elnes(tmp_path / "x.eels_mat", "l")["elnes_mat"], m)
    with open(tmp_path / "x.elnes_fmt", "w") as f:
        f.write(f"{1.0:23.10E}\n" + "header".ljust(80) + "\n")
        for cap, v in (("Norbitals", norb), ("Nbands", nb), ("Nkpoints", nk), ("Nspins", ns)):
            f.write(f"{cap:<20s} {v:5d}\n")
        for cap, a in zip(("Species_no", "Rank", "Shell", "Am_channel"), meta):
            f.write(f"{cap:<10s}" + "".join(f" {int(x):5d}" for x in a) + "\n")
        for ik in range(nk):
            for s in range(ns):
                f.write(fmt_reals(interleave(m[:, :, :, ik, s])))
    d = ob.read_elnes(tmp_path / "x.elnes_fmt", "t")
    np.testing.assert_allclose(d["elnes_mat"], m, rtol=1e-15, atol=1e-300)
    np.testing.assert_array_equal(d["elnes_shell"], fx["elnes_shell"])


def test_errors(tmp_path):
    with pytest.raises(ob.OptadosB200Error, match="Problem opening"):
        ob.read_bands(tmp_path / "missing.bands")
    p = tmp_path / "short.dome_bin"
    w = FortranWriter(p, True)
    w.header()
    w.reals(np.zeros(6))
    w.close()
    with pytest.raises(ob.OptadosB200Error, match="short record"):
        ob.read_dome(p, "b", 4, 1, 1)         # record holds 6 reals, 12 expected
    w = FortranWriter(p, True)
    w.reals([2.0])                            # a newer file version (electronic.f90:357-358)
    w.rec(b"h".ljust(80))
    w.close()
    with pytest.raises(ob.OptadosB200Error, match="newer version"):
        ob.read_dome(p, "b", 4, 1, 1)
    with pytest.raises(ob.OptadosB200Error):
        ob.read_ome(p, "x", 1, 1, 1)


# ------------------------------------------------------------------------------------------------ the reference's files
def _unbz(src, dst):
    with open(dst, "wb") as f:
        f.write(bz2.open(src, "rb").read())
    return dst


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not mounted (only in the build container)")
@pytest.mark.parametrize("setname,seed,name", [("Si2_optics", "Si2", "si2_optics"), ("Si2_DOS", "Si2", "si2_dos"),
                                               ("Si2_core", "Si2", "si2_core"), ("NaGe_core", "NaGe", "nage_core")])
def test_reference_fixture_files(tmp_path, setname, seed, name):
    d = f"{REF}/{setname}"
    if not os.path.exists(f"{d}/{seed}.bands.bz2"):
        pytest.skip("fixture set not present")
    fx = fixture(name)
    b = ob.read_bands(_unbz(f"{d}/{seed}.bands.bz2", tmp_path / "s.bands"))
    np.testing.assert_array_equal(b["band_energy"], fx["band_energy"])
    np.testing.assert_array_equal(b["kpoint_weight"], fx["kpoint_weight"])
    np.testing.assert_array_equal(b["num_electrons"], fx["num_electrons"])
    assert b["efermi_castep"] == float(fx["efermi_castep"])
    np.testing.assert_array_equal(ob.cell_find_mp_grid(b["kpoint_r"]), fx["kpoint_grid_dim"])
    np.testing.assert_array_equal(ob.cell_calc_lattice(b["real_lattice_bohr"])[1], fx["recip_lattice"])
    nb, ns, nk = b["nb"], b["ns"], b["nk"]
    if "optical_mat" in fx and os.path.exists(f"{d}/{seed}.ome_fmt.bz2"):
        om = ob.read_ome(_unbz(f"{d}/{seed}.ome_fmt.bz2", tmp_path / "s.ome_fmt"), "t", nb, nk, ns)
        np.testing.assert_array_equal(om, fx["optical_mat"])
    if "band_gradient" in fx and os.path.exists(f"{d}/{seed}.dome_fmt.bz2"):
        g = ob.read_dome(_unbz(f"{d}/{seed}.dome_fmt.bz2", tmp_path / "s.dome_fmt"), "t", nb, nk, ns)
        np.testing.assert_array_equal(g, fx["band_gradient"])
    if "pdos_weights" in fx and os.path.exists(f"{d}/{seed}.pdos_fmt.bz2"):
        p = ob.read_pdos(_unbz(f"{d}/{seed}.pdos_fmt.bz2", tmp_path / "s.pdos_fmt"), "t")
        np.testing.assert_array_equal(p["pdos_weights"], fx["pdos_weights"])
        np.testing.assert_array_equal(p["nbands_occ"], fx["nbands_occ"])
        np.testing.assert_array_equal(p["pdos_am"], fx["pdos_am"])
    if "elnes_mat" in fx and os.path.exists(f"{d}/{seed}.elnes_fmt.bz2"):
        e = ob.read_elnes(_unbz(f"{d}/{seed}.elnes_fmt.bz2", tmp_path / "s.elnes_fmt"), "t")
        np.testing.assert_array_equal(e["elnes_mat"], fx["elnes_mat"])
        np.testing.assert_array_equal(e["elnes_am"], fx["elnes_am"])
    if "symops" in fx and os.path.exists(f"{d}/{seed}-out.cell"):
        np.testing.assert_array_equal(ob.read_symmetry_ops(f"{d}/{seed}-out.cell"), fx["symops"])
