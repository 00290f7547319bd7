#!/usr/bin/env python3
"""Extract the reference's own test fixtures and golden outputs into small .npz files.

Run HERE (the container that mounts /root/reference); the outputs are committed under
tests/golden/ because /root/reference does not exist on the GPU box.

Inputs  : /root/reference/optados/test-suite/checkpoints/<set>/*.bz2 (CASTEP-format text fixtures)
          /root/reference/optados/test-suite/tests/<t>/benchmark.out.default.inp=<seed>.odi
Outputs : tests/golden/fx_<set>.npz      -- arrays in the reference's in-memory layout and units
          tests/golden/gold_<test>.npz   -- the numeric columns of the reference's golden output

The readers below restate (reference paths relative to optados/src/):
  .bands            electronic.f90:490-601   (Hartree -> eV at :601)
  lattice           cell.f90:924-980         (Bohr -> Ang, reciprocal incl. 2 pi)
  MP grid           cell.f90:87-318          (cell_find_MP_grid)
  .ome_fmt          od2od.f90:110-158, then the fmt->bin->read unit round trip
                    od2od.f90:152,229 + electronic.f90:427
  .dome_fmt         od2od.f90:254-295 (+ same round trip, electronic.f90:271)
  .pdos_fmt         od2od.f90:468-560
  .elnes_fmt        od2od.f90:639-706
  -out.cell symops  cell.f90:622-690
"""
import bz2
import json
import os
import re
import sys

import numpy as np

REF = "/root/reference/optados/test-suite"
OUT = os.path.dirname(os.path.abspath(__file__))

H2eV = 27.21138342902473
bohr2ang = 0.52917720859
pi = 3.141592653589793238462643383279502884197


def rd(path):
    if path.endswith(".bz2"):
        return bz2.open(path, "rt").read()
    return open(path).read()


def read_bands(path):
    L = rd(path).splitlines()
    nk = int(L[0].split("k-points")[1])
    ns = int(L[1].split("components")[1])
    nel = [float(x) for x in L[2].split("electrons")[1].split()][:ns]
    nb = int(L[3].split("eigenvalues")[1].split()[0])
    # electronic.f90:511-512 reads 12 chars after 'units)' as f12.4
    pos = L[4].index("units)") + 6
    efermi_castep = float(L[4][pos:pos + 12])
    real_lattice = np.zeros((3, 3))  # real_lattice(:, i) = line i
    for i in range(3):
        real_lattice[:, i] = [float(x) for x in L[6 + i].split()]
    band_energy = np.zeros((nb, ns, nk), order="F")
    kpoint_r = np.zeros((3, nk), order="F")
    kweight = np.zeros(nk)
    p = 9
    for ik in range(nk):
        t = L[p].split("K-point")[1].split()
        kpoint_r[:, ik] = [float(t[1]), float(t[2]), float(t[3])]
        kweight[ik] = float(t[4])
        p += 1
        for is_ in range(ns):
            p += 1  # 'Spin component'
            for ib in range(nb):
                band_energy[ib, is_, ik] = float(L[p])
                p += 1
    band_energy = band_energy * H2eV
    efermi_castep = efermi_castep * H2eV
    return dict(nk=nk, ns=ns, nb=nb, num_electrons=np.array(nel), efermi_castep=efermi_castep,
                real_lattice_bohr=real_lattice, band_energy=band_energy, kpoint_r=kpoint_r,
                kpoint_weight=kweight, electrons_per_state=2.0 if ns < 2 else 1.0)


def calc_lattice(real_lattice_bohr):
    """cell.f90:924-980"""
    rl = real_lattice_bohr * bohr2ang
    rc = np.zeros((3, 3))
    rc[0, 0] = rl[1, 1] * rl[2, 2] - rl[2, 1] * rl[1, 2]
    rc[1, 0] = rl[1, 2] * rl[2, 0] - rl[2, 2] * rl[1, 0]
    rc[2, 0] = rl[1, 0] * rl[2, 1] - rl[2, 0] * rl[1, 1]
    rc[0, 1] = rl[2, 1] * rl[0, 2] - rl[0, 1] * rl[2, 2]
    rc[1, 1] = rl[2, 2] * rl[0, 0] - rl[0, 2] * rl[2, 0]
    rc[2, 1] = rl[2, 0] * rl[0, 1] - rl[0, 0] * rl[2, 1]
    rc[0, 2] = rl[0, 1] * rl[1, 2] - rl[1, 1] * rl[0, 2]
    rc[1, 2] = rl[0, 2] * rl[1, 0] - rl[1, 2] * rl[0, 0]
    rc[2, 2] = rl[0, 0] * rl[1, 1] - rl[1, 0] * rl[0, 1]
    vol = rl[0, 0] * rc[0, 0] + rl[1, 0] * rc[0, 1] + rl[2, 0] * rc[0, 2]
    if vol < 0:
        vol = -vol
    rc = rc * pi * 2.0 / vol
    return rl, rc, vol


def find_mp_grid(kpoints):
    """cell.f90:87-318 (kpoints: (3, nk))"""
    nk = kpoints.shape[1]
    ktr = np.concatenate([kpoints, -kpoints], axis=1)
    ktr = ktr - np.floor(ktr + 0.5)
    sub_tol = np.finfo(float).eps
    tol = 0.000001
    grid = [0, 0, 0]
    for d in range(3):
        uniq = [ktr[d, 0]]
        for ik in range(1, 2 * nk):
            if any(abs(u - ktr[d, ik]) <= sub_tol for u in uniq):
                continue
            uniq.append(ktr[d, ik])
        n = len(uniq)
        if n == 1:
            grid[d] = 1
            continue
        if n == 2:
            mi = abs(uniq[0] - uniq[1])
            grid[d] = 2 if abs(mi - 0.5) <= tol else 1
            continue
        big = np.finfo(float).max
        m1 = big
        for i in range(n):
            for j in range(i + 1, n):
                m1 = min(m1, abs(uniq[i] - uniq[j]))
        m2 = big
        for i in range(n):
            for j in range(i, n):
                im = abs(uniq[i] - uniq[j])
                if im < m2 and im > m1 + tol:
                    m2 = im
        m3 = big
        for i in range(n):
            for j in range(i, n):
                im = abs(uniq[i] - uniq[j])
                if im < m3 and im > m2 + tol:
                    m3 = im
        if abs(2.0 * m1 - m2) < tol:
            grid[d] = int(1.0 / m1)
        else:
            grid[d] = int(1.0 / m3)
    return np.array(grid, dtype=np.int32)


def _numbers_after(text, nskip_lines):
    L = text.splitlines()
    return np.array(" ".join(L[nskip_lines:]).split(), dtype=np.float64)


def unit_round_trip(x):
    c = bohr2ang * H2eV
    v = x * c        # od2od read_*_fmt
    v = v / c        # od2od write_*_bin
    v = v * bohr2ang * H2eV  # electronic.f90:271 / :427 (left to right)
    return v


def read_ome_fmt(path, nb, nk, ns):
    v = _numbers_after(rd(path), 2)
    assert v.size == 2 * 3 * nb * nb * nk * ns, (v.size, nb, nk, ns)
    c = v[0::2] + 1j * v[1::2]
    # file order: ik, is, then (ib fastest, jb, i)
    c = c.reshape((nk, ns, 3, nb, nb))          # [ik, is, i, jb, ib]
    om = np.transpose(c, (4, 3, 2, 0, 1))       # (ib, jb, i, ik, is)
    om = np.asfortranarray(om)
    return unit_round_trip(om.real) + 1j * unit_round_trip(om.imag)


def read_dome_fmt(path, nb, nk, ns):
    v = _numbers_after(rd(path), 2)
    assert v.size == 3 * nb * nk * ns
    g = v.reshape((nk, ns, 3, nb))              # [ik, is, i, ib]
    g = np.asfortranarray(np.transpose(g, (3, 2, 0, 1)))  # (ib, i, ik, is)
    return unit_round_trip(g)


def read_pdos_fmt(path):
    L = rd(path).splitlines()
    nk = int(L[2][10:16]); ns = int(L[3][10:16]); norb = int(L[4][10:16]); nb = int(L[5][10:16])
    species = np.array(L[7].split(), dtype=np.int32)
    rank = np.array(L[9].split(), dtype=np.int32)
    am = np.array(L[11].split(), dtype=np.int32)
    w = np.zeros((norb, nb, nk, ns), order="F")
    nbocc = np.zeros((nk, ns), dtype=np.int32)
    p = 12
    for ik in range(nk):
        p += 1
        for is_ in range(ns):
            p += 1
            nbocc[ik, is_] = int(L[p]); p += 1
            for ib in range(nbocc[ik, is_]):
                w[:, ib, ik, is_] = [float(x) for x in L[p].split()]
                p += 1
    return dict(pdos_weights=w, pdos_species=species, pdos_rank=rank, pdos_am=am, nbands_occ=nbocc)


def read_elnes_fmt(path):
    L = rd(path).splitlines()
    norb = int(L[2][21:26]); nb = int(L[3][21:26]); nk = int(L[4][21:26]); ns = int(L[5][21:26])
    species = np.array(L[6][10:].split(), dtype=np.int32)
    rank = np.array(L[7][10:].split(), dtype=np.int32)
    shell = np.array(L[8][10:].split(), dtype=np.int32)
    am = np.array(L[9][10:].split(), dtype=np.int32)
    v = _numbers_after("\n".join(L), 10)
    assert v.size == 2 * norb * nb * 3 * nk * ns, (v.size, norb, nb, nk, ns)
    c = v[0::2] + 1j * v[1::2]
    c = c.reshape((nk, ns, 3, nb, norb))        # [ik, is, indx, ib, iorb]
    m = np.asfortranarray(np.transpose(c, (4, 3, 2, 0, 1)))  # (orb, ib, indx, ik, is)
    return dict(elnes_mat=m, elnes_species=species, elnes_rank=rank, elnes_shell=shell, elnes_am=am)


def read_symops(path):
    txt = rd(path).lower()
    L = [re.sub(r"[#!].*", "", l) for l in txt.splitlines()]
    s = next(i for i, l in enumerate(L) if "%block" in l and "symmetry_ops" in l)
    e = next(i for i, l in enumerate(L) if "%endblock" in l and "symmetry_ops" in l)
    rows = [l for l in L[s + 1:e] if l.strip()]
    n = len(rows) // 4
    ops = np.zeros((3, 3, n), order="F")
    for c in range(n):
        for r in range(3):
            ops[:, r, c] = [float(x) for x in rows[4 * c + r].split()]
    return ops


def fixture(setname, seed, want):
    d = f"{REF}/checkpoints/{setname}"
    b = read_bands(f"{d}/{seed}.bands.bz2")
    rl, rc, vol = calc_lattice(b["real_lattice_bohr"])
    out = dict(b)
    out.update(real_lattice=rl, recip_lattice=np.asfortranarray(rc), cell_volume=vol,
               kpoint_grid_dim=find_mp_grid(b["kpoint_r"]))
    nb, nk, ns = b["nb"], b["nk"], b["ns"]
    if "ome" in want:
        out["optical_mat"] = read_ome_fmt(f"{d}/{seed}.ome_fmt.bz2", nb, nk, ns)
    if "dome" in want:
        out["band_gradient"] = read_dome_fmt(f"{d}/{seed}.dome_fmt.bz2", nb, nk, ns)
    if "pdos" in want:
        out.update(read_pdos_fmt(f"{d}/{seed}.pdos_fmt.bz2"))
    if "elnes" in want:
        out.update(read_elnes_fmt(f"{d}/{seed}.elnes_fmt.bz2"))
    if "cell" in want:
        out["symops"] = read_symops(f"{d}/{seed}-out.cell")
    return out


def golden_table(test, seed):
    """All numeric rows of a golden .dat: returns list of 2-D blocks (blank/comment separated)."""
    path = f"{REF}/tests/{test}/benchmark.out.default.inp={seed}.odi"
    blocks, cur = [], []
    for line in open(path):
        s = line.strip()
        if not s or s.startswith("#") or s.startswith("&") or s.startswith("@"):
            if cur:
                blocks.append(np.array(cur)); cur = []
            continue
        try:
            cur.append([float(x) for x in s.split()])
        except ValueError:
            if cur:
                blocks.append(np.array(cur)); cur = []
    if cur:
        blocks.append(np.array(cur))
    return blocks


def golden_odo(test, seed):
    """every tagged number of a golden .odo (the `<- XXX` lines the reference's own test harness greps for):
    Fermi energies (EfA/EfF/EfL), band gaps (TBg, OBg, ABg, wAB: dos_utils.f90:457-750), band energies
    (BEA/BEF/BEL/BEC: :853-924), DOS at the Fermi level (DEA/DEF/DEL: :391-455) and the occupations (Occ: :841-846)"""
    path = f"{REF}/tests/{test}/benchmark.out.default.inp={seed}.odi"
    out = {}
    num = r"[-+]?\d+\.\d+(?:[Ee][-+]?\d+)?"
    for line in open(path):
        m = re.search(r"Fermi energy \((\w+) broadening\) :\s*([-\d.]+)", line)
        if m:
            out["efermi_" + m.group(1).lower()] = float(m.group(2))
        m = re.search(r"<- (\w+) \|", line)
        if not m:
            continue
        tag = m.group(1)
        body = line[:m.start()]
        if tag in ("OBg", "ABg"):
            mm = re.search(r"Spin :\s*(\d+)\s*:\s*(" + num + ")", body)
            out.setdefault(tag, []).append(float(mm.group(2)))
        elif tag in ("DEA", "DEF", "DEL"):
            mm = re.search(r"DOS at Fermi Energy :\s*(" + num + ")", body)
            out.setdefault(tag, []).append(float(mm.group(1)))
        elif tag == "Occ":
            mm = re.search(r"occupation between\s*(" + num + r")\s*and\s*(" + num + ")", body)
            out.setdefault(tag, []).append([float(mm.group(1)), float(mm.group(2))])
        elif tag in ("TBg", "wAB", "BEA", "BEF", "BEL", "BEC"):
            mm = re.findall(num, body)
            out[tag] = float(mm[-1])
    # multiplicities and the direct / indirect verdict printed beside the thermal gap
    txt = open(path).read()
    mm = re.findall(r"Spin :\s*(\d+)\s*:\s*(\d+)\s+(\d+)\s", txt)
    if mm:
        out["vbm_cbm_multiplicity"] = [[int(a[1]), int(a[2])] for a in mm]
    for key in ("Direct Gap", "Indirect Gap", "Gap unknown"):
        if key in txt:
            out["gap_kind"] = key
    return out


def golden_header_scalars(test, seed):
    """sum rules printed in the headers of Si2_epsilon.dat / Si2_loss_fn.dat (optics.f90:895, :1030-1031)"""
    path = f"{REF}/tests/{test}/benchmark.out.default.inp={seed}.odi"
    out = {}
    for line in open(path):
        if not line.lstrip().startswith("#"):
            break
        for key, pat in (("N_eff", r"Result of sum rule: Neff\(E\) =\s*([-+\dEe.]+)"),
                         ("N_eff2", r"Result of first sum rule: Neff\(E\) =\s*([-+\dEe.]+)"),
                         ("N_eff3", r"Result of second sum rule \(pi/2 = 1.570796327\):\s*([-+\dEe.]+)")):
            m = re.search(pat, line)
            if m:
                out[key] = float(m.group(1))
    return out


def copy_testsuite():
    """the INPUT files of the reference's test-suite (the .odi of every optados test, and the CASTEP-format checkpoint
    files they link to, still bz2-compressed) -> tests/golden/testsuite/, with index.json = test -> {seed, files, output,
    parser}.  These are data fixtures for the files-in / files-out tests of od_run_seed (tests/test_run_files.py)."""
    import configparser
    import shutil
    dst = os.path.join(OUT, "testsuite")
    os.makedirs(os.path.join(dst, "checkpoints"), exist_ok=True)
    cfg = configparser.ConfigParser()
    cfg.read(f"{REF}/tests/jobconfig")
    index = {}
    for test in sorted(os.listdir(f"{REF}/tests")):
        d = f"{REF}/tests/{test}"
        if not (test.startswith("testopt_") and os.path.isdir(d)):
            continue
        odis = [f for f in os.listdir(d) if f.endswith(".odi") and not f.startswith("benchmark")]
        if not odis:
            continue
        seed = odis[0][:-4]
        os.makedirs(os.path.join(dst, "tests", test), exist_ok=True)
        shutil.copyfile(f"{d}/{odis[0]}", os.path.join(dst, "tests", test, odis[0]))
        files, complete = {}, True
        for f in sorted(os.listdir(d)):
            if f.startswith("benchmark") or f in ("Makefile", ".gitignore") or f.endswith(".odi"):
                continue
            src = os.path.realpath(f"{d}/{f}")
            if not os.path.exists(src):
                complete = False          # a dangling link (Al4.ome_fmt.bz2 is not in the checkout)
                continue
            if "/checkpoints/" in src:
                rel = src.split("/checkpoints/")[1]
                os.makedirs(os.path.dirname(os.path.join(dst, "checkpoints", rel)), exist_ok=True)
                shutil.copyfile(src, os.path.join(dst, "checkpoints", rel))
                files[f] = "checkpoints/" + rel
            else:
                shutil.copyfile(src, os.path.join(dst, "tests", test, f))
                files[f] = f"tests/{test}/{f}"
        out = cfg[test]["output"].strip() if cfg.has_section(test) and "output" in cfg[test] else None
        prog = cfg[test]["program"].strip() if cfg.has_section(test) and "program" in cfg[test] else None
        index[test] = {"seed": seed, "files": files, "output": out, "program": prog, "complete": complete}
    with open(os.path.join(dst, "index.json"), "w") as f:
        json.dump(index, f, indent=1, sort_keys=True)
    print("testsuite inputs:", len(index), "tests")


def save(name, d):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in d.items()})
    print(f"{name}.npz  {os.path.getsize(path) / 1024:.1f} KiB")


def main():
    save("fx_si2_optics", fixture("Si2_optics", "Si2", {"ome", "pdos", "cell"}))
    save("fx_si2_optics_no_spin", fixture("Si2_optics_no_spin", "Si2", {"ome", "pdos", "cell"}))
    save("fx_si2_dos", fixture("Si2_DOS", "Si2", {"dome", "pdos", "cell"}))
    save("fx_si2_core", fixture("Si2_core", "Si2", {"dome", "elnes", "cell"}))
    save("fx_nage_core", fixture("NaGe_core_chemical_shift", "NaGe", {"dome", "elnes", "cell"}))
    # testopt_dos_fixed ships its own 10-k-point, 23-band Si2.bands (no gradients)
    b = read_bands(f"{REF}/tests/testopt_dos_fixed/Si2.bands")
    rl, rc, vol = calc_lattice(b["real_lattice_bohr"])
    b.update(real_lattice=rl, recip_lattice=np.asfortranarray(rc), cell_volume=vol, kpoint_grid_dim=find_mp_grid(b["kpoint_r"]))
    save("fx_si2_fixed", b)

    dat_tests = {
        "testopt_dos_adaptive_dat": "Si2", "testopt_dos_adaptive_no_spin": "Si2",
        "testopt_dos_linear_dome": "Si2", "testopt_jdos_adaptive_dome": "Si2",
        "testopt_jdos_linear_no_spin": "Si2", "testopt_pdos_angular": "Si2",
        "testopt_pdos_string2": "Si2", "testopt_pdos_string3": "Si2",
        "testopt_optics_polar": "Si2", "testopt_optics_tensor": "Si2", "testopt_optics_unpolar": "Si2",
        "testopt_core_polarised": "Si2", "testopt_core_emisson": "Si2", "testopt_core_all": "Si2",
        "testopt_core_chemical_shift": "NaGe",
    }
    # (testopt_pdos_string1 / 4, testopt_jdos_fixed_odo, testopt_optics_poly, testopt_core_absorb are .odo tests:
    #  their goldens are the tagged numbers of gold_odo.json)
    for t, seed in dat_tests.items():
        blocks = golden_table(t, seed)
        save("gold_" + t, {f"block{i}": b for i, b in enumerate(blocks)})
    odo = {}
    for t in ["testopt_dos_fixed", "testopt_dos_adaptive", "testopt_dos_linear", "testopt_jdos_fixed_odo",
              "testopt_pdos_string1", "testopt_pdos_string4", "testopt_optics_poly", "testopt_core_absorb"]:
        odo[t] = golden_odo(t, "Si2")
    for t in ["testopt_optics_polar", "testopt_optics_unpolar"]:
        odo[t] = golden_header_scalars(t, "Si2")
    with open(os.path.join(OUT, "gold_odo.json"), "w") as f:
        json.dump(odo, f, indent=1, sort_keys=True)
    print("gold_odo.json", odo)
    copy_testsuite()


if __name__ == "__main__":
    sys.exit(main())
