"""the reference's file contract.  Every optados test of the reference's test-suite whose
inputs are complete is run FILES IN -> FILES OUT through od_run_seed (the .odi parser, the CASTEP-format decoders, the
task drivers on the GPU, the writers) and the file its jobconfig names as `output` is compared with the committed
benchmark -- all numeric columns, to the print precision of the benchmark, not only the last line the reference's own
harness samples (test-suite/tools/parsers/parse_dos_dat.py).  Inputs: tests/golden/testsuite (copied by make_golden.py)."""
import bz2
import json
import os
import re
import shutil

import numpy as np
import pytest

from conftest import GOLDEN, load_npz

TS = os.path.join(GOLDEN, "testsuite")
INDEX = json.load(open(os.path.join(TS, "index.json")))
RUNNABLE = sorted(t for t, v in INDEX.items() if v["complete"] and v["program"] != "OPTADOS_MIZ_OUT_OK")


def stage(test, tmp_path):
    """what the test-suite Makefiles do, minus od2od: unpack the checkpoint files next to the .odi (the formatted *_fmt
    twins are read directly, through the same fmt -> bin -> read unit round trip)"""
    info = INDEX[test]
    d = tmp_path / test
    d.mkdir()
    shutil.copyfile(os.path.join(TS, "tests", test, info["seed"] + ".odi"), d / (info["seed"] + ".odi"))
    for name, rel in info["files"].items():
        src = os.path.join(TS, rel)
        if name.endswith(".bz2"):
            (d / name[:-4]).write_bytes(bz2.open(src).read())
        else:
            shutil.copyfile(src, d / name)
    return d, info["seed"]


def numeric_blocks(path):
    """the numeric rows of a .dat file as blocks (blank / comment separated), like make_golden.golden_table"""
    blocks, cur = [], []
    for line in open(path):
        s = line.strip()
        if not s or s.startswith("#") or s.startswith("&") or s.startswith("@"):
            if cur:
                blocks.append(np.array(cur))
                cur = []
            continue
        try:
            cur.append([float(x) for x in s.split()])
        except ValueError:
            if cur:
                blocks.append(np.array(cur))
                cur = []
    if cur:
        blocks.append(np.array(cur))
    return blocks


def odo_tags(path):
    """the tagged numbers of an .odo, with the rules of make_golden.golden_odo"""
    out = {}
    num = r"[-+]?\d+\.\d+(?:[Ee][-+]?\d+)?"
    for line in open(path):
        m = re.search(r"Fermi energy \((\w+) broadening\) :\s*([-\d.]+)", line)
        if m:
            out["efermi_" + m.group(1).lower()] = float(m.group(2))
        m = re.search(r"<- (\w+) \|", line)
        if not m:
            continue
        tag, body = m.group(1), line[:m.start()]
        if tag in ("OBg", "ABg"):
            out.setdefault(tag, []).append(float(re.search(r"Spin :\s*(\d+)\s*:\s*(" + num + ")", body).group(2)))
        elif tag in ("DEA", "DEF", "DEL"):
            out.setdefault(tag, []).append(float(re.search(r"DOS at Fermi Energy :\s*(" + num + ")", body).group(1)))
        elif tag == "Occ":
            mm = re.search(r"occupation between\s*(" + num + r")\s*and\s*(" + num + ")", body)
            out.setdefault(tag, []).append([float(mm.group(1)), float(mm.group(2))])
        elif tag in ("TBg", "wAB", "BEA", "BEF", "BEL", "BEC"):
            out[tag] = float(re.findall(num, body)[-1])
    return out


def test_param_read_every_odi_of_the_test_suite():
    """od_param_read (param_read, parameters.f90:138-465) on every .odi the reference ships with its tests (host only)"""
    import optados_b200 as ob
    seen = 0
    for test, info in INDEX.items():
        p = ob.param_read(os.path.join(TS, "tests", test, info["seed"] + ".odi"))
        txt = open(os.path.join(TS, "tests", test, info["seed"] + ".odi")).read().lower()
        task = re.search(r"^\s*task\s*[:=]?\s*(\w+)", txt, re.M).group(1)
        assert p[task] == 1, (test, task)
        m = re.search(r"^\s*broadening\s*[:=]?\s*(\w+)", txt, re.M)
        assert p[m.group(1) if m else "adaptive"] == 1, test
        assert p["fixed"] + p["adaptive"] + p["linear"] == 1
        seen += 1
    assert seen >= 25
    p = ob.param_read(os.path.join(TS, "tests", "testopt_core_chemical_shift", "NaGe.odi"))
    assert p["core"] and p["adaptive_smearing"] == 0.2 and p["core_lai_broadening"] and p["lai_gaussian_width"] == 1.0
    assert p["lai_lorentzian_width"] == 0.5 and p["lai_lorentzian_scale"] == 0.1 and abs(p["core_chemical_shift"] - 11142.275222) < 1e-9
    assert p["dos_spacing"] == 0.1 and p["compute_band_gap"] == 1
    p = ob.param_read(os.path.join(TS, "tests", "testopt_optics_polar", "Si2.odi"))
    assert p["optics"] and p["optics_geom"].startswith("polar") and len(p["optics_qdir"]) == 3
    p = ob.param_read(os.path.join(TS, "tests", "testopt_pdos_string1", "Si2.odi"))
    assert p["pdos"] and p["projectors_string"] == "sum:si1-2(s)" and p["fixed"]


def test_param_read_rules(tmp_path):
    """the tokenizer: comments, `=` / `:` / blank separators, logicals by letter, ranges, unknown keywords, defaults"""
    import optados_b200 as ob
    f = tmp_path / "x.odi"
    f.write_text("TASK = compare_dos   ! three schemes\n# a comment\nDOS_NBINS 300\nexclude_bands : 1,3-5 9\nefermi = 4.25\n"
                 "numerical_intdos : .true.\nFINITE_BIN_CORRECTION = F\n")
    p = ob.param_read(f)
    assert p["dos"] and p["compare_dos"] and p["fixed"] and p["adaptive"] and p["linear"]
    assert p["dos_nbins"] == 300 and p["exclude_bands"] == [1, 3, 4, 5, 9]
    assert p["efermi_choice"] == 2 and p["efermi_user"] == 4.25 and p["numerical_intdos"] == 1 and p["finite_bin_correction"] == 0
    f.write_text("task : dos\n")
    p = ob.param_read(f)
    assert p["adaptive"] == 1 and p["dos_spacing"] == float(np.float32(0.005)) and p["dos_nbins"] == -1      # parameters.f90:459-461
    f.write_text("task : dos\nno_such_keyword : 1\n")
    with pytest.raises(ob.OptadosB200Error):
        ob.param_read(f)
    f.write_text("task : dos\ntask : jdos\n")
    with pytest.raises(ob.OptadosB200Error):
        ob.param_read(f)
    f.write_text("task : optics\noptics_geom : polarised\n")      # optics_qdir missing
    with pytest.raises(ob.OptadosB200Error):
        ob.param_read(f)


@pytest.mark.gpu
@pytest.mark.parametrize("test", RUNNABLE)
def test_files_in_files_out(test, tmp_path):
    import optados_b200 as ob
    info = INDEX[test]
    d, seed = stage(test, tmp_path)
    ctx = ob.Context(0)
    ctx.run_seed(d, seed)
    ctx.close()
    out = d / info["output"]
    assert out.exists(), sorted(os.listdir(d))
    if info["program"] == "OPTADOS_ODO_OK":
        G = json.load(open(os.path.join(GOLDEN, "gold_odo.json")))[test]
        T = odo_tags(out)
        for k, v in G.items():
            if k in ("vbm_cbm_multiplicity", "gap_kind"):
                continue
            assert k in T, (k, T)
            # print precision of the .odo: f15.10 gaps, f10.5 occupations, f8.4 / f12.4 energies
            tol = 5.1e-11 if k in ("TBg", "OBg", "ABg", "wAB") else (5.1e-6 if k == "Occ" else 5.1e-5)
            assert np.abs(np.asarray(T[k], dtype=float) - np.asarray(v, dtype=float)).max() <= tol, (k, T[k], v)
        if "gap_kind" in G:
            assert G["gap_kind"] in open(out).read()
        return
    G = load_npz("gold_" + test)
    B = numeric_blocks(out)
    assert len(B) == len(G), (len(B), len(G))
    pdos = ".pdos." in info["output"]
    tensor = test == "testopt_optics_tensor"
    for i, b in enumerate(B):
        g = G[f"block{i}"]
        assert b.shape == g.shape, (i, b.shape, g.shape)
        assert np.abs(b[:, 0] - g[:, 0]).max() <= (1e-6 if pdos else 1e-9) * max(1.0, np.abs(g[:, 0]).max())
        for c in range(1, g.shape[1]):
            peak = np.abs(g[:, c]).max()
            if tensor:      # the off-diagonal components vanish by symmetry (1e-17 of noise in the benchmark): relative to the diagonal
                peak = max(np.abs(G[f"block{k}"][:, c]).max() for k in range(len(G)))
            if peak == 0.0:
                assert np.abs(b[:, c]).max() < 1e-20
                continue
            # Print precision of the benchmarks: E21.13 tables (DOS, JDOS) 5e-13 of peak, es14.7 PDOS tables 6e-8, list-directed
            # files 17 digits.  DOS / JDOS tables are held to 2e-12; epsilon, loss function and core edges (Kramers-Kronig and
            # LAI sums over the grid on top of the spectra; the linear scheme's closed-form corners) to north_star's 1e-9 of
            # peak -- what the GPU path measures against the oracle is 4e-11 at worst (core_all).
            tol = 2e-7 if pdos else (1e-9 if ("epsilon" in info["output"] or "loss_fn" in info["output"] or "core_edge" in info["output"]) else 2e-12)
            assert np.abs(b[:, c] - g[:, c]).max() <= tol * peak, (i, c, np.abs(b[:, c] - g[:, c]).max() / peak)
