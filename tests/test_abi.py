"""The C-ABI library loads and exports every symbol include/optados_b200.h declares (CPU only;
no compute calls)."""
import os
import re

import pytest

from conftest import ROOT


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "optados_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(od_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_header_symbols():
    import optados_b200
    if not os.path.exists(optados_b200.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    L = optados_b200.load()
    syms = header_symbols()
    assert len(syms) > 40
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(optados_b200.EXPORTS) == syms
    assert b"sm_100a" in L.od_version()


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly, not fall back."""
    import torch
    import optados_b200
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(optados_b200.OptadosB200Error):
        optados_b200.Context(0)


def test_dist_array_matches_reference():
    """algor_dist_array (algorithms.f90:309-340)"""
    import optados_b200
    assert optados_b200.dist_array(10, 4) == [3, 3, 2, 2]
    assert optados_b200.dist_array(8, 8) == [1] * 8
    assert optados_b200.dist_array(210, 8) == [27, 27] + [26] * 6      # remainder > 1: DO bounds are evaluated once
    assert optados_b200.dist_array(1000000, 8) == [125000] * 8
    with pytest.raises(optados_b200.OptadosB200Error):
        optados_b200.dist_array(3, 4)


def test_product_does_not_touch_oracle():
    """The shipped library and package must not import, link or execute anything under oracle/."""
    import subprocess
    import optados_b200
    out = subprocess.run(["ldd", optados_b200.LIB_PATH], capture_output=True, text=True).stdout
    assert "od_oracle" not in out
    for root, _, files in os.walk(os.path.join(ROOT, "optados_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".f90")):
                src = open(os.path.join(root, f), errors="ignore").read()
                assert "libod_oracle" not in src and "from oracle" not in src and "import oracle" not in src, f


def test_structure_layouts_match_the_bindings():
    """sizeof of every public structure as the library was compiled, against the ctypes mirrors of the Python binding
    (same member order as the ISO_C_BINDING derived types of optados_b200/fortran/od_b200.f90, which checks itself against
    od_abi_sizeof in b200_setup) -- the place where an ABI mismatch of a never-compiled interface would hide."""
    import ctypes as C
    import optados_b200 as ob
    L = ob.load()
    pairs = {"od_broadening": ob.Broadening, "od_dos_params": ob.DosParams, "od_dos_result": ob.DosResult,
             "od_jdos_params": ob.JdosParams, "od_jdos_result": ob.JdosResult, "od_optics_params": ob.OpticsParams,
             "od_optics_result": ob.OpticsResult, "od_core_params": ob.CoreParams, "od_core_result": ob.CoreResult,
             "od_bandgap_result": ob.BandgapResult, "od_odi": ob.Odi}
    for name, cls in pairs.items():
        assert L.od_abi_sizeof(name.encode()) == C.sizeof(cls), name
    assert L.od_abi_sizeof(b"no_such_struct") == -1
    # every interface of the Fortran shim names a symbol the library exports, and every derived type a structure it knows
    src = open(os.path.join(ROOT, "optados_b200", "fortran", "od_b200.f90")).read()
    for sym in set(re.findall(r"bind\(c, name='(od_[a-z0-9_]+)'\)", src)):
        assert hasattr(L, sym), sym
    for t in set(re.findall(r"type, bind\(c\), public :: (od_[a-z_]+)", src)):
        assert L.od_abi_sizeof(t.encode()) > 0, t
    # member order of the shim's derived types = member order of the header's structs (names, in sequence)
    hdr = open(os.path.join(ROOT, "include", "optados_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    for t in set(re.findall(r"type, bind\(c\), public :: (od_[a-z_]+)", src)):
        body = re.search(r"type, bind\(c\), public :: " + t + r"\n(.*?)end type", src, re.S).group(1)
        body = re.sub(r"!.*", "", body)
        f_members = []
        for line in body.splitlines():
            if "::" in line:
                f_members += [m.split("(")[0].strip() for m in re.sub(r"\([^)]*\)", "", line.split("::")[1]).split(",")]
        m = re.search(r"typedef struct " + t + r" \{(.*?)\} " + t + ";", hdr, re.S)
        if not m:
            m = re.search(r"struct " + t + r" \{(.*?)\}", hdr, re.S)
        c_members = []
        for decl in m.group(1).split(";"):
            decl = decl.strip()
            if not decl:
                continue
            decl = re.sub(r"^(const\s+)?(unsigned\s+)?(long long|double|int|char|od_[a-z_]+)\s*\*?", "", decl)
            c_members += [x.strip().lstrip("*").split("[")[0].strip() for x in decl.split(",")]
        assert f_members == c_members, (t, f_members, c_members)
