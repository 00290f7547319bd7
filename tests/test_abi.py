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
