"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol include/vdn.h declares, and
refuses to work without a GPU (no fallback)."""
import ctypes as C
import os
import re

import pytest

import varden_b200 as V

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "vdn.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(vdn_[a-z_0-9]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = V.load_library()
    syms = header_symbols()
    assert len(syms) >= 25
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(V.ABI_SYMBOLS) == syms


def test_field_enum_matches_header():
    txt = open(os.path.join(ROOT, "include", "vdn.h")).read()
    body = re.search(r"enum vdn_field \{(.*?)\};", txt, re.S).group(1)
    names = [n.strip().split("=")[0].strip() for n in body.replace("\n", " ").split(",") if n.strip()]
    names = [n[4:] for n in names if n != "VDN_NFIELDS"]
    assert names == V.FIELDS


def test_params_default_matches_reference_defaults():
    p = V.default_params()
    # src/_parameters: nscal 2, slope_order 4, use_minion F, stencil_order 2; mac_multigrid.f90:56 bottom eps 1e-3
    assert (p.nscal, p.slope_order, p.use_minion, p.stencil_order) == (2, 4, 0, 2)
    assert p.mg_bottom_eps == 1e-3 and p.mg_nu1 == 2 and p.mg_nu2 == 2


def test_no_cpu_fallback():
    """without a CUDA device context creation must fail loudly"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(V.VdnError):
        V.Context(3, [([0, 0, 0], [7, 7, 7])], [0, 0, 0], [7, 7, 7], [[-1, -1]] * 3, [0.125] * 3)


def test_product_does_not_import_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "varden_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert "oracle" not in src.replace("no oracle", "").replace("CPU oracle", "").replace("the oracle", "").replace("oracle/oracle.py", ""), f
