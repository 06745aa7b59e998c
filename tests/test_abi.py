"""The C-ABI library loads and exports every symbol include/s2m.h declares (no GPU needed)."""
import ctypes
import os
import re

from conftest import ROOT


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "s2m.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(s2m_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree(s2m):
    assert declared_symbols() == sorted(s2m.EXPORTS)


def test_library_exports_every_symbol(s2m, built):
    lib = ctypes.CDLL(s2m.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_struct_sizes(s2m, built):
    # s2m_params: 2 floats + 11 ints; s2m_stats: 15 ints (+pad) + 4 doubles
    assert ctypes.sizeof(s2m.Params) == 52
    assert ctypes.sizeof(s2m.Stats) == 96
    p = s2m.default_params()
    assert abs(p.line_res - 0.4) < 1e-7 and abs(p.plane_res - 0.8) < 1e-7 and p.batch == 1


def test_strerror(s2m, built):
    L = s2m.load_library()
    assert L.s2m_strerror(0) == b"ok"
    assert b"not enough" in L.s2m_strerror(1)


def test_no_cpu_fallback(s2m, built):
    """Without a CUDA device the product must fail loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        return
    import pytest
    with pytest.raises(s2m.S2MError):
        s2m.Registrar()


def test_product_never_touches_oracle():
    pk = os.path.join(ROOT, "sc-a-loam_b200")
    for dp, _, fs in os.walk(pk):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "orc_" not in txt and "libs2m_oracle" not in txt, f
