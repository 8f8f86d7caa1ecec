"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol
include/ehdg.h declares; without a device it refuses to create a context (no CPU path)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "ehdg.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(ehdg_[a-z0-9_]+)\s*\(", hdr)))


def test_header_and_binding_agree():
    from convectiondiffusionequation_b200 import _lib
    assert _declared_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol(libehdg_built):
    from convectiondiffusionequation_b200 import _lib
    lib = _lib.load()
    for name in _declared_symbols():
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.ehdg_version()


def test_struct_layouts_match_header():
    from convectiondiffusionequation_b200 import _lib
    assert C.sizeof(_lib.LayerExpr) == 48
    assert C.sizeof(_lib.Problem) == 4 + 4 + 8 + 16 + 4 + 8 + 4 + 16 + 48 + 48      # incl. padding after n_enrich arrays
    assert C.sizeof(_lib.Mesh) == 16 + 4 * 8
    assert C.sizeof(_lib.GmresInfo) == 128


def test_no_cpu_path(libehdg_built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from convectiondiffusionequation_b200 import _lib
    from convectiondiffusionequation_b200.hot_path import EHDGPipeline, ProblemSpec
    from convectiondiffusionequation_b200.mesh import unit_square_mesh
    lib = _lib.load()
    p = _lib.Problem()
    p.order, p.epsilon = 2, 0.01
    ctx = C.c_void_p()
    assert lib.ehdg_create(C.byref(p), C.byref(ctx)) == -2
    assert b"no CPU path" in lib.ehdg_last_error()
    with pytest.raises(_lib.EhdgError):
        EHDGPipeline(unit_square_mesh(2), ProblemSpec(order=1, epsilon=0.1, beta=(1, 0)))


def test_product_never_imports_the_oracle():
    """the package and the measurement helpers under tools/ never touch oracle/ (host studies that do live in tests/studies)"""
    for top in ("convectiondiffusionequation_b200", "tools"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".h", ".cpp")):
                    src = open(os.path.join(dirpath, f)).read()
                    assert "ehdg_oracle" not in src and "from oracle" not in src and "import oracle" not in src, f
