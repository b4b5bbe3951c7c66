"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol include/chs.h
declares, validates its arguments, and the product refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    import __graft_entry__ as ge
    from casualhdrsplat_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        ge.build()
    return _lib.lib()


def test_exports_every_declared_symbol(L):
    from casualhdrsplat_b200 import _lib

    hdr = open(os.path.join(ROOT, "include", "chs.h")).read()
    declared = set(re.findall(r"CHS_API[^;(]*?\b(chs_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 17
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(L, name), name


def test_version_and_struct_layout(L):
    from casualhdrsplat_b200 import _lib

    assert L.chs_version() == 200
    assert ctypes.sizeof(_lib.ChsConfig) == 4 * (6 + 3 + 5 + 3 + 2 + 8) == L.chs_sizeof(0)
    assert ctypes.sizeof(_lib.ChsWorkspaceSizes) == L.chs_sizeof(1) and ctypes.sizeof(_lib.ChsTensors) == L.chs_sizeof(2)
    assert L.chs_sizeof(3) == 0


def test_argument_validation_sets_error_message(L):
    from casualhdrsplat_b200 import _lib

    cfg = _lib.make_config(10, 1, 1, 32, 32, tile_size=8)
    out = _lib.ChsWorkspaceSizes()
    st = L.chs_workspace_query(ctypes.byref(cfg), 0, 0, ctypes.byref(out))
    assert st == -1 and b"tile_size" in L.chs_last_error()
    cfg = _lib.make_config(10, 1, 0, 32, 32)
    assert L.chs_workspace_query(ctypes.byref(cfg), 0, 0, ctypes.byref(out)) == -1
    with pytest.raises(RuntimeError, match="tile_size"):
        _lib.check(L.chs_workspace_query(ctypes.byref(_lib.make_config(1, 1, 1, 8, 8, tile_size=4)), 0, 0, ctypes.byref(out)))
    st = L.chs_spline_fwd(7, None, 0, 0.0, 1.0, None, None, 0, 1, None, None)
    assert st == -1 and b"kind" in L.chs_last_error()


def test_no_cpu_fallback():
    from casualhdrsplat_b200 import rasterize

    z = torch.zeros
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        rasterize(z(4, 3), z(4, 4), z(4, 3), z(4), z(4, 3), z(1, 4, 4), z(1, 3, 3), 16, 16, z(1), 1)


def test_loss_entry_points_validate_and_refuse_cpu(L):
    from casualhdrsplat_b200 import _lib
    from casualhdrsplat_b200.train import photometric_loss, ssim_loss

    z = torch.zeros
    with pytest.raises(RuntimeError, match="no CPU path"):
        ssim_loss(z(1, 8, 8, 3), z(1, 8, 8, 3))
    with pytest.raises(RuntimeError, match="no CPU path"):
        photometric_loss(z(1, 8, 8, 3), z(1, 8, 8, 3))
    # argument validation happens before any device work
    assert L.chs_ssim_loss(None, None, 1, 8, 8, 0.8, 0.2, None, None, None, 0, None) == -1 and b"null" in L.chs_last_error()
    assert L.chs_ssim_loss(None, None, 1, 0, 8, 0.8, 0.2, None, None, None, 0, None) == -1 and b"shape" in L.chs_last_error()
    assert L.chs_ssim_loss(None, None, 0, 8, 8, 0.8, 0.2, None, None, None, 0, None) == 0  # no images: nothing to do
    assert L.chs_loss(2, None, None, 0, 1.0, None, None, None) == -1 and b"kind" in L.chs_last_error()


def test_crf_kinds_and_sizes(L):
    from casualhdrsplat_b200 import _lib

    out = _lib.ChsWorkspaceSizes()
    for kind, size, ok in [(_lib.CHS_CRF_MLP, 64, True), (_lib.CHS_CRF_MLP, 129, False), (_lib.CHS_CRF_LUT, 256, True),
                           (_lib.CHS_CRF_LUT, 1, False), (_lib.CHS_CRF_LUT, 1025, False), (3, 8, False)]:
        cfg = _lib.make_config(0, 1, 1, 32, 32, crf_kind=kind, crf_hidden=size)  # no Gaussians: no CUB size query, no device needed
        assert (L.chs_workspace_query(ctypes.byref(cfg), 0, 4, ctypes.byref(out)) == 0) == ok, (kind, size)
    # parameter-tensor shapes of the operator
    assert _lib.crf_size(_lib.CHS_CRF_MLP, torch.zeros(3, 3 * 16 + 1)) == 16
    assert _lib.crf_size(_lib.CHS_CRF_LUT, torch.zeros(3, 34)) == 32
    assert _lib.crf_size(_lib.CHS_CRF_IDENTITY, None) == 0
    for kind, shape in [(_lib.CHS_CRF_MLP, (3, 48)), (_lib.CHS_CRF_LUT, (3, 3)), (_lib.CHS_CRF_MLP, (2, 49)), (7, (3, 49))]:
        with pytest.raises(RuntimeError):
            _lib.crf_size(kind, torch.zeros(*shape))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "casualhdrsplat_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
    # nothing under tools/ touches the oracle either (diagnostics that need it live under tests/tools/)
    for f in os.listdir(os.path.join(ROOT, "tools")):
        if f.endswith(".py"):
            src = open(os.path.join(ROOT, "tools", f)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle|oracle_run", src, flags=re.M), f
