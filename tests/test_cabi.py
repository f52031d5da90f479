"""CPU tests of the drop-in boundary: the C-ABI library builds, loads and exports every symbol include/*.h declares;
geometry helpers agree with the oracle; the Python shim mirrors the reference op surface and fails loudly."""
import os
import re

import numpy as np
import pytest
import torch

import helpers as H


def _declared_symbols():
    hdr = open(os.path.join(H.ROOT, "include", "micromix_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(mmx_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol(mmx_lib):
    from micromix_b200 import _lib
    declared = _declared_symbols()
    assert len(declared) >= 12
    assert set(declared) == set(_lib.SYMBOLS), "ctypes table and header disagree"
    for name in declared:
        assert hasattr(mmx_lib, name), name
    assert mmx_lib.mmx_version() >= 100


def test_sf_geometry_matches_oracle(mmx_lib):
    O = H.O
    rng = np.random.default_rng(0)
    for _ in range(200):
        k = int(rng.integers(1, 64)) * 128
        r = int(rng.integers(0, 5000))
        g = int(rng.integers(0, k // 32))
        assert mmx_lib.mmx_sf_offset(r, g, k) == int(O.sf_offset(r, g, k))
    for M in (1, 127, 128, 129, 2048):
        assert mmx_lib.mmx_sf_bytes_act(M, 2560) == (M // 128 + 1) * 128 * 2560 // 32  # bindings.cpp:120-123
    assert mmx_lib.mmx_sf_bytes_wgt(4096, 1024) == 4096 * 1024 // 32                    # bindings.cpp:170-172


def test_argument_validation_without_gpu(mmx_lib):
    """Invalid shapes are rejected before any CUDA call (no GPU needed), with a readable message."""
    from micromix_b200 import _lib
    rc = mmx_lib.mmx_reorder_quantize_x(None, 4, 256, None, 100, 100, 56, None, None, None, None, None, None, None)
    assert rc == -1 and "multiples of 128" in _lib.last_error()
    rc = mmx_lib.mmx_reorder_quantize_x(None, 4, 256, None, 128, 128, 128, None, None, None, None, None, None, None)
    assert rc == -1 and "bad shape" in _lib.last_error()
    rc = mmx_lib.mmx_matmul(*([None] * 12), 16, 100, 128, 0, 0, 1, None, None, None)
    assert rc == -1 and "multiple of 128" in _lib.last_error()
    assert mmx_lib.mmx_set_option(b"no_such_option", 1) == -1
    rc = mmx_lib.mmx_activate_quantize_x(None, None, 4, 100, 100, 56, None, None, None, None, None, None, None)
    assert rc == -1 and "multiples of 128" in _lib.last_error()
    rc = mmx_lib.mmx_downproj_quantize_w4(None, 4, 128, 0, 0, None, None, None, None, None, None, None)
    assert rc == -1 and "null" in _lib.last_error()
    rc = mmx_lib.mmx_rmsnorm_quantize_x(None, None, 1e-5, 4, 256, None, 128, 128, 0, *([None] * 7))
    assert rc == -1 and "null" in _lib.last_error()


def test_shim_mirrors_reference_surface():
    from micromix_b200 import mixedgemm
    import inspect
    ref_ops = {  # bindings.cpp:686-735 names and argument names
        "matmul": ["AN", "BN", "AS", "BS", "AO", "BO", "SFAN", "SFBN", "SFAS", "SFBS", "SFAO", "SFBO"],
        "reorder_quantize_x": ["X", "reorder_index", "KN", "KS", "KO"],
        "reorder_quantize_w": ["W", "reorder_index", "KN", "KS", "KO"],
        "reorder_quantize_w4": ["W", "reorder_index", "KN", "KS", "KO"],
        "rmsnorm_quantize_x": ["X", "W", "eps", "reorder_index", "KN", "KS", "KO"],
        "activate_quantize_x": ["A", "B", "KN", "KS", "KO"],
        "downproj_quantize_w": ["W", "KN", "KS", "KO"],
        "downproj_quantize_w4": ["W", "KN", "KS", "KO"],
    }
    for name, args in ref_ops.items():
        params = list(inspect.signature(getattr(mixedgemm, name)).parameters)
        assert params[:len(args)] == args, name
    assert mixedgemm.test_function() == "Hello from test_function!"


def test_no_cpu_fallback():
    from micromix_b200 import mixedgemm
    x = torch.zeros(4, 128, dtype=torch.bfloat16)
    idx = torch.arange(128, dtype=torch.int16)
    with pytest.raises(RuntimeError, match="CUDA"):
        mixedgemm.reorder_quantize_x(x, idx, 128, 0, 0)
    with pytest.raises(RuntimeError, match="CUDA"):
        mixedgemm.matmul(*[torch.zeros(4, 64, dtype=torch.uint8)] * 12)
    with pytest.raises(RuntimeError, match="CUDA"):
        mixedgemm.activate_quantize_x(x, x, 128, 0, 0)
    with pytest.raises(RuntimeError, match="CUDA"):
        mixedgemm.downproj_quantize_w(x, 128, 0, 0)
    with pytest.raises(RuntimeError, match="CUDA"):
        mixedgemm.rmsnorm_quantize_x(x, x[0], 1e-5, idx, 128, 0, 0)


def test_product_does_not_import_oracle():
    """The product package must never import, include, link or dlopen anything under oracle/."""
    pkg = os.path.join(H.ROOT, "micromix_b200")
    bad = re.compile(r"(^\s*(from|import)\s+oracle\b)|(libmmx_oracle)|(#include\s+[\"<].*oracle)|(mmxo_)", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert not bad.search(src), f


def test_tp_entry_points_validate_without_a_gpu(mmx_lib):
    """Argument checks of the tensor-parallel entry points that need no device."""
    import ctypes
    assert mmx_lib.mmx_tp_workspace_bytes(8192, 4096, 3) > 0 or mmx_lib.mmx_tp_workspace_bytes(8192, 4096, 3) == -1
    assert mmx_lib.mmx_tp_workspace_bytes(0, 4096, 2) == -1
    assert mmx_lib.mmx_tp_workspace_bytes(8192, 4096, 9) == -1
    b2, b8 = mmx_lib.mmx_tp_workspace_bytes(8192, 4096, 2), mmx_lib.mmx_tp_workspace_bytes(8192, 4096, 8)
    assert b2 > 2 * 8192 * 4096 * 2 and b8 > 2 * 8192 * 4096 * 2  # two parities of C + staging + flags
    assert mmx_lib.mmx_tp_ctx_set_multicast(None, None, 0) != 0
    assert b"mmx_tp_ctx_set_multicast" in mmx_lib.mmx_last_error()
    ctx = ctypes.c_void_p()
    arr = (ctypes.c_void_p * 2)(None, None)
    assert mmx_lib.mmx_tp_ctx_create(arr, 2, 0, 8192, 4096, ctypes.byref(ctx)) != 0  # null workspaces
    assert mmx_lib.mmx_tp_ctx_create(arr, 3, 0, 8192, 4096, ctypes.byref(ctx)) != 0  # tp must be 1, 2, 4 or 8
    out = (ctypes.c_uint64 * 8)()
    assert mmx_lib.mmx_tp_debug_times(out, 0) != 0
