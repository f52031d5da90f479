"""ctypes loader for libmicromix_b200.so (the C ABI declared in include/micromix_b200.h).

There is NO fallback: if the library is missing or cannot be loaded, importing the ops raises.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmicromix_b200.so")

# every symbol include/micromix_b200.h declares: name -> (restype, argtypes)
_vp, _i64, _i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
_QUANT_ARGS = [_vp, _i64, _i32, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]
SYMBOLS = {
    "mmx_version": (_i32, []),
    "mmx_last_error": (ctypes.c_char_p, []),
    "mmx_sf_bytes_act": (_i64, [_i64, _i64]),
    "mmx_sf_bytes_wgt": (_i64, [_i64, _i64]),
    "mmx_sf_offset": (_i64, [_i64, _i64, _i64]),
    "mmx_reorder_quantize_x": (_i32, _QUANT_ARGS),
    "mmx_reorder_quantize_w": (_i32, _QUANT_ARGS),
    "mmx_reorder_quantize_w4": (_i32, _QUANT_ARGS),
    "mmx_rmsnorm_quantize_x": (_i32, [_vp, _vp, ctypes.c_float, _i64, _i32, _vp, _i32, _i32, _i32] + [_vp] * 7),
    "mmx_activate_quantize_x": (_i32, [_vp, _vp, _i64, _i32, _i32, _i32] + [_vp] * 7),
    "mmx_activate_quantize_x_strided": (_i32, [_vp, _vp, _i64, _i64, _i32, _i32, _i32] + [_vp] * 7),
    "mmx_downproj_quantize_w": (_i32, [_vp, _i64, _i32, _i32, _i32] + [_vp] * 7),
    "mmx_downproj_quantize_w4": (_i32, [_vp, _i64, _i32, _i32, _i32] + [_vp] * 7),
    "mmx_matmul": (_i32, [_vp] * 12 + [_i64, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "mmx_matmul_residual": (_i32, [_vp] * 12 + [_i64, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "mmx_matmul_grouped_activate_quantize": (_i32, [_vp] * 12 + [_i64, _i64] + [_i32] * 6 + [_vp, _vp] + [_i32] * 3 + [_vp] * 7),
    "mmx_matmul_rope": (_i32, [_vp] * 12 + [_i64, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _i64, _i32, _vp, _vp]),
    "mmx_moe_route": (_i32, [_vp, _vp, _i64, _i32, _i32, _i32, _i32, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mmx_debug_silu_table": (_i32, [_vp, _vp, _vp]),
    "mmx_matmul_activate_quantize": (_i32, [_vp] * 12 + [_i64, _i64] + [_i32] * 7 + [_vp] * 7),
    "mmx_peer_alloc": (_i32, [_i64, ctypes.POINTER(_vp), ctypes.c_char_p]),
    "mmx_peer_open": (_i32, [ctypes.c_char_p, ctypes.POINTER(_vp)]),
    "mmx_peer_close": (_i32, [_vp]),
    "mmx_peer_free": (_i32, [_vp]),
    "mmx_tp_workspace_bytes": (_i64, [_i64, _i64, _i32]),
    "mmx_tp_ctx_create": (_i32, [ctypes.POINTER(_vp), _i32, _i32, _i64, _i64, ctypes.POINTER(_vp)]),
    "mmx_tp_ctx_set_multicast": (_i32, [_vp, _vp, _i32]),
    "mmx_tp_ctx_destroy": (_i32, [_vp]),
    "mmx_tp_status": (_i32, [_vp, ctypes.POINTER(ctypes.c_uint32)]),
    "mmx_matmul_allreduce": (_i32, [_vp] + [_vp] * 12 + [_i64, _i64, _i32, _i32, _i32, _i32, _vp,
                                                       ctypes.POINTER(_vp), _vp]),
    "mmx_tp_workspace_bytes_ex": (_i64, [_i64, _i64, _i32, _i64, _i64]),
    "mmx_tp_ctx_create_ex": (_i32, [ctypes.POINTER(_vp), _i32, _i32, _i64, _i64, _i64, _i64, ctypes.POINTER(_vp)]),
    "mmx_tp_shard_rows": (_i64, [_i64, _i32]),
    "mmx_matmul_reduce_scatter": (_i32, [_vp] + [_vp] * 12 + [_i64, _i64, _i32, _i32, _i32, _i32, _vp,
                                                            ctypes.POINTER(_vp), ctypes.POINTER(_i64),
                                                            ctypes.POINTER(_i64), _vp]),
    "mmx_tp_quantize_allgather": (_i32, [_vp, _vp, _i64, _i32, _vp, _i32, _i32, _i32, _vp, ctypes.c_float,
                                         ctypes.POINTER(_vp), _vp]),
    "mmx_tp_matmul_gathered": (_i32, [_vp] + [_vp] * 6 + [_i64, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "mmx_tp_quantize_alltoall": (_i32, [_vp, _vp, _i64, _i32, _vp, _i32, _i32, _i32, ctypes.POINTER(ctypes.c_int32),
                                        ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(_vp), _vp]),
    "mmx_tp_matmul_exchanged": (_i32, [_vp] + [_vp] * 6 + [_i64, _i64, _i32, _i32, _i32, _i32, _vp, _vp,
                                                         ctypes.POINTER(_i64), ctypes.POINTER(_i64), _vp]),
    "mmx_reorder_quantize_x_grouped": (_i32, [_vp, _i64, _i32, _vp, _vp, _vp, _vp, _i32, _i32, _i32] + [_vp] * 7),
    "mmx_activate_quantize_x_rows": (_i32, [_vp, _vp, _i64, _i64, _vp, _i32, _i32, _i32] + [_vp] * 7),
    "mmx_rope_inplace": (_i32, [_vp, _i64, _i64, _i32, _i32, _vp, _vp, _i64, _vp]),
    "mmx_moe_combine": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp]),
    "mmx_matmul_grouped": (_i32, [_vp] * 12 + [_i64, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "mmx_debug_mma_peak": (_i32, [_i32, _i32, _i32, _i32, ctypes.POINTER(ctypes.c_double),
                                  ctypes.POINTER(ctypes.c_double)]),
    "mmx_launch_count": (_i64, []),
    "mmx_set_option": (_i32, [ctypes.c_char_p, _i64]),
    "mmx_gemm_debug_status": (_i32, [ctypes.POINTER(ctypes.c_uint32), _i32]),
    "mmx_tp_debug_times": (_i32, [ctypes.POINTER(ctypes.c_uint64), _i32]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load the CUDA library; raise loudly when it is absent (no CPU / eager fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  micromix_b200 has no fallback path.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the header and the library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error() -> str:
    return load().mmx_last_error().decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = last_error()
        if rc == -1:
            raise ValueError(f"{what}: {msg}")
        raise RuntimeError(f"{what} failed (rc={rc}): {msg}")
