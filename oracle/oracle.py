"""CPU oracle for the MicroMix hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` leg may
import this module.  Nothing under ``micromix_b200/`` does.

Thin numpy/ctypes wrapper over ``oracle/mmx_oracle.c`` (the scalar C restatement of
``/root/reference/mgemm/src/reorder.cu:94-432``) plus the GEMM semantics of
``/root/reference/mgemm/src/gemm.cu:53-78`` (three block-scaled GEMMs chained through a bf16 ``D`` with beta=1:
``w4a4.cu:176``, ``w4a6.cu:178``, ``w4a8.cu:178``) stated as dequantise -> fp32 matmul -> bf16, the recipe of the
CUTLASS host reference (``cutlass/tools/util/include/cutlass/util/reference/host/gett.hpp:549-600``).

Parity pin: see the header of ``mmx_oracle.c`` -- pinned against the vendored CUTLASS converters run on the host
(``tests/golden/cutlass_convert_table.npz``) and against the reference's own ``reorder.cu`` run on a B200
(``tests/golden/ref_reorder_*.npz``).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmmx_oracle.so")
_lib = None

FMT_X = (4, 6, 8)   # reorder_quantize_x / reorder_quantize_w
FMT_W4 = (4, 4, 4)  # reorder_quantize_w4


def build(force: bool = False) -> str:
    """Compile the C oracle (gcc, a second).  Building the checker is not using it."""
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(
            os.path.join(_HERE, "mmx_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "liboracle"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        u8p, i16p, u16p, f32p = (ctypes.POINTER(t) for t in (ctypes.c_uint8, ctypes.c_int16, ctypes.c_uint16,
                                                             ctypes.c_float))
        L.mmxo_reorder_quantize.argtypes = [u16p, ctypes.c_int64, ctypes.c_int, i16p] + [ctypes.c_int] * 6 + [u8p] * 6
        L.mmxo_reorder_quantize.restype = ctypes.c_int
        L.mmxo_dequant.argtypes = [u8p, u8p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, f32p]
        L.mmxo_dequant.restype = ctypes.c_int
        L.mmxo_sf_offset.argtypes = [ctypes.c_int64] * 3
        L.mmxo_sf_offset.restype = ctypes.c_int64
        L.mmxo_encode.argtypes = [ctypes.c_float, ctypes.c_int]
        L.mmxo_encode.restype = ctypes.c_uint8
        L.mmxo_decode.argtypes = [ctypes.c_uint8, ctypes.c_int]
        L.mmxo_decode.restype = ctypes.c_float
        L.mmxo_ue8m0_from_float.argtypes = [ctypes.c_float]
        L.mmxo_ue8m0_from_float.restype = ctypes.c_uint8
        L.mmxo_scale_byte_int.argtypes = [ctypes.c_uint16, ctypes.c_int]
        L.mmxo_scale_byte_int.restype = ctypes.c_int
        L.mmxo_scale_byte_float.argtypes = [ctypes.c_uint16, ctypes.c_int]
        L.mmxo_scale_byte_float.restype = ctypes.c_int
        L.mmxo_num_threads.restype = ctypes.c_int
        L.mmxo_cuda_log2f.argtypes = [ctypes.c_float]
        L.mmxo_cuda_log2f.restype = ctypes.c_float
        L.mmxo_act_scale_exp.argtypes = [ctypes.c_float, ctypes.c_int]
        L.mmxo_act_scale_exp.restype = ctypes.c_int
        L.mmxo_act_scale_exp_fast.argtypes = [ctypes.c_float, ctypes.c_int]
        L.mmxo_act_scale_exp_fast.restype = ctypes.c_int
        L.mmxo_check_scale_shortcut.argtypes = [ctypes.c_int, ctypes.c_int]
        L.mmxo_check_scale_shortcut.restype = ctypes.c_int64
        L.mmxo_silu_mul.argtypes = [u16p, u16p, ctypes.c_int64, f32p]
        L.mmxo_silu_mul.restype = None
        L.mmxo_quantize_f32.argtypes = [f32p, ctypes.c_int64, ctypes.c_int] + [ctypes.c_int] * 6 + [u8p] * 6
        L.mmxo_quantize_f32.restype = ctypes.c_int
        L.mmxo_rmsnorm.argtypes = [u16p, u16p, ctypes.c_float, ctypes.c_int64, ctypes.c_int, u16p]
        L.mmxo_rmsnorm.restype = ctypes.c_int
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


# ------------------------------------------------------------------ bf16 <-> numpy
def f32_to_bf16_bits(a: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even fp32 -> bf16 bit patterns (uint16)."""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    lsb = (u >> 16) & 1
    r = ((u + 0x7FFF + lsb) >> 16).astype(np.uint16)
    nan = (u & 0x7FFFFFFF) > 0x7F800000
    r[nan] = 0x7FFF
    return r


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (np.ascontiguousarray(b, dtype=np.uint16).astype(np.uint32) << 16).view(np.float32)


# ------------------------------------------------------------------ sizes (bindings.cpp:115-123,165-172,216-223)
def packed_width(kseg: int, fmt: int) -> int:
    return kseg * fmt // 8


def sf_bytes(rows: int, kseg: int, is_act: bool) -> int:
    if is_act:
        return (rows // 128 + 1) * 128 * kseg // 32
    return -(-rows // 128) * 128 * kseg // 32  # == rows*kseg/32 when rows % 128 == 0 (the reference's assumption)


def sf_offset(r, g, kseg):
    """Vectorised closed form of the SfKMajorAtom layout (sm100_blockscaled_layout.hpp:54-55,93)."""
    r = np.asarray(r, dtype=np.int64)
    g = np.asarray(g, dtype=np.int64)
    katoms = (kseg + 127) // 128
    return (r // 128) * katoms * 512 + (g // 4) * 512 + (r % 32) * 16 + ((r // 32) % 4) * 4 + (g % 4)


# ------------------------------------------------------------------ quantize
def reorder_quantize(x_bits: np.ndarray, idx: np.ndarray, KN: int, KS: int, KO: int, mode: str = "x",
                     sf_fill: int = 0):
    """x_bits: uint16 [rows, K] bf16 bit patterns.  mode 'x' | 'w' | 'w4' (bindings.cpp:104,155,206).

    Returns (QN, QS, QO, SFN, SFS, SFO) as uint8 arrays with the reference's shapes.  SF bytes the reference
    leaves unwritten (padding rows) are set to ``sf_fill``.
    """
    x_bits = np.ascontiguousarray(x_bits, dtype=np.uint16)
    idx = np.ascontiguousarray(idx, dtype=np.int16)
    rows, K = x_bits.shape
    assert KN + KS + KO == K and idx.shape == (K,)
    fm = FMT_W4 if mode == "w4" else FMT_X
    is_act = mode == "x"
    q = [np.zeros((rows, packed_width(k, f)), dtype=np.uint8) for k, f in zip((KN, KS, KO), fm)]
    sf = [np.full((sf_bytes(rows, k, is_act),), sf_fill, dtype=np.uint8) for k in (KN, KS, KO)]
    L = lib()
    rc = L.mmxo_reorder_quantize(_p(x_bits, ctypes.c_uint16), rows, K, _p(idx, ctypes.c_int16), KN, KS, KO, *fm,
                                 *[_p(a, ctypes.c_uint8) for a in q], *[_p(a, ctypes.c_uint8) for a in sf])
    if rc != 0:
        raise ValueError(f"mmxo_reorder_quantize rc={rc}")
    return (*q, *sf)


# ------------------------------------------------------------------ ops without a permutation (activate.cu)
def silu_mul(a_bits: np.ndarray, b_bits: np.ndarray) -> np.ndarray:
    """fp32 silu(a) * b of activate.cu:29,107 with libm's expf (CUDA's ends in ex2.approx: not restatable)."""
    a_bits = np.ascontiguousarray(a_bits, dtype=np.uint16)
    b_bits = np.ascontiguousarray(b_bits, dtype=np.uint16)
    out = np.empty(a_bits.shape, dtype=np.float32)
    lib().mmxo_silu_mul(_p(a_bits, ctypes.c_uint16), _p(b_bits, ctypes.c_uint16), a_bits.size, _p(out, ctypes.c_float))
    return out


def quantize_f32(v: np.ndarray, KN: int, KS: int, KO: int, w4: bool = False, sf_fill: int = 0):
    """The activate / downproj quantizer (activate.cu:109-176) on fp32 values [rows, K] already in segment order:
    scale 2^ceil(log2f(amax/QMAX)) (1.0 when amax <= 1e-6), codes RNE straight from fp32.  SF buffers are sized like
    activations for all three ops (bindings.cpp:320-322,346-348,373-375)."""
    v = np.ascontiguousarray(v, dtype=np.float32)
    rows, K = v.shape
    assert KN + KS + KO == K
    fm = FMT_W4 if w4 else FMT_X
    q = [np.zeros((rows, packed_width(k, f)), dtype=np.uint8) for k, f in zip((KN, KS, KO), fm)]
    sf = [np.full((sf_bytes(rows, k, True),), sf_fill, dtype=np.uint8) for k in (KN, KS, KO)]
    rc = lib().mmxo_quantize_f32(_p(v, ctypes.c_float), rows, K, KN, KS, KO, *fm, *[_p(a, ctypes.c_uint8) for a in q],
                                 *[_p(a, ctypes.c_uint8) for a in sf])
    if rc != 0:
        raise ValueError(f"mmxo_quantize_f32 rc={rc}")
    return (*q, *sf)


def downproj_quantize(w_bits: np.ndarray, KN: int, KS: int, KO: int, w4: bool):
    """downproj_quantize_w / _w4 (activate.cu:204-507): bit-exact restatement (no transcendental but log2f)."""
    return quantize_f32(bf16_bits_to_f32(w_bits), KN, KS, KO, w4)


def rmsnorm(x_bits: np.ndarray, w_bits: np.ndarray, eps: float) -> np.ndarray:
    """bf16((x * w) * rinv) with the fixed-order sum of squares (see mmx_oracle.c); returns bf16 bits [rows, K]."""
    x_bits = np.ascontiguousarray(x_bits, dtype=np.uint16)
    w_bits = np.ascontiguousarray(w_bits, dtype=np.uint16)
    rows, K = x_bits.shape
    y = np.empty_like(x_bits)
    rc = lib().mmxo_rmsnorm(_p(x_bits, ctypes.c_uint16), _p(w_bits, ctypes.c_uint16), float(np.float32(eps)), rows, K,
                            _p(y, ctypes.c_uint16))
    if rc != 0:
        raise ValueError(f"mmxo_rmsnorm rc={rc}")
    return y


def rmsnorm_quantize(x_bits, w_bits, eps, idx, KN, KS, KO):
    """rmsnorm_quantize_x (bindings.cpp:257-303), intended semantics: norm -> bf16 -> reorder_quantize_x."""
    return reorder_quantize(rmsnorm(x_bits, w_bits, eps), idx, KN, KS, KO, "x")


def dequant(q: np.ndarray, sf: np.ndarray, rows: int, kseg: int, fmt: int) -> np.ndarray:
    out = np.zeros((rows, kseg), dtype=np.float32)
    if kseg == 0:
        return out
    q = np.ascontiguousarray(q, dtype=np.uint8)
    sf = np.ascontiguousarray(sf, dtype=np.uint8)
    rc = lib().mmxo_dequant(_p(q, ctypes.c_uint8), _p(sf, ctypes.c_uint8), rows, kseg, fmt, _p(out, ctypes.c_float))
    if rc != 0:
        raise ValueError(f"mmxo_dequant rc={rc}")
    return out


def sf_valid_mask(rows: int, kseg: int, nbytes: int) -> np.ndarray:
    """Boolean mask over an SF buffer: True where the reference writes (rows < M); padding is don't-care."""
    m = np.zeros((nbytes,), dtype=bool)
    if kseg:
        r = np.arange(rows)[:, None]
        g = np.arange(kseg // 32)[None, :]
        m[sf_offset(r, g, kseg).ravel()] = True
    return m


# ------------------------------------------------------------------ GEMM
def _segment_formats(AS, BS, AO, BO):
    """bindings.cpp:74 dispatch rule: equal packed widths -> symmetric (w6a6/w8a8), else w4 weights."""
    sym = AS.shape[1] == BS.shape[1] and AO.shape[1] == BO.shape[1]
    return (4, 6, 8), ((4, 6, 8) if sym else (4, 4, 4))


def matmul(AN, BN, AS, BS, AO, BO, SFAN, SFBN, SFAS, SFBS, SFAO, SFBO, chain: bool = True, f64: bool = False):
    """Oracle for mixedgemm.matmul (bindings.cpp:50-102).  Returns bf16 bit patterns uint16 [M, N].

    chain=True  : the reference's numerics -- one GEMM per segment, D rounded to bf16 after each (beta=1 chain).
    chain=False : one fp32 accumulator over all three segments, one bf16 rounding (what a fused kernel computes).
    """
    M, N = AN.shape[0], BN.shape[0]
    KN, KS, KO = AN.shape[1] * 2, AS.shape[1] * 4 // 3, AO.shape[1]
    fa, fb = _segment_formats(AS, BS, AO, BO)
    segs = [(AN, SFAN, BN, SFBN, KN, fa[0], fb[0]), (AS, SFAS, BS, SFBS, KS, fa[1], fb[1]),
            (AO, SFAO, BO, SFBO, KO, fa[2], fb[2])]
    dt = np.float64 if f64 else np.float32
    acc = np.zeros((M, N), dtype=dt)
    d_bits = np.zeros((M, N), dtype=np.uint16)  # torch::zeros C (bindings.cpp:72)
    for (a, sfa, b, sfb, k, f_a, f_b) in segs:
        if k == 0:
            continue
        af = dequant(a, sfa, M, k, f_a).astype(dt)
        bf = dequant(b, sfb, N, k, f_b).astype(dt)
        part = af @ bf.T
        if chain:
            d_bits = f32_to_bf16_bits((part + bf16_bits_to_f32(d_bits).astype(dt)).astype(np.float32))
        else:
            acc += part
    if not chain:
        d_bits = f32_to_bf16_bits(acc.astype(np.float32))
    return d_bits


def fake_quant_linear(x_bits, w_bits, idx, KN, KS, KO, chain=True):
    """The whole hot path on the CPU: quantize x, quantize w (MXFP4), mixed matmul.  bench.py's cpu_baseline."""
    a = reorder_quantize(x_bits, idx, KN, KS, KO, "x")
    b = reorder_quantize(w_bits, idx, KN, KS, KO, "w4")
    return matmul(a[0], b[0], a[1], b[1], a[2], b[2], a[3], b[3], a[4], b[4], a[5], b[5], chain=chain)


# ---------------------------------------------------------------------------------------------------- calibration
def calibrate_reference(calls, lamda=1.0):
    """/root/reference/reorder_indices.py:40-113 restated literally for ONE linear input: `calls` is the list of
    activation tensors its forward hook saw.  Keeps and concatenates every |x| row, exactly as the reference does.
    Returns (reorder_index, p8_num, p6_num, p4_num) -- p4_num unguarded, as in the reference."""
    import math
    import torch
    act_scale, total = None, []
    for tensor in calls:
        hidden_dim = tensor.shape[-1]
        tensor = tensor.view(-1, hidden_dim).float().detach().cpu().abs()      # :42
        comming_scales = torch.mean(tensor, dim=0).float()                      # :43
        if act_scale is not None:
            total.append(tensor)
            act_scale = torch.max(act_scale, comming_scales)                    # :45-46
        else:
            total = [tensor]
            act_scale = comming_scales
    _, sorted_index = torch.sort(act_scale, descending=False)                   # :66
    value = torch.cat(total, dim=0)                                             # :100
    _, in_features = value.shape
    p4_threshold = value.max(dim=-1, keepdim=True)[0] * 448 / 6 / math.pow(2, 10) * lamda    # :103
    p6_threshold = value.max(dim=-1, keepdim=True)[0] * 448 / 28 / math.pow(2, 6) * lamda    # :104
    p4_ratio = (value < p4_threshold).sum() / value.numel()                     # :106
    p6_ratio = (value < p6_threshold).sum() / value.numel() - p4_ratio          # :107
    p8_ratio = 1 - p4_ratio - p6_ratio                                          # :108
    p6_num = math.ceil(in_features * p6_ratio / 128) * 128                      # :109
    p8_num = math.ceil(in_features * p8_ratio / 128) * 128                      # :110
    p4_num = in_features - p8_num - p6_num                                      # :111
    return sorted_index, p8_num, p6_num, p4_num
