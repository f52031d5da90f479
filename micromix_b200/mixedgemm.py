"""`mixedgemm` -- drop-in for the reference's pybind module (/root/reference/mgemm/src/bindings.cpp:682-742).

Same op names, positional/keyword argument names, return tuples, shapes and dtypes.  PyTorch is used only to
allocate the outputs and to find the current CUDA stream; all compute is in libmicromix_b200.so (hand-written
sm_100a kernels) reached through the C ABI of include/micromix_b200.h.  There is no CPU or eager fallback: inputs
that are not CUDA tensors raise, and a missing library raises at first use.

    import micromix_b200.mixedgemm as mixedgemm          # instead of sys.path.append('./mgemm/build/')

Differences from the reference, all deliberate:
  * any K = KN+KS+KO with KN, KS, KO multiples of 128 and K <= 32767 (the reference: 10 hard-coded K, :134-147);
  * inputs are validated (device, dtype, contiguity, shapes) -> ValueError/RuntimeError instead of UB;
  * kernels run on torch's current stream (the reference: legacy default stream) and are graph-capturable;
  * matmul: one launch, fp32 accumulation across all three segments, one bf16 rounding; C is not pre-zeroed;
  * rmsnorm_quantize_x implements the INTENDED semantics of the reference kernel (norm -> bf16 -> the reorder
    quantizer) for any K <= 16384; the reference's own kernel rounds to integers and mis-reduces for K != 4096;
  * the six flashinfer decode ops (bindings.cpp:418-674) are outside the hot path and absent.
"""
from __future__ import annotations

import torch

from . import _lib

__all__ = ["matmul", "reorder_quantize_x", "reorder_quantize_w", "reorder_quantize_w4", "rmsnorm_quantize_x",
           "activate_quantize_x", "downproj_quantize_w", "downproj_quantize_w4", "test_function", "launch_count",
           "reorder_quantize_x_grouped", "matmul_grouped", "moe_combine", "rope_inplace"]


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return t.data_ptr() if t is not None and t.numel() > 0 else None


def _check_cuda(name, t, dtype, ndim=None):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (micromix_b200 has no CPU path)")
    if t.dtype != dtype:
        raise ValueError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    if ndim is not None and t.dim() != ndim:
        raise ValueError(f"{name} must be {ndim}-D, got shape {tuple(t.shape)}")


def _check_split(K, KN, KS, KO):
    KN, KS, KO = int(KN), int(KS), int(KO)
    if min(KN, KS, KO) < 0 or KN + KS + KO != K:
        raise ValueError(f"KN+KS+KO must equal K={K}, got ({KN},{KS},{KO})")
    if KN % 128 or KS % 128 or KO % 128:
        raise ValueError(f"KN, KS, KO must be multiples of 128, got ({KN},{KS},{KO})")
    if K > 32767:
        raise ValueError(f"K={K} exceeds the int16 reorder_index range")
    return KN, KS, KO


def _reorder_quantize(fn_name, T, reorder_index, KN, KS, KO, widths, is_act):
    lib = _lib.load()
    _check_cuda("X" if is_act else "W", T, torch.bfloat16, 2)
    _check_cuda("reorder_index", reorder_index, torch.int16, 1)
    rows, K = T.shape
    KN, KS, KO = _check_split(K, KN, KS, KO)
    if reorder_index.numel() != K:
        raise ValueError(f"reorder_index must have K={K} entries, got {reorder_index.numel()}")
    if reorder_index.device != T.device:
        raise ValueError("reorder_index must live on the same device as the input")
    opts = dict(dtype=torch.uint8, device=T.device)
    with torch.cuda.device(T.device):
        q = [torch.empty((rows, w), **opts) for w in widths(KN, KS, KO)]
        if is_act:  # bindings.cpp:120-123
            sf = [torch.empty((int(lib.mmx_sf_bytes_act(rows, k)),), **opts) for k in (KN, KS, KO)]
        else:       # bindings.cpp:170-172 (rounded up to whole 128-row blocks so N % 128 != 0 stays in bounds)
            sf = [torch.empty((int(lib.mmx_sf_bytes_wgt(rows, k)),), **opts) for k in (KN, KS, KO)]
        rc = 0
        if rows > 0:
            rc = getattr(lib, fn_name)(_ptr(T), rows, K, _ptr(reorder_index), KN, KS, KO, _ptr(q[0]), _ptr(q[1]),
                                       _ptr(q[2]), _ptr(sf[0]), _ptr(sf[1]), _ptr(sf[2]), _stream())
    _lib.check(rc, fn_name)
    return (q[0], q[1], q[2], sf[0], sf[1], sf[2])


def reorder_quantize_x(X, reorder_index, KN, KS, KO):
    """Reorder and quantize activation (bindings.cpp:104-151): -> (XN, XS, XO, SFXN, SFXS, SFXO), all uint8."""
    return _reorder_quantize("mmx_reorder_quantize_x", X, reorder_index, KN, KS, KO,
                             lambda kn, ks, ko: (kn // 2, ks // 4 * 3, ko), True)


def reorder_quantize_w(W, reorder_index, KN, KS, KO):
    """Reorder and quantize weight, FP4|FP6|FP8 (bindings.cpp:155-202)."""
    return _reorder_quantize("mmx_reorder_quantize_w", W, reorder_index, KN, KS, KO,
                             lambda kn, ks, ko: (kn // 2, ks // 4 * 3, ko), False)


def reorder_quantize_w4(W, reorder_index, KN, KS, KO):
    """Reorder and quantize weight to MXFP4 in all three segments (bindings.cpp:206-253)."""
    return _reorder_quantize("mmx_reorder_quantize_w4", W, reorder_index, KN, KS, KO,
                             lambda kn, ks, ko: (kn // 2, ks // 2, ko // 2), False)


def matmul(AN, BN, AS, BS, AO, BO, SFAN, SFBN, SFAS, SFBS, SFAO, SFBO, bias=None, out=None, residual=None, rope=None):
    """output = A @ B^T over the three mixed-MX segments (bindings.cpp:50-102) -> bf16 [M, N].

    `bias` (bf16 [N]), `out` and `residual` are extensions: bias is added in the epilogue with the rounding of the
    reference's separate `y + self.bias` (model/qLinearLayer.py:70-71); `residual` (bf16 [M, N], may be `out`) is then added
    with the rounding of the decoder layer's separate `residual + hidden_states` (model/qLlamaLayer.py:116-158);
    `rope` = (cos, sin, rope_cols), cos / sin bf16 [S, 128]: the first rope_cols output columns are q / k heads of 128
    channels whose weight rows were stored with pair_adjacent_rows; HF's apply_rotary_pos_emb (qLlamaLayer.py:25-54) runs in
    the epilogue and every value lands at its original column -- the result equals matmul on the unpermuted weight followed
    by rope_inplace, bit for bit.  Row m uses table row m % S.
    """
    lib = _lib.load()
    names = ("AN", "BN", "AS", "BS", "AO", "BO", "SFAN", "SFBN", "SFAS", "SFBS", "SFAO", "SFBO")
    tensors = (AN, BN, AS, BS, AO, BO, SFAN, SFBN, SFAS, SFBS, SFAO, SFBO)
    for n, t in zip(names[:6], tensors[:6]):
        _check_cuda(n, t, torch.uint8, 2)
    for n, t in zip(names[6:], tensors[6:]):
        _check_cuda(n, t, torch.uint8)
    M, N = AN.size(0), BN.size(0)
    KN, KS, KO = AN.size(1) * 2, AS.size(1) * 4 // 3, AO.size(1)  # bindings.cpp:68-70
    sym = AS.size(1) == BS.size(1) and AO.size(1) == BO.size(1)    # bindings.cpp:74
    w4 = 0 if sym else 1
    if KS == 0 and KO == 0:
        w4 = 1  # both branches coincide when only the FP4 segment exists
    exp_b = (KN // 2, KS // 2 if w4 else KS // 4 * 3, KO // 2 if w4 else KO)
    for n, t, w in zip(("BN", "BS", "BO"), (BN, BS, BO), exp_b):
        if t.size(0) != N or t.size(1) != w:
            raise ValueError(f"{n} must be [{N}, {w}], got {tuple(t.shape)}")
    for n, t in zip(("AS", "AO"), (AS, AO)):
        if t.size(0) != M:
            raise ValueError(f"{n} must have M={M} rows, got {tuple(t.shape)}")
    for n, t, k in zip(("SFAN", "SFAS", "SFAO"), (SFAN, SFAS, SFAO), (KN, KS, KO)):
        if t.numel() < -(-M // 128) * 128 * k // 32:
            raise ValueError(f"{n} has {t.numel()} bytes, too small for M={M}, K={k}")
    for n, t, k in zip(("SFBN", "SFBS", "SFBO"), (SFBN, SFBS, SFBO), (KN, KS, KO)):
        if t.numel() < N * k // 32:
            raise ValueError(f"{n} has {t.numel()} bytes, too small for N={N}, K={k}")
    if bias is not None:
        _check_cuda("bias", bias, torch.bfloat16, 1)
        if bias.numel() != N:
            raise ValueError(f"bias must have N={N} entries")
    with torch.cuda.device(AN.device):
        if out is None:
            out = torch.empty((M, N), dtype=torch.bfloat16, device=AN.device)
        else:
            _check_cuda("out", out, torch.bfloat16, 2)
            if tuple(out.shape) != (M, N):
                raise ValueError(f"out must be [{M}, {N}]")
        rc = 0
        if rope is not None:
            cos, sin, rope_cols = rope
            _check_cuda("cos", cos, torch.bfloat16, 2)
            _check_cuda("sin", sin, torch.bfloat16, 2)
            if residual is not None:
                raise ValueError("rope and residual do not combine")
            if cos.shape != sin.shape or cos.shape[1] != 128 or not cos.is_contiguous() or not sin.is_contiguous():
                raise ValueError("cos / sin must be contiguous bf16 [S, 128] tables")
            if M > 0:
                rc = lib.mmx_matmul_rope(_ptr(AN), _ptr(BN), _ptr(AS), _ptr(BS), _ptr(AO), _ptr(BO), _ptr(SFAN), _ptr(SFBN),
                                         _ptr(SFAS), _ptr(SFBS), _ptr(SFAO), _ptr(SFBO), M, N, KN, KS, KO, w4, _ptr(bias),
                                         _ptr(cos), _ptr(sin), cos.shape[0], int(rope_cols), _ptr(out), _stream())
        elif residual is not None:
            _check_cuda("residual", residual, torch.bfloat16, 2)
            if tuple(residual.shape) != (M, N) or not residual.is_contiguous():
                raise ValueError(f"residual must be a contiguous [{M}, {N}] tensor")
            if M > 0:
                rc = lib.mmx_matmul_residual(_ptr(AN), _ptr(BN), _ptr(AS), _ptr(BS), _ptr(AO), _ptr(BO), _ptr(SFAN), _ptr(SFBN),
                                             _ptr(SFAS), _ptr(SFBS), _ptr(SFAO), _ptr(SFBO), M, N, KN, KS, KO, w4, _ptr(bias),
                                             _ptr(residual), _ptr(out), _stream())
        elif M > 0:
            rc = lib.mmx_matmul(_ptr(AN), _ptr(BN), _ptr(AS), _ptr(BS), _ptr(AO), _ptr(BO), _ptr(SFAN), _ptr(SFBN),
                                _ptr(SFAS), _ptr(SFBS), _ptr(SFAO), _ptr(SFBO), M, N, KN, KS, KO, w4, _ptr(bias),
                                _ptr(out), _stream())
    _lib.check(rc, "mmx_matmul")
    return out


def pair_adjacent_rows(t, heads, head_dim=128, inverse=False):
    """Row order matmul(..., rope=) expects of the q / k weights (and biases): inside every head, row 2j = channel j and row
    2j + 1 = channel j + head_dim / 2, so that the two partners of a rotary-embedding rotation are neighbouring output
    columns.  t: dim 0 = heads * head_dim channels.  inverse=True undoes it."""
    if t.shape[0] != heads * head_dim or head_dim % 2:
        raise ValueError(f"dim 0 must be heads * head_dim = {heads * head_dim}, got {tuple(t.shape)}")
    rest = t.shape[1:]
    if inverse:
        return t.reshape(heads, head_dim // 2, 2, *rest).transpose(1, 2).reshape(heads * head_dim, *rest).contiguous()
    return t.reshape(heads, 2, head_dim // 2, *rest).transpose(1, 2).reshape(heads * head_dim, *rest).contiguous()


def interleave_gate_up(gate, up, block=128):
    """Row layout matmul_activate_quantize expects of the fused weight: [gate rows of channels 0..127 | up rows of channels
    0..127 | gate 128..255 | up 128..255 | ...] (any tensors whose dim 0 is the channel; dim 0 a multiple of `block`)."""
    if gate.shape != up.shape or gate.shape[0] % block:
        raise ValueError(f"gate / up must have equal shapes with dim 0 a multiple of {block}, got {tuple(gate.shape)}, "
                         f"{tuple(up.shape)}")
    n = gate.shape[0] // block
    g = gate.reshape(n, block, *gate.shape[1:])
    u = up.reshape(n, block, *up.shape[1:])
    return torch.stack((g, u), dim=1).reshape(2 * gate.shape[0], *gate.shape[1:]).contiguous()


def matmul_activate_quantize(AN, BN, AS, BS, AO, BO, SFAN, SFBN, SFAS, SFBS, SFAO, SFBO, DN, DS, DO):
    """Extension: matmul(...) against gate / up weights interleaved per 128 channels (interleave_gate_up), followed by
    activate_quantize_x(gate, up, DN, DS, DO) -- in the GEMM's epilogue, without the bf16 [M, 2 * inter] round trip.
    -> (XN, XS, XO, SFXN, SFXS, SFXO) of the down projection's operand, bit-identical to the two separate ops."""
    lib = _lib.load()
    names = ("AN", "BN", "AS", "BS", "AO", "BO", "SFAN", "SFBN", "SFAS", "SFBS", "SFAO", "SFBO")
    tensors = (AN, BN, AS, BS, AO, BO, SFAN, SFBN, SFAS, SFBS, SFAO, SFBO)
    for n, t in zip(names[:6], tensors[:6]):
        _check_cuda(n, t, torch.uint8, 2)
    for n, t in zip(names[6:], tensors[6:]):
        _check_cuda(n, t, torch.uint8)
    M, N = AN.size(0), BN.size(0)
    KN, KS, KO = AN.size(1) * 2, AS.size(1) * 4 // 3, AO.size(1)
    sym = AS.size(1) == BS.size(1) and AO.size(1) == BO.size(1)
    w4 = 0 if sym else 1
    if KS == 0 and KO == 0:
        w4 = 1
    exp_b = (KN // 2, KS // 2 if w4 else KS // 4 * 3, KO // 2 if w4 else KO)
    for n, t, w in zip(("BN", "BS", "BO"), (BN, BS, BO), exp_b):
        if t.size(0) != N or t.size(1) != w:
            raise ValueError(f"{n} must be [{N}, {w}], got {tuple(t.shape)}")
    for n, t in zip(("AS", "AO"), (AS, AO)):
        if t.size(0) != M:
            raise ValueError(f"{n} must have M={M} rows, got {tuple(t.shape)}")
    for n, t, k in zip(("SFAN", "SFAS", "SFAO"), (SFAN, SFAS, SFAO), (KN, KS, KO)):
        if t.numel() < -(-M // 128) * 128 * k // 32:
            raise ValueError(f"{n} has {t.numel()} bytes, too small for M={M}, K={k}")
    for n, t, k in zip(("SFBN", "SFBS", "SFBO"), (SFBN, SFBS, SFBO), (KN, KS, KO)):
        if t.numel() < N * k // 32:
            raise ValueError(f"{n} has {t.numel()} bytes, too small for N={N}, K={k}")
    DN, DS, DO = _check_split(N // 2, DN, DS, DO)
    if N != 2 * (DN + DS + DO) or DN <= 0:
        raise ValueError(f"N={N} must be 2 * (DN + DS + DO) with DN > 0, got ({DN}, {DS}, {DO})")
    opts = dict(dtype=torch.uint8, device=AN.device)
    with torch.cuda.device(AN.device):
        q = [torch.empty((M, w), **opts) for w in (DN // 2, DS // 4 * 3, DO)]
        sf = [torch.empty((int(lib.mmx_sf_bytes_act(M, k)),), **opts) for k in (DN, DS, DO)]
        rc = 0
        if M > 0:
            rc = lib.mmx_matmul_activate_quantize(_ptr(AN), _ptr(BN), _ptr(AS), _ptr(BS), _ptr(AO), _ptr(BO), _ptr(SFAN),
                                                  _ptr(SFBN), _ptr(SFAS), _ptr(SFBS), _ptr(SFAO), _ptr(SFBO), M, N, KN, KS, KO,
                                                  w4, DN, DS, DO, _ptr(q[0]), _ptr(q[1]), _ptr(q[2]), _ptr(sf[0]), _ptr(sf[1]),
                                                  _ptr(sf[2]), _stream())
    _lib.check(rc, "mmx_matmul_activate_quantize")
    return (q[0], q[1], q[2], sf[0], sf[1], sf[2])


def rmsnorm_quantize_x(X, W, eps, reorder_index, KN, KS, KO):
    """RMSNorm fused into reorder+quantize (bindings.cpp:257-303): X bf16 [M, K], W bf16 [K] -> the six tensors of
    reorder_quantize_x computed on bf16((x * w) * rsqrt(mean(x^2) + eps)).  Any K <= 16384 (the reference: four K)."""
    lib = _lib.load()
    _check_cuda("X", X, torch.bfloat16, 2)
    _check_cuda("W", W, torch.bfloat16, 1)
    _check_cuda("reorder_index", reorder_index, torch.int16, 1)
    M, K = X.shape
    KN, KS, KO = _check_split(K, KN, KS, KO)
    if W.numel() != K or reorder_index.numel() != K:
        raise ValueError(f"W and reorder_index must have K={K} entries, got {W.numel()} and {reorder_index.numel()}")
    if W.device != X.device or reorder_index.device != X.device:
        raise ValueError("W and reorder_index must live on the same device as X")
    if K > 16384:
        raise ValueError(f"rmsnorm_quantize_x supports K <= 16384, got {K}")
    opts = dict(dtype=torch.uint8, device=X.device)
    with torch.cuda.device(X.device):
        q = [torch.empty((M, w), **opts) for w in (KN // 2, KS // 4 * 3, KO)]
        sf = [torch.empty((int(lib.mmx_sf_bytes_act(M, k)),), **opts) for k in (KN, KS, KO)]
        rc = 0
        if M > 0:
            rc = lib.mmx_rmsnorm_quantize_x(_ptr(X), _ptr(W), float(eps), M, K, _ptr(reorder_index), KN, KS, KO,
                                            _ptr(q[0]), _ptr(q[1]), _ptr(q[2]), _ptr(sf[0]), _ptr(sf[1]), _ptr(sf[2]),
                                            _stream())
    _lib.check(rc, "mmx_rmsnorm_quantize_x")
    return (q[0], q[1], q[2], sf[0], sf[1], sf[2])


def _rowwise(fn_name, tensors, names, KN, KS, KO, widths):
    lib = _lib.load()
    for n, t in zip(names, tensors):
        _check_cuda(n, t, torch.bfloat16, 2)
    rows, K = tensors[0].shape
    for n, t in zip(names[1:], tensors[1:]):
        if t.shape != tensors[0].shape or t.device != tensors[0].device:
            raise ValueError(f"{n} must match {names[0]} in shape and device")
    KN, KS, KO = int(KN), int(KS), int(KO)
    if min(KN, KS, KO) < 0 or KN + KS + KO != K:
        raise ValueError(f"KN+KS+KO must equal K={K}, got ({KN},{KS},{KO})")
    if KN % 128 or KS % 128 or KO % 128:
        raise ValueError(f"KN, KS, KO must be multiples of 128, got ({KN},{KS},{KO})")
    dev = tensors[0].device
    opts = dict(dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        q = [torch.empty((rows, w), **opts) for w in widths(KN, KS, KO)]
        # all three ops size their scale buffers like activations (bindings.cpp:320-322, 346-348, 373-375)
        sf = [torch.empty((int(lib.mmx_sf_bytes_act(rows, k)),), **opts) for k in (KN, KS, KO)]
        rc = 0
        if rows > 0:
            rc = getattr(lib, fn_name)(*[_ptr(t) for t in tensors], rows, KN, KS, KO, _ptr(q[0]), _ptr(q[1]),
                                       _ptr(q[2]), _ptr(sf[0]), _ptr(sf[1]), _ptr(sf[2]), _stream())
    _lib.check(rc, fn_name)
    return (q[0], q[1], q[2], sf[0], sf[1], sf[2])


def activate_quantize_x(A, B, KN, KS, KO, rows_used=None):
    """SiLU(A) * B -> MX quantize without a permutation (bindings.cpp:307-334): A = gate, B = up, both bf16 [M, K]
    already in down_proj's channel order -> (XN, XS, XO, SFXN, SFXS, SFXO).
    Extension (rows_used: int32 [1] on the device = rows that exist, strided inputs only; see reorder_quantize_x_grouped).
    Extension: A and B may be column slices of one wider row-major matrix (same row stride, unit column stride), e.g. the
    two halves of a fused gate_up GEMM output; they are then read in place."""
    if (A.dim() == 2 and B.dim() == 2 and not (A.is_contiguous() and B.is_contiguous()) and A.shape == B.shape
            and A.stride(1) == 1 and B.stride(1) == 1 and A.stride(0) == B.stride(0) and A.stride(0) % 8 == 0
            and A.is_cuda and B.is_cuda and A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16
            and A.data_ptr() % 16 == 0 and B.data_ptr() % 16 == 0):
        lib = _lib.load()
        rows, K = A.shape
        KN, KS, KO = _check_split(K, KN, KS, KO)
        opts = dict(dtype=torch.uint8, device=A.device)
        with torch.cuda.device(A.device):
            q = [torch.empty((rows, w), **opts) for w in (KN // 2, KS // 4 * 3, KO)]
            sf = [torch.empty((int(lib.mmx_sf_bytes_act(rows, k)),), **opts) for k in (KN, KS, KO)]
            rc = 0
            if rows > 0 and rows_used is not None:  # grouped MoE: the used row count lives on the device
                rc = lib.mmx_activate_quantize_x_rows(A.data_ptr(), B.data_ptr(), A.stride(0), rows, _ptr(rows_used), KN, KS, KO,
                                                      _ptr(q[0]), _ptr(q[1]), _ptr(q[2]), _ptr(sf[0]), _ptr(sf[1]), _ptr(sf[2]),
                                                      _stream())
            elif rows > 0:
                rc = lib.mmx_activate_quantize_x_strided(A.data_ptr(), B.data_ptr(), A.stride(0), rows, KN, KS, KO, _ptr(q[0]),
                                                         _ptr(q[1]), _ptr(q[2]), _ptr(sf[0]), _ptr(sf[1]), _ptr(sf[2]), _stream())
        _lib.check(rc, "mmx_activate_quantize_x_strided")
        return (q[0], q[1], q[2], sf[0], sf[1], sf[2])
    return _rowwise("mmx_activate_quantize_x", (A, B), ("A", "B"), KN, KS, KO,
                    lambda kn, ks, ko: (kn // 2, ks // 4 * 3, ko))


def downproj_quantize_w(W, KN, KS, KO):
    """Quantize already-ordered weight rows to FP4|FP6|FP8 (bindings.cpp:336-360)."""
    return _rowwise("mmx_downproj_quantize_w", (W,), ("W",), KN, KS, KO,
                    lambda kn, ks, ko: (kn // 2, ks // 4 * 3, ko))


def downproj_quantize_w4(W, KN, KS, KO):
    """Quantize already-ordered weight rows to MXFP4 in all three segments (bindings.cpp:362-387)."""
    return _rowwise("mmx_downproj_quantize_w4", (W,), ("W",), KN, KS, KO,
                    lambda kn, ks, ko: (kn // 2, ks // 2, ko // 2))


# ---------------------------------------------------------------------------------------------- grouped forms (Mixtral)
def reorder_quantize_x_grouped(X, reorder_index, group_of_rowblock, KN, KS, KO, row_src=None, rows=None, rows_used=None):
    """Extension (the reference loops over experts in Python, qMixtralLayer.py:437-450): ONE quantize launch over the
    expert-sorted, per-expert-padded token matrix.  reorder_index int16 [groups, K]; group_of_rowblock int32 [rows/128]
    (device); row_src int32 [rows] (device, optional): sorted row r is X[row_src[r]] -- the gather is fused; `rows` = rows
    of the sorted matrix (defaults to X.size(0)); rows_used int32 [1] (device, optional): rows that actually exist (a multiple
    of 128) -- `rows` is then only the static upper bound and row blocks past rows_used are not touched.
    -> the six tensors of reorder_quantize_x for the sorted matrix."""
    lib = _lib.load()
    _check_cuda("X", X, torch.bfloat16, 2)
    _check_cuda("reorder_index", reorder_index, torch.int16, 2)
    _check_cuda("group_of_rowblock", group_of_rowblock, torch.int32, 1)
    K = X.size(1)
    M = int(rows) if rows is not None else X.size(0)
    KN, KS, KO = _check_split(K, KN, KS, KO)
    if reorder_index.size(1) != K:
        raise ValueError(f"reorder_index must be [groups, {K}]")
    if M % 128 or group_of_rowblock.numel() < M // 128:
        raise ValueError("the sorted matrix must be padded to whole 128-row blocks, one group id per block")
    if row_src is not None:
        _check_cuda("row_src", row_src, torch.int32, 1)
        if row_src.numel() < M:
            raise ValueError(f"row_src must have {M} entries")
    elif X.size(0) != M:
        raise ValueError("without row_src, X must be the sorted matrix itself")
    opts = dict(dtype=torch.uint8, device=X.device)
    with torch.cuda.device(X.device):
        q = [torch.empty((M, w), **opts) for w in (KN // 2, KS // 4 * 3, KO)]
        sf = [torch.empty((int(lib.mmx_sf_bytes_act(M, k)),), **opts) for k in (KN, KS, KO)]
        rc = 0
        if M > 0:
            rc = lib.mmx_reorder_quantize_x_grouped(_ptr(X), M, K, _ptr(reorder_index), _ptr(group_of_rowblock), _ptr(row_src),
                                                    _ptr(rows_used), KN, KS, KO, _ptr(q[0]), _ptr(q[1]), _ptr(q[2]), _ptr(sf[0]), _ptr(sf[1]),
                                                    _ptr(sf[2]), _stream())
    _lib.check(rc, "mmx_reorder_quantize_x_grouped")
    return (q[0], q[1], q[2], sf[0], sf[1], sf[2])


def matmul_grouped(A, W, group_of_mtile, groups, tile_rows, out=None, rows_used=None, act_split=None):
    """Extension: ONE persistent mixed GEMM over all experts.  A = the six tensors of the sorted, padded activation
    ([M, *]); W = six tensors with the experts' MXFP4 weights stacked on N ([groups * N, *]); group_of_mtile int32
    [M / tile_rows] (device): expert of each m-tile, -1 = padding tile (skipped); rows_used int32 [1] (device, optional):
    rows that exist -- the tile walk stops there -> bf16 [M, N].
    act_split = (DN, DS, DO): every expert's rows are interleave_gate_up(w1, w3) and the epilogue emits the MX-quantized
    SiLU(w1 x) * (w3 x) instead -> the six operand tensors of the grouped w2 GEMM (see matmul_activate_quantize)."""
    lib = _lib.load()
    for t in list(A) + list(W):
        _check_cuda("operand", t, torch.uint8)
    _check_cuda("group_of_mtile", group_of_mtile, torch.int32, 1)
    M = A[0].size(0)
    KN, KS, KO = A[0].size(1) * 2, A[1].size(1) * 4 // 3, A[2].size(1)
    if W[0].size(0) % groups:
        raise ValueError("the stacked weights must hold `groups` equal blocks of rows")
    N = W[0].size(0) // groups
    if tile_rows not in (128, 256) or M % tile_rows or group_of_mtile.numel() < M // tile_rows:
        raise ValueError("M must be padded to whole m-tiles (128 | 256 rows), one group id per m-tile")
    if act_split is not None:
        DN, DS, DO = _check_split(N // 2, *act_split)
        opts = dict(dtype=torch.uint8, device=A[0].device)
        with torch.cuda.device(A[0].device):
            q = [torch.empty((M, w), **opts) for w in (DN // 2, DS // 4 * 3, DO)]
            sf = [torch.empty((int(lib.mmx_sf_bytes_act(M, k)),), **opts) for k in (DN, DS, DO)]
            rc = 0
            if M > 0:
                rc = lib.mmx_matmul_grouped_activate_quantize(
                    _ptr(A[0]), _ptr(W[0]), _ptr(A[1]), _ptr(W[1]), _ptr(A[2]), _ptr(W[2]), _ptr(A[3]), _ptr(W[3]), _ptr(A[4]),
                    _ptr(W[4]), _ptr(A[5]), _ptr(W[5]), M, N, KN, KS, KO, 1, int(groups), int(tile_rows), _ptr(group_of_mtile),
                    _ptr(rows_used), DN, DS, DO, _ptr(q[0]), _ptr(q[1]), _ptr(q[2]), _ptr(sf[0]), _ptr(sf[1]), _ptr(sf[2]),
                    _stream())
        _lib.check(rc, "mmx_matmul_grouped_activate_quantize")
        return (q[0], q[1], q[2], sf[0], sf[1], sf[2])
    with torch.cuda.device(A[0].device):
        if out is None:
            out = torch.empty((M, N), dtype=torch.bfloat16, device=A[0].device)
        rc = 0
        if M > 0:
            rc = lib.mmx_matmul_grouped(_ptr(A[0]), _ptr(W[0]), _ptr(A[1]), _ptr(W[1]), _ptr(A[2]), _ptr(W[2]), _ptr(A[3]),
                                        _ptr(W[3]), _ptr(A[4]), _ptr(W[4]), _ptr(A[5]), _ptr(W[5]), M, N, KN, KS, KO, 1,
                                        int(groups), int(tile_rows), _ptr(group_of_mtile), _ptr(rows_used), _ptr(out),
                                        _stream())
    _lib.check(rc, "mmx_matmul_grouped")
    return out


MOE_ROUTE_MAX_LOCAL = 64
MOE_ROUTE_MAX_EXPERTS = 256


def moe_route(sel, local_slot, n_local, tile):
    """Extension: the routing tables of the grouped expert path in ONE kernel (qMixtralLayer.route_tables is the same
    function in torch ops, ~20 launches): sel int64 [T, k], local_slot int64 [experts] (slot on this rank or n_local) ->
    (row_src int32 [Mp], pair_row int32 [T, k], grp_rowblk int32 [Mp/128], grp_mtile int32 [Mp/tile], Mp, rows_used int32 [1])."""
    lib = _lib.load()
    _check_cuda("sel", sel, torch.int64, 2)
    _check_cuda("local_slot", local_slot, torch.int64, 1)
    T, k = sel.shape
    if not (1 <= n_local <= MOE_ROUTE_MAX_LOCAL) or tile not in (128, 256):
        raise ValueError(f"n_local must be 1..{MOE_ROUTE_MAX_LOCAL} and tile 128 | 256")
    Mp = (T * k + n_local * (tile - 1) + tile - 1) // tile * tile
    i32 = dict(dtype=torch.int32, device=sel.device)
    with torch.cuda.device(sel.device):
        row_src, pair_row = torch.empty(Mp, **i32), torch.empty((T, k), **i32)
        grp_rowblk, grp_mtile, used = torch.empty(Mp // 128, **i32), torch.empty(Mp // tile, **i32), torch.empty(1, **i32)
        rc = lib.mmx_moe_route(_ptr(sel.contiguous()), _ptr(local_slot), T, k, int(local_slot.numel()), int(n_local), int(tile),
                               Mp, _ptr(row_src),
                               pair_row.data_ptr(), _ptr(grp_rowblk), _ptr(grp_mtile), _ptr(used), _stream())
    _lib.check(rc, "mmx_moe_route")
    return row_src, pair_row, grp_rowblk, grp_mtile, Mp, used


def moe_combine(Y, row, expert, weight, out=None):
    """out[t] = sum over token t's slots, ascending expert id, of bf16(Y[row[t,s]] * weight[t,s]) with a bf16 rounding
    after every add (the reference's index_add_ loop, qMixtralLayer.py:446-450); row < 0 = expert not on this rank."""
    lib = _lib.load()
    _check_cuda("Y", Y, torch.bfloat16, 2)
    _check_cuda("row", row, torch.int32, 2)
    _check_cuda("expert", expert, torch.int32, 2)
    _check_cuda("weight", weight, torch.bfloat16, 2)
    T, k = row.shape
    H = Y.size(1)
    with torch.cuda.device(Y.device):
        if out is None:
            out = torch.empty((T, H), dtype=torch.bfloat16, device=Y.device)
        rc = lib.mmx_moe_combine(_ptr(Y), _ptr(row), _ptr(expert), _ptr(weight), T, k, H, _ptr(out), _stream()) if T else 0
    _lib.check(rc, "mmx_moe_combine")
    return out


def rope_inplace(Y, heads, head_dim, cos, sin):
    """Extension: HF rotary embedding (q * cos + rotate_half(q) * sin, every op rounded to bf16 like the torch ops of
    qLlamaLayer.py:25-54) applied IN PLACE to the first heads * head_dim columns of the rows of Y (bf16 [M, ld], unit column
    stride) -- the q and k columns of a fused qkv output.  cos / sin: bf16 [S, head_dim] with M % S == 0 (row m uses m % S)."""
    lib = _lib.load()
    if Y.dim() != 2 or Y.dtype != torch.bfloat16 or not Y.is_cuda or Y.stride(1) != 1:
        raise ValueError("Y must be a CUDA bf16 [M, ld] matrix with unit column stride")
    _check_cuda("cos", cos, torch.bfloat16, 2)
    _check_cuda("sin", sin, torch.bfloat16, 2)
    M, S = Y.size(0), cos.size(0)
    if cos.shape != sin.shape or cos.size(1) != head_dim or M % S or heads * head_dim > Y.size(1):
        raise ValueError("cos / sin must be [S, head_dim] with M % S == 0, and heads * head_dim columns must exist")
    with torch.cuda.device(Y.device):
        rc = lib.mmx_rope_inplace(Y.data_ptr(), Y.stride(0), M, int(heads), int(head_dim), cos.data_ptr(), sin.data_ptr(), S,
                                  _stream()) if M else 0
    _lib.check(rc, "mmx_rope_inplace")
    return Y


def test_function():
    return "Hello from test_function!"


def launch_count() -> int:
    """Kernels launched by the library so far (bench.py's gpu_launches)."""
    return int(_lib.load().mmx_launch_count())
