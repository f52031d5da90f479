"""QLinearLayer -- drop-in for /root/reference/model/qLinearLayer.py:20-74.

Same constructor and forward signature.  `__init__` quantizes the weight once to MXFP4 with the activation's channel
permutation (reference :50, mixedgemm.reorder_quantize_w4); `forward` runs the hot path:
reorder+quantize the activation (:67) then the three-segment mixed GEMM (:68), bias fused in the GEMM epilogue with
the same rounding as the reference's separate add (:70-71).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import mixedgemm

PACKED_NAMES = ("BN", "BS", "BO", "SFBN", "SFBS", "SFBO")  # the reference's attribute names, qLinearLayer.py:50-56
PACKED_FORMAT = "micromix_b200.packed_linear.v1"


def find_qlinear_layers(module, name=''):
    """qLinearLayer.py:8-17 (the reference tests a non-existent `enable_quant` attribute; every layer counts here)."""
    if type(module) == QLinearLayer:
        return {name: module}
    res = {}
    for name1, child in module.named_children():
        res.update(find_qlinear_layers(child, name=name + '.' + name1 if name != '' else name1))
    return res


class QLinearLayer(nn.Module):
    def __init__(
        self,
        originalLayer: nn.Linear,
        p8_num,
        p6_num,
        reorder_index,
        out_reorder_index=None,
    ):
        super().__init__()
        self.in_features = originalLayer.in_features
        self.out_features = originalLayer.out_features

        if originalLayer.bias is not None:
            self.register_buffer('bias', originalLayer.bias.detach().to(torch.bfloat16).cuda().contiguous())
        else:
            self.bias = None

        self.p6_num = int(p6_num)  # p4_num, p6_num, p8_num must be multiples of 128 (reference :40)
        self.p8_num = int(p8_num)
        self.p4_num = self.in_features - self.p8_num - self.p6_num
        if self.p4_num < 0 or self.p4_num % 128 or self.p6_num % 128 or self.p8_num % 128:
            raise ValueError(f"p4/p6/p8 = {self.p4_num}/{self.p6_num}/{self.p8_num} must be non-negative multiples "
                             f"of 128 summing to in_features={self.in_features}")

        self.register_buffer('reorder_index', reorder_index.to(torch.int16).cuda().contiguous())
        w = originalLayer.weight.data.to(device='cuda', dtype=torch.bfloat16).contiguous()
        packed = mixedgemm.reorder_quantize_w4(w, self.reorder_index, self.p4_num, self.p6_num, self.p8_num)
        del w
        # the six packed-weight tensors are BUFFERS: they follow .to() / .cuda() and appear in state_dict(), so a model
        # of QLinearLayers can be saved once and reloaded without re-quantizing (the reference re-quantizes every weight at
        # every start, model/qLinearLayer.py:50, and saves nothing: SURVEY.md section 5.4)
        for name, t in zip(PACKED_NAMES, packed):
            self.register_buffer(name, t)

    # ---- packed-weight checkpoint: six uint8 tensors + index + split
    def packed_state(self):
        st = {"format": PACKED_FORMAT, "in_features": self.in_features, "out_features": self.out_features,
              "p4_num": self.p4_num, "p6_num": self.p6_num, "p8_num": self.p8_num,
              "reorder_index": self.reorder_index.detach().cpu(),
              "bias": None if self.bias is None else self.bias.detach().cpu()}
        for name in PACKED_NAMES:
            st[name] = getattr(self, name).detach().cpu()
        return st

    def save_packed(self, path):
        torch.save(self.packed_state(), path)

    @classmethod
    def from_packed(cls, state, device="cuda"):
        """Rebuild a layer from packed_state() (or the path of a save_packed() file) WITHOUT touching the bf16 weight:
        start-up cost is one host->device copy of ~0.53 bytes per weight."""
        if not isinstance(state, dict):
            state = torch.load(state, map_location="cpu", weights_only=True)
        if state.get("format") != PACKED_FORMAT:
            raise ValueError(f"not a {PACKED_FORMAT} checkpoint: format={state.get('format')!r}")
        K, N = int(state["in_features"]), int(state["out_features"])
        p4, p6, p8 = int(state["p4_num"]), int(state["p6_num"]), int(state["p8_num"])
        if p4 + p6 + p8 != K or min(p4, p6, p8) < 0 or p4 % 128 or p6 % 128 or p8 % 128:
            raise ValueError(f"inconsistent split {p4}/{p6}/{p8} for in_features={K}")
        shapes = {"BN": (N, p4 // 2), "BS": (N, p6 // 2), "BO": (N, p8 // 2)}
        for name, k in (("SFBN", p4), ("SFBS", p6), ("SFBO", p8)):
            shapes[name] = (-(-N // 128) * 128 * k // 32,)
        self = cls.__new__(cls)
        nn.Module.__init__(self)
        self.in_features, self.out_features = K, N
        self.p4_num, self.p6_num, self.p8_num = p4, p6, p8
        if state.get("bias") is not None:
            self.register_buffer("bias", state["bias"].to(device=device, dtype=torch.bfloat16).contiguous())
        else:
            self.bias = None
        idx = state["reorder_index"]
        if idx.dtype != torch.int16 or idx.numel() != K:
            raise ValueError("reorder_index must be int16 [in_features]")
        self.register_buffer("reorder_index", idx.to(device).contiguous())
        for name in PACKED_NAMES:
            t = state[name]
            if t.dtype != torch.uint8 or tuple(t.shape) != shapes[name]:
                raise ValueError(f"{name}: expected uint8 {shapes[name]}, got {t.dtype} {tuple(t.shape)}")
            self.register_buffer(name, t.to(device).contiguous())
        return self

    @torch.no_grad()
    def forward(self, x, residual=None):
        """`residual` (extension, bf16 [bsz, q_len, out_features]): returns residual + linear(x), added in the GEMM's
        epilogue with the rounding of the separate torch add."""
        bsz, q_len, _ = x.shape
        x = x.reshape(bsz * q_len, -1).contiguous()
        AN, AS, AO, SFAN, SFAS, SFAO = mixedgemm.reorder_quantize_x(
            x, self.reorder_index, self.p4_num, self.p6_num, self.p8_num)
        if residual is not None:
            residual = residual.reshape(bsz * q_len, -1).contiguous()
        y = mixedgemm.matmul(AN, self.BN, AS, self.BS, AO, self.BO, SFAN, self.SFBN, SFAS, self.SFBS, SFAO,
                             self.SFBO, bias=self.bias, residual=residual)
        return y.reshape(bsz, q_len, -1)
