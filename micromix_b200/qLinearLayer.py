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

        self.register_buffer('reorder_index', reorder_index.to(torch.int16).cuda().contiguous(), persistent=False)
        w = originalLayer.weight.data.to(device='cuda', dtype=torch.bfloat16).contiguous()
        (self.BN, self.BS, self.BO, self.SFBN, self.SFBS, self.SFBO) = mixedgemm.reorder_quantize_w4(
            w, self.reorder_index, self.p4_num, self.p6_num, self.p8_num)
        del w

    @torch.no_grad()
    def forward(self, x):
        bsz, q_len, _ = x.shape
        x = x.reshape(bsz * q_len, -1).contiguous()
        AN, AS, AO, SFAN, SFAS, SFAO = mixedgemm.reorder_quantize_x(
            x, self.reorder_index, self.p4_num, self.p6_num, self.p8_num)
        y = mixedgemm.matmul(AN, self.BN, AS, self.BS, AO, self.BO, SFAN, self.SFBN, SFAS, self.SFBS, SFAO,
                             self.SFBO, bias=self.bias)
        return y.reshape(bsz, q_len, -1)
