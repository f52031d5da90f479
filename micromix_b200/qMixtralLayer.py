"""QMixtralDecoderLayer / QMixtralSparseMoeBlock -- drop-in for /root/reference/model/qMixtralLayer.py:72-519.

The reference loops over the 8 experts in Python on one GPU (:437-450) and passes a stale tuple protocol into
QLinearLayer (:507-519).  Here an expert is w1||w3 fused (one quantize + one GEMM) followed by
SiLU*mul and w2, tokens are gathered per expert once, and with `ep_group` the experts are sharded across ranks
(expert e lives on rank e % ep): activations are already replicated after attention, so every rank computes its own
experts' weighted outputs into a zero buffer and ONE all-reduce combines them -- no all-to-all.
The router (`gate`) stays bf16, as in the reference (:396,420).
"""
from __future__ import annotations

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from ._qdecoder import QAttention as QMixtralAttention  # noqa: F401
from ._qdecoder import QDecoderLayer, QGatedMLP, tp_info

_KEY = 'layers.{}.{}.{}.{}.{}.{}'  # qMixtralLayer.py:467


class QMixtralBlockSparseTop2MLP(QGatedMLP):
    """One expert: w2(act(w1 x) * w3 x) (qMixtralLayer.py:454-519)."""

    def __init__(self, originalBlock, p8_nums, p6_nums, reorder_index, layer_idx, moe_idx, fused=False):
        super().__init__(originalBlock, p8_nums, p6_nums, reorder_index, layer_idx, None, names=('w1', 'w3', 'w2'),
                         key_fmt=lambda n: _KEY.format(layer_idx, 'block_sparse_moe', 'experts', moe_idx, n, 'input'),
                         fused_act=fused)

    @torch.no_grad()
    def forward(self, x):  # x: [tokens, hidden]
        return super().forward(x.unsqueeze(0)).squeeze(0)


def group_tokens_by_expert(sel: torch.Tensor, num_experts: int):
    """sel [tokens, top_k] expert ids -> (order, tok_sorted, counts): `order` lists the flattened (token, slot) pairs
    grouped by expert with tokens ascending inside an expert (stable sort = the order torch.where(sel == e) gives the
    reference's per-expert loop, qMixtralLayer.py:437-450), tok_sorted = the token of each pair, counts[e] = pairs of
    expert e (ONE device->host read for all experts)."""
    flat = sel.reshape(-1)
    order = torch.argsort(flat, stable=True)
    counts = torch.bincount(flat, minlength=num_experts).tolist()
    tok_sorted = torch.div(order, sel.shape[-1], rounding_mode="floor")
    return order, tok_sorted, counts


class QMixtralSparseMoeBlock(nn.Module):
    def __init__(self, originalSparseMoeBlock, p8_nums, p6_nums, reorder_index, i, ep_group=None, fused=False):
        super().__init__()
        self.num_experts = getattr(originalSparseMoeBlock, "num_experts", len(originalSparseMoeBlock.experts))
        self.top_k = originalSparseMoeBlock.top_k
        self.gate = originalSparseMoeBlock.gate
        self.ep_group = ep_group
        self.ep, self.rank = tp_info(ep_group)
        self.experts = nn.ModuleDict()
        for j in range(self.num_experts):
            if j % self.ep == self.rank:
                self.experts[str(j)] = QMixtralBlockSparseTop2MLP(originalSparseMoeBlock.experts[j], p8_nums, p6_nums,
                                                                  reorder_index, i, j, fused=fused)

    @torch.no_grad()
    def forward(self, hidden_states):
        b, s, h = hidden_states.shape
        x = hidden_states.view(-1, h)
        router_logits = self.gate(x)
        w = F.softmax(router_logits, dim=1, dtype=torch.float)
        w, sel = torch.topk(w, self.top_k, dim=-1)
        w = (w / w.sum(dim=-1, keepdim=True)).to(x.dtype)
        out = torch.zeros_like(x)
        # group the (token, slot) pairs by expert ONCE: a stable sort keeps tokens ascending inside an expert (the order
        # torch.where gives the reference's loop, qMixtralLayer.py:437-450) and one host read of the counts replaces a
        # device->host synchronisation per expert
        order, tok_sorted, counts = group_tokens_by_expert(sel, self.num_experts)
        w_sorted = w.reshape(-1)[order]
        off = 0
        for j in range(self.num_experts):
            n, lo = counts[j], off
            off += n
            expert = self.experts[str(j)] if str(j) in self.experts else None
            if expert is None or n == 0:
                continue
            tok = tok_sorted[lo:lo + n]
            y = expert(x.index_select(0, tok)) * w_sorted[lo:lo + n, None]
            out.index_add_(0, tok, y.to(x.dtype))
        if self.ep > 1:
            dist.all_reduce(out, group=self.ep_group)
        return out.view(b, s, h), router_logits


class QMixtralDecoderLayer(QDecoderLayer):
    def __init__(self, originalLayer, kv_cache, p8_nums, p6_nums, reorder_index, layer_idx, tp_group=None,
                 ep_group=None, fused=False):
        self._ep_group = ep_group
        super().__init__(originalLayer, kv_cache, p8_nums, p6_nums, reorder_index, layer_idx, tp_group, fused)
        self.block_sparse_moe = self.mlp

    def _build_mlp(self, originalLayer, p8_nums, p6_nums, reorder_index, layer_idx, tp_group):
        moe = getattr(originalLayer, "block_sparse_moe", None) or originalLayer.mlp
        return QMixtralSparseMoeBlock(moe, p8_nums, p6_nums, reorder_index, layer_idx, self._ep_group, fused=self.fused)

    @torch.no_grad()
    def forward(self, hidden_states, attention_mask=None, position_ids=None, past_key_value=None,
                output_attentions=False, output_router_logits=False, use_cache=False, cache_position=None,
                position_embeddings=None, **kwargs):
        """qMixtralLayer.py:119-163: (hidden_states,) [+ attn_weights] [+ present_key_value] [+ router_logits]."""
        residual = hidden_states
        hidden_states = self.input_layernorm(hidden_states)
        hidden_states, attn_weights, present = self.self_attn(
            hidden_states=hidden_states, attention_mask=attention_mask, position_ids=position_ids,
            past_key_value=past_key_value, output_attentions=output_attentions, use_cache=use_cache,
            cache_position=cache_position, position_embeddings=position_embeddings, **kwargs)
        hidden_states = residual + hidden_states
        residual = hidden_states
        hidden_states, router_logits = self.block_sparse_moe(self.post_attention_layernorm(hidden_states))
        hidden_states = residual + hidden_states
        outputs = (hidden_states,)
        if output_attentions:
            outputs += (attn_weights,)
        if use_cache:
            outputs += (present,)
        if output_router_logits:
            outputs += (router_logits,)
        return outputs
