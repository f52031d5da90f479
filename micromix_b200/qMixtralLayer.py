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


def route_tables(sel: torch.Tensor, local_slot: torch.Tensor, n_local: int, tile: int):
    """Device-side routing tables of the grouped path -- pure tensor code, NO host synchronisation.

    sel int64 [T, k] expert ids; local_slot int64 [E]: position of expert e among this rank's experts, or n_local for an
    expert that lives elsewhere.  The (token, slot) pairs of local experts are sorted by expert (stable: tokens ascending
    inside an expert, the order torch.where gives the reference's loop) and each expert's rows are padded to whole m-tiles.
    Returns, for the static upper bound Mp = roundup(T*k + n_local*(tile-1), tile) of padded rows:
      row_src  int32 [Mp]      token whose activation row sits at padded row r (padding rows: token 0)
      pair_row int32 [T, k]    padded row holding the pair's expert output, -1 = expert not on this rank
      grp_rowblk int32 [Mp/128], grp_mtile int32 [Mp/tile]   expert slot per 128-row block / m-tile (-1 = unused m-tile)
      rows_used int32 [1]      padded rows in use (<= Mp); the quantizers skip the rest, the GEMM skips the -1 m-tiles
    """
    T, k = sel.shape
    dev = sel.device
    flat = sel.reshape(-1)
    key = local_slot[flat]                                    # [T*k], n_local = "not mine"
    order = torch.argsort(key, stable=True)
    ks = key[order]
    # (not torch.bincount: it reads the maximum back to the host -- a synchronisation, and illegal during graph capture)
    counts = torch.zeros(n_local + 1, dtype=torch.int64, device=dev).scatter_add_(0, key, torch.ones_like(key))[:n_local]
    padded = (counts + tile - 1) // tile * tile
    ends_p = torch.cumsum(padded, 0)
    starts_p = ends_p - padded
    starts_u = torch.cumsum(counts, 0) - counts
    Mp = (T * k + n_local * (tile - 1) + tile - 1) // tile * tile
    pos = torch.arange(T * k, device=dev)
    valid = ks < n_local
    ksc = ks.clamp(max=n_local - 1)
    dst = torch.where(valid, starts_p[ksc] + pos - starts_u[ksc], torch.full_like(pos, Mp))  # Mp = trash slot
    row_src = torch.zeros(Mp + 1, dtype=torch.int32, device=dev)
    row_src[dst] = torch.div(order, k, rounding_mode="floor").to(torch.int32)
    pair_row = torch.full((T * k + 1,), -1, dtype=torch.int32, device=dev)
    pair_row[torch.where(valid, order, torch.full_like(order, T * k))] = dst.to(torch.int32)
    blk = torch.arange(Mp // 128, device=dev) * 128
    grp_rowblk = torch.searchsorted(ends_p, blk, right=True).clamp(max=n_local - 1).to(torch.int32)
    mt = torch.arange(Mp // tile, device=dev) * tile
    g = torch.searchsorted(ends_p, mt, right=True)
    grp_mtile = torch.where(g < n_local, g, torch.full_like(g, -1)).to(torch.int32)
    rows_used = ends_p[-1:].to(torch.int32)  # padded rows that actually exist (the kernels skip everything past them)
    return row_src[:Mp], pair_row[:T * k].view(T, k), grp_rowblk, grp_mtile, Mp, rows_used


class QMixtralSparseMoeBlock(nn.Module):
    """qMixtralLayer.py:393-452.  `grouped=True` (default whenever this rank's experts share one (p4, p6, p8) per input):
    the experts run as ONE grouped quantize + ONE grouped GEMM per projection over the expert-sorted token matrix, with the
    gather of the routed tokens fused into the quantizer's loads and the weighted scatter-add done by one combine kernel;
    no device->host synchronisation anywhere.  `grouped=False` is the reference's op sequence: a Python loop over experts."""

    def __init__(self, originalSparseMoeBlock, p8_nums, p6_nums, reorder_index, i, ep_group=None, fused=False, grouped=None,
                 _emulate_ep=None, act_epilogue=True):
        """_emulate_ep = (ep, rank): profiling aid -- this process plays ONE rank of an ep-way expert-parallel block (its
        experts only, no combine across ranks), so that the per-rank work can be profiled on one GPU."""
        super().__init__()
        self.num_experts = getattr(originalSparseMoeBlock, "num_experts", len(originalSparseMoeBlock.experts))
        self.top_k = originalSparseMoeBlock.top_k
        self.gate = originalSparseMoeBlock.gate
        self.ep_group = ep_group
        self.ep, self.rank = tp_info(ep_group)
        self._emulated = _emulate_ep is not None
        if self._emulated:
            self.ep, self.rank = int(_emulate_ep[0]), int(_emulate_ep[1])
        self.fused = bool(fused)
        # fused + grouped: SiLU(w1 x) * (w3 x) and the quantization of w2's operand run in the w1||w3 GEMM's epilogue
        self.act_epilogue = bool(act_epilogue) and self.fused
        self.local = [j for j in range(self.num_experts) if j % self.ep == self.rank]
        key = lambda j, n: _KEY.format(i, 'block_sparse_moe', 'experts', j, n, 'input')
        same = lambda n: all(int(p8_nums[key(j, n)]) == int(p8_nums[key(self.local[0], n)]) and
                             int(p6_nums[key(j, n)]) == int(p6_nums[key(self.local[0], n)]) for j in self.local)
        shared13 = all(torch.equal(reorder_index[key(j, 'w1')].cpu(), reorder_index[key(j, 'w3')].cpu()) and
                       int(p8_nums[key(j, 'w1')]) == int(p8_nums[key(j, 'w3')]) and
                       int(p6_nums[key(j, 'w1')]) == int(p6_nums[key(j, 'w3')]) for j in self.local)
        can_group = bool(self.local) and shared13 and same('w1') and same('w2')
        self.grouped = can_group if grouped is None else bool(grouped)
        if self.grouped and not can_group:
            raise ValueError("grouped=True needs w1/w3 to share their calibration and all local experts to share one split")
        self.experts = nn.ModuleDict()
        if not self.grouped:
            for j in self.local:
                self.experts[str(j)] = QMixtralBlockSparseTop2MLP(originalSparseMoeBlock.experts[j], p8_nums, p6_nums,
                                                                  reorder_index, i, j, fused=fused)
            return
        # ---- grouped: the experts' MXFP4 weights stacked on N, one permutation per expert
        from . import mixedgemm
        j0 = self.local[0]
        e0 = originalSparseMoeBlock.experts[j0]
        self.hidden, self.inter = e0.w1.in_features, e0.w1.out_features
        self.s13 = (self.hidden - int(p6_nums[key(j0, 'w1')]) - int(p8_nums[key(j0, 'w1')]), int(p6_nums[key(j0, 'w1')]),
                    int(p8_nums[key(j0, 'w1')]))
        self.s2 = (self.inter - int(p6_nums[key(j0, 'w2')]) - int(p8_nums[key(j0, 'w2')]), int(p6_nums[key(j0, 'w2')]),
                   int(p8_nums[key(j0, 'w2')]))
        W13, W2, idx13, idx2 = [], [], [], []
        for j in self.local:
            e = originalSparseMoeBlock.experts[j]
            i13 = reorder_index[key(j, 'w1')].to(torch.int16).cuda().contiguous()
            i2 = reorder_index[key(j, 'w2')].to(torch.int16).cuda().contiguous()
            w1 = e.w1.weight.data.to(device='cuda', dtype=torch.bfloat16)
            w3 = e.w3.weight.data.to(device='cuda', dtype=torch.bfloat16)
            w2 = e.w2.weight.data.to(device='cuda', dtype=torch.bfloat16)
            if self.fused:
                # w1 / w3 rows in w2's channel order: SiLU(w1 x) * w3 x is born permuted and quantized in place
                perm = i2.to(torch.int64)
                w1, w3 = w1[perm], w3[perm]
                W2.append(mixedgemm.downproj_quantize_w4(w2[:, perm].contiguous(), *self.s2))
            else:
                W2.append(mixedgemm.reorder_quantize_w4(w2.contiguous(), i2, *self.s2))
            if self.act_epilogue and self.inter % 128 == 0 and self.s2[0] > 0:
                w13 = mixedgemm.interleave_gate_up(w1, w3)
            else:
                self.act_epilogue = False
                w13 = torch.cat([w1, w3], 0).contiguous()
            W13.append(mixedgemm.reorder_quantize_w4(w13, i13, *self.s13))
            del w13
            idx13.append(i13)
            idx2.append(i2)
            del w1, w2, w3
        n2 = self.hidden  # rows of w2
        # (scale buffers of the row-wise quantizer are sized like activations: keep the N * K/32 bytes the GEMM reads)
        cut = lambda t, rows, k: t[: rows * k // 32]
        for name, Ws, rows, split in (("W13", W13, 2 * self.inter, self.s13), ("W2", W2, n2, self.s2)):
            for c in range(3):
                self.register_buffer(f"{name}_q{c}", torch.cat([w[c] for w in Ws], 0).contiguous(), persistent=False)
                self.register_buffer(f"{name}_s{c}", torch.cat([cut(w[3 + c], rows, split[c]) for w in Ws], 0).contiguous(),
                                     persistent=False)
        self.register_buffer("idx13", torch.stack(idx13).contiguous(), persistent=False)
        self.register_buffer("idx2", torch.stack(idx2).contiguous(), persistent=False)
        slot = torch.full((self.num_experts,), len(self.local), dtype=torch.int64)
        for s_, j in enumerate(self.local):
            slot[j] = s_
        self.register_buffer("local_slot", slot.cuda(), persistent=False)

    def _w(self, name):
        return tuple(getattr(self, f"{name}_q{c}") for c in range(3)) + tuple(getattr(self, f"{name}_s{c}") for c in range(3))

    @torch.no_grad()
    def _forward_grouped(self, x, w, sel):
        from . import mixedgemm
        T = x.shape[0]
        n_local = len(self.local)
        tile = 256 if T * self.top_k >= 256 * self.num_experts else 128
        if n_local <= mixedgemm.MOE_ROUTE_MAX_LOCAL and self.num_experts <= mixedgemm.MOE_ROUTE_MAX_EXPERTS and T > 0:
            row_src, pair_row, grp_rowblk, grp_mtile, Mp, used = mixedgemm.moe_route(sel, self.local_slot, n_local, tile)
        else:
            row_src, pair_row, grp_rowblk, grp_mtile, Mp, used = route_tables(sel, self.local_slot, n_local, tile)
        a = mixedgemm.reorder_quantize_x_grouped(x, self.idx13, grp_rowblk, *self.s13, row_src=row_src, rows=Mp, rows_used=used)
        I = self.inter
        if self.fused and self.act_epilogue:
            a2 = mixedgemm.matmul_grouped(a, self._w("W13"), grp_mtile, n_local, tile, rows_used=used, act_split=self.s2)
            y = mixedgemm.matmul_grouped(a2, self._w("W2"), grp_mtile, n_local, tile, rows_used=used)   # [Mp, hidden]
            return mixedgemm.moe_combine(y, pair_row, sel.to(torch.int32), w.contiguous())
        h = mixedgemm.matmul_grouped(a, self._w("W13"), grp_mtile, n_local, tile, rows_used=used)   # [Mp, 2 * inter]
        if self.fused:
            a2 = mixedgemm.activate_quantize_x(h[:, :I], h[:, I:], *self.s2, rows_used=used)
        else:
            act = F.silu(h[:, :I]) * h[:, I:]
            a2 = mixedgemm.reorder_quantize_x_grouped(act, self.idx2, grp_rowblk, *self.s2, rows_used=used)
        y = mixedgemm.matmul_grouped(a2, self._w("W2"), grp_mtile, n_local, tile, rows_used=used)   # [Mp, hidden]
        return mixedgemm.moe_combine(y, pair_row, sel.to(torch.int32), w.contiguous())

    @torch.no_grad()
    def forward(self, hidden_states):
        b, s, h = hidden_states.shape
        x = hidden_states.view(-1, h)
        router_logits = self.gate(x)
        w = F.softmax(router_logits, dim=1, dtype=torch.float)
        w, sel = torch.topk(w, self.top_k, dim=-1)
        w = (w / w.sum(dim=-1, keepdim=True)).to(x.dtype)
        if self.grouped:
            out = self._forward_grouped(x.contiguous(), w, sel)
            if self.ep > 1 and not self._emulated:
                dist.all_reduce(out, group=self.ep_group)
            return out.view(b, s, h), router_logits
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
        if self.ep > 1 and not self._emulated:
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
