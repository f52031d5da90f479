"""Shared pieces of the quantized decoder-layer wrappers (qLlamaLayer / qQwenLayer / qMixtralLayer).

The reference wraps every `nn.Linear` of a HF decoder layer in a `QLinearLayer` and keeps attention / RoPE / norms in
stock PyTorch (/root/reference/model/qLlamaLayer.py:69-387, qQwenLayer.py, qMixtralLayer.py).  The B200 version keeps the
constructors and forward signatures and changes what the hot path costs:

  * q/k/v (and gate/up) read the same activation with the same reorder_index and split, so they are ONE QLinearLayer
    over the row-concatenated weight: one reorder+quantize and one mixed GEMM instead of three (two);
  * `fused=True` (extension, off by default = the reference's op sequence): the two RMSNorms run inside the quantizer
    (`rmsnorm_quantize_x`: no normalised [M, hidden] round trip through HBM) and SiLU(gate) * up + the quantization of
    down_proj's operand run in the EPILOGUE of the gate_up GEMM (`matmul_activate_quantize`: the [M, 2 * intermediate] bf16
    result never exists in HBM; bit-identical to `matmul` followed by the reference's `activate_quantize_x`, which remains
    the path when gate / up have biases or different calibrations); for that the rows of gate_proj / up_proj are stored in
    down_proj's channel order, interleaved per 128 channels, at construction, so the intermediate activation is born
    permuted (the reference's abandoned `out_reorder_index` idea, qLinearLayer.py:27, qLlamaLayer.py:341,354);
  * with a `tp_group`, qkv / gate_up are column-parallel (no collective) and o / down row-parallel with an NCCL all-reduce
    (micromix_b200.parallel_utils); heads are sharded so every rank runs attention on its own heads only.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from . import mixedgemm
from .parallel_utils import RowParallelQLinear, TokenParallelQLinear, forward_row_shard
from .qLinearLayer import QLinearLayer

NAME = 'layers.{}.{}.{}.{}'  # the key template of the reference's calibration dicts (qLlamaLayer.py:211)


def _meta_linear(weight: torch.Tensor, bias: Optional[torch.Tensor]) -> nn.Linear:
    lin = nn.Linear(weight.shape[1], weight.shape[0], bias=bias is not None, device="meta", dtype=torch.bfloat16)
    lin.weight = nn.Parameter(weight, requires_grad=False)
    if bias is not None:
        lin.bias = nn.Parameter(bias, requires_grad=False)
    return lin


def _same(a, b) -> bool:
    if isinstance(a, torch.Tensor):
        return isinstance(b, torch.Tensor) and a.shape == b.shape and bool(torch.equal(a.cpu(), b.cpu()))
    return int(a) == int(b)


class FusedQLinear(nn.Module):
    """Several linears that read the same activation: one quantize + one GEMM, outputs split on the last dim.

    `row_slices[i]` selects the output rows of linear i this rank owns (column parallelism); None = all rows.
    """

    def __init__(self, linears: Sequence[nn.Linear], p8_num, p6_num, reorder_index, row_slices=None):
        super().__init__()
        ws, bs, self.splits = [], [], []
        any_bias = any(l.bias is not None for l in linears)
        for i, l in enumerate(linears):
            sl = row_slices[i] if row_slices is not None else slice(None)
            w = l.weight.data[sl]
            ws.append(w)
            self.splits.append(w.shape[0])
            if any_bias:
                b = l.bias.data[sl] if l.bias is not None else torch.zeros(w.shape[0], dtype=w.dtype, device=w.device)
                bs.append(b)
        weight = torch.cat(ws, dim=0).contiguous()
        bias = torch.cat(bs, dim=0).contiguous() if any_bias else None
        self.inner = QLinearLayer(_meta_linear(weight, bias), p8_num, p6_num, reorder_index)
        self.in_features, self.out_features = weight.shape[1], weight.shape[0]

    @torch.no_grad()
    def forward_whole(self, x, norm=None, rope=None):
        """The concatenated output [b, s, sum(splits)] (one GEMM); norm = (weight bf16 [K], eps): x is the UN-normalised
        input and RMSNorm runs inside the quantizer; rope = (cos, sin, rope_cols): see mixedgemm.matmul."""
        if norm is None and rope is None:
            return self.inner(x)
        lin = self.inner
        bsz, q_len, _ = x.shape
        x2 = x.reshape(bsz * q_len, -1).contiguous()
        if norm is None:
            a = mixedgemm.reorder_quantize_x(x2, lin.reorder_index, lin.p4_num, lin.p6_num, lin.p8_num)
        else:
            a = mixedgemm.rmsnorm_quantize_x(x2, norm[0], norm[1], lin.reorder_index, lin.p4_num, lin.p6_num, lin.p8_num)
        return mixedgemm.matmul(a[0], lin.BN, a[1], lin.BS, a[2], lin.BO, a[3], lin.SFBN, a[4], lin.SFBS, a[5], lin.SFBO,
                                bias=lin.bias, rope=rope).reshape(bsz, q_len, -1)

    @torch.no_grad()
    def forward(self, x, norm=None):
        y = self.forward_whole(x, norm)
        return y.split(self.splits, dim=-1) if len(self.splits) > 1 else (y,)

    @torch.no_grad()
    def forward_activated(self, x, dsplit, norm=None):
        """This module holds gate / up rows interleaved per 128 channels (mixedgemm.interleave_gate_up): one quantize + one
        GEMM whose epilogue emits the MX-quantized SiLU(gate) * up -> the six operand tensors of the down projection."""
        lin = self.inner
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        if norm is None:
            a = mixedgemm.reorder_quantize_x(x2, lin.reorder_index, lin.p4_num, lin.p6_num, lin.p8_num)
        else:
            a = mixedgemm.rmsnorm_quantize_x(x2, norm[0], norm[1], lin.reorder_index, lin.p4_num, lin.p6_num, lin.p8_num)
        return mixedgemm.matmul_activate_quantize(a[0], lin.BN, a[1], lin.BS, a[2], lin.BO, a[3], lin.SFBN, a[4], lin.SFBS,
                                                  a[5], lin.SFBO, *dsplit)

    @torch.no_grad()
    def forward_gathered(self, x_shard, M, workspace, norm=None):
        """Sequence-parallel form: x_shard = THIS rank's rows of the [M, K] activation.  The rank quantizes only those
        (RMSNorm fused when `norm` is given), the packed codes are multicast into every rank's gather channel, and the
        column-parallel GEMM runs on the gathered operand -> [M, N/tp] splits."""
        y = self.forward_gathered_whole(x_shard, M, workspace, norm)
        return y.split(self.splits, dim=-1) if len(self.splits) > 1 else (y,)

    @torch.no_grad()
    def forward_gathered_whole(self, x_shard, M, workspace, norm=None):
        lin = self.inner
        a = workspace.quantize_allgather(x_shard, M, lin.reorder_index, lin.p4_num, lin.p6_num, lin.p8_num, norm=norm)
        del a  # (the GEMM reads the channel directly)
        return workspace.matmul_gathered(M, (lin.BN, lin.BS, lin.BO, lin.SFBN, lin.SFBS, lin.SFBO), lin.p4_num, lin.p6_num,
                                         lin.p8_num, bias=lin.bias)


def build_input_group(linears, keys, p8_nums, p6_nums, reorder_index, row_slices=None):
    """One FusedQLinear if all linears share (index, p6, p8) -- calibration gives linears with the same input the same
    statistics (reorder_indices.py:72-78) -- else one FusedQLinear per linear (the reference's behaviour)."""
    k0 = keys[0]
    shared = all(_same(reorder_index[k], reorder_index[k0]) and _same(p8_nums[k], p8_nums[k0]) and
                 _same(p6_nums[k], p6_nums[k0]) for k in keys[1:])
    if shared:
        return nn.ModuleList([FusedQLinear(linears, p8_nums[k0], p6_nums[k0], reorder_index[k0], row_slices)])
    return nn.ModuleList([FusedQLinear([l], p8_nums[k], p6_nums[k], reorder_index[k],
                                       None if row_slices is None else [row_slices[i]])
                          for i, (l, k) in enumerate(zip(linears, keys))])


def run_input_group(group, x, norm=None):
    outs = []
    for m in group:
        outs.extend(m(x, norm))
    return outs


def fusable_rmsnorm(norm):
    """(weight, eps) if `norm` is a plain RMSNorm (HF Llama / Qwen2 / Mixtral RMSNorm or model_shapes.RMSNorm), else None."""
    w = getattr(norm, "weight", None)
    if w is None or getattr(norm, "bias", None) is not None or "RMSNorm" not in type(norm).__name__:
        return None
    if w.dtype != torch.bfloat16 or not w.is_cuda:
        return None
    eps = getattr(norm, "variance_epsilon", None)
    if eps is None:
        eps = getattr(norm, "eps", None)
    return None if eps is None else (w.detach().contiguous(), float(eps))


def tp_info(group):
    if group is None or not dist.is_available() or not dist.is_initialized():
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


@torch.no_grad()
def quantize_int_group(w, nbits, group_size):
    """Asymmetric int fake-quant of the KV cache (qLlamaLayer.py:13-23), used when kv_cache=True."""
    shape = w.shape
    w = w.reshape(-1, group_size)
    w_max, w_min = w.amax(dim=-1, keepdim=True), w.amin(dim=-1, keepdim=True)
    q_max = 2 ** nbits - 1
    scales = (w_max - w_min).clamp(min=1e-5) / q_max
    base = torch.round(-w_min / scales).clamp_(min=0, max=q_max)
    w = (torch.clamp(torch.round(w / scales) + base, 0, q_max) - base) * scales
    return w.reshape(shape)


def rope_tables_2d(position_embeddings, bsz, q_len, head_dim):
    """(cos, sin) as contiguous bf16 [S, d] tables for mixedgemm.rope_inplace (row m of the flattened [b*s] tokens uses table
    row m % S), or None when the tables do not have that form (then RoPE stays in torch)."""
    cos, sin = position_embeddings
    if cos.dtype != torch.bfloat16 or not cos.is_cuda or cos.shape != sin.shape or cos.dim() != 3 or cos.shape[-1] != head_dim:
        return None
    if cos.shape[1] != q_len or cos.shape[0] not in (1, bsz):
        return None
    if cos.shape[0] == 1 or (cos.stride(0) == 0 and sin.stride(0) == 0):  # one table for the whole batch (prefill)
        c, s_ = cos[0], sin[0]
    else:
        c, s_ = cos.reshape(bsz * q_len, head_dim), sin.reshape(bsz * q_len, head_dim)
    if not (c.is_contiguous() and s_.is_contiguous()) or c.data_ptr() % 16 or s_.data_ptr() % 16:
        return None
    return c, s_


def apply_rope(q, k, cos, sin):
    """q, k: [b, heads, s, d]; cos, sin: [b, s, d] (HF convention: halves rotated, qLlamaLayer.py:25-54)."""
    cos, sin = cos.unsqueeze(1), sin.unsqueeze(1)

    def rot(x):
        h = x.shape[-1] // 2
        return torch.cat((-x[..., h:], x[..., :h]), dim=-1)

    return q * cos + rot(q) * sin, k * cos + rot(k) * sin


class QAttention(nn.Module):
    """Llama / Qwen2 / Mixtral attention with quantized projections (qLlamaLayer.py:196-321, qQwenLayer.py:205-327)."""

    def __init__(self, originalAttn, kv_cache, p8_nums, p6_nums, reorder_index, i, tp_group=None, workspace=None,
                 sequence_parallel=False, token_parallel_rows=False, rope_epilogue=False):
        """rope_epilogue (extension, single GPU): the q / k weight rows are stored pair-adjacent inside each head
        (mixedgemm.pair_adjacent_rows) and the qkv GEMM's epilogue applies the rotary embedding (matmul(..., rope=)); the
        output is bit-identical to GEMM + rope_inplace / HF's apply_rotary_pos_emb."""
        super().__init__()
        self.sp = bool(sequence_parallel)
        self.tpr = bool(token_parallel_rows) and self.sp
        self.workspace = workspace
        cfg = originalAttn.config
        self.config = cfg
        self.q_kv_cache = kv_cache
        self.layer_idx = i
        self.tp, self.rank = tp_info(tp_group)
        self.head_dim = getattr(cfg, "head_dim", None) or cfg.hidden_size // cfg.num_attention_heads
        if cfg.num_attention_heads % self.tp or cfg.num_key_value_heads % self.tp:
            raise ValueError(f"heads {cfg.num_attention_heads}/{cfg.num_key_value_heads} do not shard {self.tp} ways")
        self.num_heads = cfg.num_attention_heads // self.tp          # local
        self.num_key_value_heads = cfg.num_key_value_heads // self.tp  # local
        self.num_key_value_groups = self.num_heads // self.num_key_value_heads
        self.attention_dropout = getattr(originalAttn, "attention_dropout", 0.0)
        keys = [NAME.format(i, 'self_attn', n, 'input') for n in ('q_proj', 'k_proj', 'v_proj')]
        d = self.head_dim
        slices = None
        if self.tp > 1:
            slices = [slice(self.rank * self.num_heads * d, (self.rank + 1) * self.num_heads * d),
                      slice(self.rank * self.num_key_value_heads * d, (self.rank + 1) * self.num_key_value_heads * d),
                      slice(self.rank * self.num_key_value_heads * d, (self.rank + 1) * self.num_key_value_heads * d)]
        qkv = [originalAttn.q_proj, originalAttn.k_proj, originalAttn.v_proj]
        shared = all(_same(reorder_index[k], reorder_index[keys[0]]) and _same(p8_nums[k], p8_nums[keys[0]]) and
                     _same(p6_nums[k], p6_nums[keys[0]]) for k in keys[1:])
        self.rope_epilogue = bool(rope_epilogue) and self.tp == 1 and not self.sp and shared and d == 128
        if self.rope_epilogue:
            def paired(lin, heads):
                w = mixedgemm.pair_adjacent_rows(lin.weight.data, heads, d)
                b = None if lin.bias is None else mixedgemm.pair_adjacent_rows(lin.bias.data, heads, d)
                return _meta_linear(w, b)
            qkv = [paired(qkv[0], self.num_heads), paired(qkv[1], self.num_key_value_heads), qkv[2]]
        self.qkv_proj = build_input_group(qkv, keys, p8_nums, p6_nums, reorder_index, slices)
        ko = NAME.format(i, 'self_attn', 'o_proj', 'input')
        if self.tp > 1 and self.tpr:
            self.o_proj = TokenParallelQLinear(originalAttn.o_proj, p8_nums[ko], p6_nums[ko], reorder_index[ko], tp_group,
                                               workspace)
        elif self.tp > 1:
            self.o_proj = RowParallelQLinear(originalAttn.o_proj, p8_nums[ko], p6_nums[ko], reorder_index[ko], tp_group,
                                             workspace=workspace)
        else:
            self.o_proj = QLinearLayer(originalAttn.o_proj, p8_nums[ko], p6_nums[ko], reorder_index[ko])

    @torch.no_grad()
    def forward(self, hidden_states, attention_mask=None, position_ids=None, past_key_value=None,
                output_attentions=False, use_cache=False, cache_position=None, position_embeddings=None, **kwargs):
        past_key_value = kwargs.get("past_key_values", past_key_value)
        norm = kwargs.pop("mmx_norm", None)
        res = kwargs.pop("mmx_residual", None)  # returned as residual + o_proj(...), added in the GEMM's epilogue
        if self.sp:
            # hidden_states = this rank's token rows [1, rows, hidden]; the batch shape comes from the rotary tables
            bsz, q_len = position_embeddings[0].shape[0], position_embeddings[0].shape[1]
        else:
            bsz, q_len, _ = hidden_states.size()
        # q, k, v from ONE GEMM: the rotary embedding is applied IN PLACE to the q and k columns of that output by one
        # kernel (same arithmetic and roundings as the torch ops below) instead of ~10 strided elementwise / cat kernels
        tables = None
        if len(self.qkv_proj) == 1 and position_embeddings is not None and self.head_dim % 16 == 0:
            tables = rope_tables_2d(position_embeddings, bsz, q_len, self.head_dim)
        if self.rope_epilogue:
            m = self.qkv_proj[0]
            nrope = (self.num_heads + self.num_key_value_heads) * self.head_dim
            if tables is not None:
                y = m.forward_whole(hidden_states, norm, rope=(tables[0], tables[1], nrope))
                tables = None  # done
                position_embeddings_applied = True
            else:
                # no [S, d] tables (or no rotary embedding at all): undo the pair-adjacent column order of q / k
                y = m.forward_whole(hidden_states, norm).reshape(bsz * q_len, -1)
                qk = mixedgemm.pair_adjacent_rows(y[:, :nrope].t(), self.num_heads + self.num_key_value_heads, self.head_dim,
                                                  inverse=True).t()
                y = torch.cat([qk, y[:, nrope:]], dim=-1)
                position_embeddings_applied = False
            q, k, v = y.view(bsz, q_len, -1).split(m.splits, dim=-1)
        elif tables is not None:
            m = self.qkv_proj[0]
            if self.sp:
                y = m.forward_gathered_whole(hidden_states.reshape(-1, hidden_states.shape[-1]).contiguous(), bsz * q_len,
                                             self.workspace, norm)
            else:
                y = m.forward_whole(hidden_states, norm)
            y = y.reshape(bsz * q_len, -1)
            mixedgemm.rope_inplace(y, self.num_heads + self.num_key_value_heads, self.head_dim, tables[0], tables[1])
            q, k, v = y.view(bsz, q_len, -1).split(m.splits, dim=-1)
        elif self.sp:
            q, k, v = self.qkv_proj[0].forward_gathered(hidden_states.reshape(-1, hidden_states.shape[-1]).contiguous(),
                                                        bsz * q_len, self.workspace, norm)
        else:
            q, k, v = run_input_group(self.qkv_proj, hidden_states, norm)
        q = q.view(bsz, q_len, self.num_heads, self.head_dim).transpose(1, 2)
        k = k.view(bsz, q_len, self.num_key_value_heads, self.head_dim).transpose(1, 2)
        v = v.view(bsz, q_len, self.num_key_value_heads, self.head_dim).transpose(1, 2)
        if position_embeddings is not None:
            cos, sin = position_embeddings
            if self.rope_epilogue:
                if not position_embeddings_applied:
                    q, k = apply_rope(q, k, cos, sin)
            elif tables is None:
                q, k = apply_rope(q, k, cos, sin)
        else:
            cos = sin = None
        if past_key_value is not None:
            k, v = past_key_value.update(k, v, self.layer_idx, {"sin": sin, "cos": cos, "cache_position": cache_position})
        if self.q_kv_cache:
            v = quantize_int_group(v, nbits=4, group_size=128)
            k = quantize_int_group(k, nbits=4, group_size=128)
        mask = attention_mask
        if mask is not None:
            mask = mask[:, :, :, : k.shape[-2]]
        out = F.scaled_dot_product_attention(q, k, v, attn_mask=mask,
                                             dropout_p=self.attention_dropout if self.training else 0.0,
                                             is_causal=mask is None and q_len > 1,
                                             enable_gqa=self.num_key_value_groups > 1)
        out = out.transpose(1, 2).reshape(bsz, q_len, -1)
        if self.sp:
            # this rank's rows only: reduce-scatter of the K-sharded product, or the token-parallel full-K product
            y, _ = self.o_proj(out) if self.tpr else forward_row_shard(self.o_proj, out)
            return y.unsqueeze(0), None, past_key_value
        if res is not None:
            return self.o_proj(out, residual=res), None, past_key_value
        return self.o_proj(out), None, past_key_value


class QGatedMLP(nn.Module):
    """down(act(gate(x)) * up(x)) with quantized projections (qLlamaLayer.py:324-387, qQwenLayer.py:330-393)."""

    def __init__(self, originalMLP, p8_nums, p6_nums, reorder_index, i, tp_group=None, names=('gate_proj', 'up_proj',
                                                                                               'down_proj'),
                 key_fmt=None, fused_act=False, workspace=None, sequence_parallel=False, token_parallel_rows=False,
                 act_epilogue=True):
        super().__init__()
        self.act_epilogue = False
        self.sp = bool(sequence_parallel)
        self.tpr = bool(token_parallel_rows) and self.sp
        self.workspace = workspace
        self.tp, self.rank = tp_info(tp_group)
        gate, up, down = (getattr(originalMLP, n) for n in names)
        key = key_fmt or (lambda n: NAME.format(i, 'mlp', n, 'input'))
        inter = gate.out_features
        self.act_fn = getattr(originalMLP, "act_fn", F.silu)
        is_silu = self.act_fn is F.silu or isinstance(self.act_fn, nn.SiLU) or type(self.act_fn).__name__ in ("SiLU", "SiLUActivation")
        self.fused_act = bool(fused_act) and self.tp == 1 and is_silu
        if self.fused_act:
            # gate / up rows in down_proj's channel order: the intermediate activation is born permuted and
            # SiLU(gate) * up is quantized in place by activate_quantize_x; down_proj's columns follow the same order
            kd = key(names[2])
            perm = reorder_index[kd].to(torch.int64)
            pw = lambda l: _meta_linear(l.weight.data[perm.to(l.weight.device)],
                                        None if l.bias is None else l.bias.data[perm.to(l.bias.device)])
            self.d_p8, self.d_p6 = int(p8_nums[kd]), int(p6_nums[kd])
            self.d_p4 = inter - self.d_p8 - self.d_p6
            kg, ku = key(names[0]), key(names[1])
            shared = (_same(reorder_index[kg], reorder_index[ku]) and _same(p8_nums[kg], p8_nums[ku])
                      and _same(p6_nums[kg], p6_nums[ku]))
            # SiLU * up + quantize in the gate_up GEMM's epilogue: one calibration for both projections, no biases
            self.act_epilogue = bool(act_epilogue and shared and gate.bias is None and up.bias is None and inter % 128 == 0
                                     and self.d_p4 > 0)
            if self.act_epilogue:
                w = mixedgemm.interleave_gate_up(gate.weight.data[perm.to(gate.weight.device)],
                                                 up.weight.data[perm.to(up.weight.device)])
                self.gate_up_proj = nn.ModuleList([FusedQLinear([_meta_linear(w, None)], p8_nums[kg], p6_nums[kg],
                                                                reorder_index[kg])])
                del w
            else:
                self.gate_up_proj = build_input_group([pw(gate), pw(up)], [kg, ku], p8_nums, p6_nums, reorder_index, None)
            wd = down.weight.data.to(device='cuda', dtype=torch.bfloat16)[:, perm.cuda()].contiguous()
            self.dW = mixedgemm.downproj_quantize_w4(wd, self.d_p4, self.d_p6, self.d_p8)
            del wd
            self.d_bias = None if down.bias is None else down.bias.detach().to(device='cuda', dtype=torch.bfloat16).contiguous()
            self.down_proj = None
            return
        slices = None
        if self.tp > 1:
            per = inter // self.tp
            slices = [slice(self.rank * per, (self.rank + 1) * per)] * 2
        self.gate_up_proj = build_input_group([gate, up], [key(names[0]), key(names[1])], p8_nums, p6_nums,
                                              reorder_index, slices)
        kd = key(names[2])
        if self.tp > 1 and self.tpr:
            self.down_proj = TokenParallelQLinear(down, p8_nums[kd], p6_nums[kd], reorder_index[kd], tp_group, workspace)
        elif self.tp > 1:
            self.down_proj = RowParallelQLinear(down, p8_nums[kd], p6_nums[kd], reorder_index[kd], tp_group,
                                                workspace=workspace)
        else:
            self.down_proj = QLinearLayer(down, p8_nums[kd], p6_nums[kd], reorder_index[kd])

    @torch.no_grad()
    def forward(self, x, norm=None, tokens=None, residual=None):
        """`residual` (single-GPU forms): returns residual + down(...), added in down_proj's GEMM epilogue."""
        if self.sp:
            # x = this rank's token rows [1, rows, hidden] of `tokens` in total
            g, u = self.gate_up_proj[0].forward_gathered(x.reshape(-1, x.shape[-1]).contiguous(), int(tokens), self.workspace,
                                                         norm)
            h = (self.act_fn(g) * u).unsqueeze(0)
            y, _ = self.down_proj(h) if self.tpr else forward_row_shard(self.down_proj, h)
            return y.unsqueeze(0)
        W = self.dW if self.fused_act else None
        if self.act_epilogue:
            bsz, q_len, _ = x.shape
            a = self.gate_up_proj[0].forward_activated(x, (self.d_p4, self.d_p6, self.d_p8), norm)
            r2 = None if residual is None else residual.reshape(bsz * q_len, -1).contiguous()
            y = mixedgemm.matmul(a[0], W[0], a[1], W[1], a[2], W[2], a[3], W[3], a[4], W[4], a[5], W[5], bias=self.d_bias,
                                 residual=r2)
            return y.reshape(bsz, q_len, -1)
        g, u = run_input_group(self.gate_up_proj, x, norm)
        if not self.fused_act:
            if residual is not None:
                return self.down_proj(self.act_fn(g) * u, residual=residual)
            return self.down_proj(self.act_fn(g) * u)
        bsz, q_len, inter = g.shape
        a = mixedgemm.activate_quantize_x(g.reshape(bsz * q_len, inter), u.reshape(bsz * q_len, inter), self.d_p4, self.d_p6,
                                          self.d_p8)
        r2 = None if residual is None else residual.reshape(bsz * q_len, -1).contiguous()
        y = mixedgemm.matmul(a[0], W[0], a[1], W[1], a[2], W[2], a[3], W[3], a[4], W[4], a[5], W[5], bias=self.d_bias,
                             residual=r2)
        return y.reshape(bsz, q_len, -1)


class QDecoderLayer(nn.Module):
    """norm -> attention -> residual -> norm -> MLP -> residual, the reference's forward contract
    (qLlamaLayer.py:116-158): returns (hidden_states,) [+ (attn_weights,)] [+ (present_key_value,)]."""

    def __init__(self, originalLayer, kv_cache, p8_nums, p6_nums, reorder_index, layer_idx, tp_group=None, fused=False,
                 workspace=None, sequence_parallel=False, token_parallel_rows=False, rope_epilogue=False):
        """`rope_epilogue` (with fused, single GPU; OFF by default): RoPE in the qkv GEMM's epilogue instead of the in-place
        kernel.  Bit-identical, but measured SLOWER (Llama-3-8B prefill 79.9 vs 75.2 ms, profiles/r02_rope_epilogue.txt): a
        TMEM lane is a token row, so the cos / sin loads and the un-permuting stores of a warp touch 32 different lines per
        instruction and the epilogue outlasts the MMAs of a K = 4096 tile.
        `workspace` (extension): a parallel_utils.PeerWorkspace shared by the model's row-parallel linears -- o_proj and
        down_proj then run as GEMMs fused with their all-reduce instead of GEMM + NCCL all-reduce.
        `sequence_parallel` (extension, needs a workspace with a gather channel): the layer takes and returns THIS rank's
        token rows only ([1, rows, hidden], rows = workspace.shard_range(b*s)): o_proj / down_proj end in a reduce-scatter,
        residual + RMSNorm + quantize run on the shard, and the packed MX codes (not bf16 activations) are all-gathered
        into the column-parallel GEMMs through NVSwitch multicast.
        `token_parallel_rows` (with sequence_parallel): o_proj / down_proj keep a REPLICATED MXFP4 weight and exchange the
        packed codes of their input all-to-all instead of reducing bf16 partial sums (parallel_utils.TokenParallelQLinear)."""
        super().__init__()
        self.fused = bool(fused)
        self.sp = bool(sequence_parallel)
        self.tpr = bool(token_parallel_rows) and self.sp
        if self.sp and (workspace is None or tp_info(tp_group)[0] < 2 or not workspace.gather[0]):
            raise ValueError("sequence_parallel needs tp >= 2 and a PeerWorkspace(..., gather=(tokens, hidden))")
        self._workspace = workspace
        self.hidden_size = getattr(originalLayer, "hidden_size", None) or originalLayer.self_attn.config.hidden_size
        self.self_attn = QAttention(originalLayer.self_attn, kv_cache, p8_nums, p6_nums, reorder_index, layer_idx,
                                    tp_group, workspace, sequence_parallel=self.sp, token_parallel_rows=self.tpr,
                                    rope_epilogue=self.fused and rope_epilogue)
        if self.sp and len(self.self_attn.qkv_proj) != 1:
            raise ValueError("sequence_parallel needs q/k/v to share one (reorder_index, p6, p8): one gather feeds ONE GEMM")
        self.mlp = self._build_mlp(originalLayer, p8_nums, p6_nums, reorder_index, layer_idx, tp_group)
        self.input_layernorm = originalLayer.input_layernorm
        self.post_attention_layernorm = originalLayer.post_attention_layernorm

    def _build_mlp(self, originalLayer, p8_nums, p6_nums, reorder_index, layer_idx, tp_group):
        return QGatedMLP(originalLayer.mlp, p8_nums, p6_nums, reorder_index, layer_idx, tp_group, fused_act=self.fused,
                         workspace=self._workspace, sequence_parallel=self.sp, token_parallel_rows=self.tpr)

    @torch.no_grad()
    def forward(self, hidden_states, attention_mask=None, position_ids=None, past_key_value=None,
                output_attentions=False, use_cache=False, cache_position=None, position_embeddings=None, **kwargs):
        residual = hidden_states
        if self.sp:
            # the RMSNorms always run inside the quantizer of the shard when they can: that IS the sequence-parallel hand-over
            norm1 = fusable_rmsnorm(self.input_layernorm)
            if norm1 is None:
                hidden_states = self.input_layernorm(hidden_states)
            else:
                kwargs["mmx_norm"] = norm1
            hidden_states, attn_weights, present = self.self_attn(
                hidden_states=hidden_states, attention_mask=attention_mask, position_ids=position_ids,
                past_key_value=past_key_value, output_attentions=output_attentions, use_cache=use_cache,
                cache_position=cache_position, position_embeddings=position_embeddings, **kwargs)
            hidden_states = residual + hidden_states
            residual = hidden_states
            tokens = position_embeddings[0].shape[0] * position_embeddings[0].shape[1]
            norm2 = fusable_rmsnorm(self.post_attention_layernorm)
            if norm2 is None:
                hidden_states = self.mlp(self.post_attention_layernorm(hidden_states), None, tokens)
            else:
                hidden_states = self.mlp(hidden_states, norm2, tokens)
            hidden_states = residual + hidden_states
            outputs = (hidden_states,)
            if output_attentions:
                outputs += (attn_weights,)
            if use_cache:
                outputs += (present,)
            return outputs
        norm1 = fusable_rmsnorm(self.input_layernorm) if self.fused else None
        if norm1 is None:
            hidden_states = self.input_layernorm(hidden_states)
        else:
            kwargs["mmx_norm"] = norm1  # RMSNorm runs inside the qkv quantizer
        # fused, single GPU: the two residual adds run in the epilogues of o_proj / down_proj (same rounding as the torch add)
        res_o = self.fused and isinstance(self.self_attn.o_proj, QLinearLayer)
        if res_o:
            kwargs["mmx_residual"] = residual
        hidden_states, attn_weights, present = self.self_attn(
            hidden_states=hidden_states, attention_mask=attention_mask, position_ids=position_ids,
            past_key_value=past_key_value, output_attentions=output_attentions, use_cache=use_cache,
            cache_position=cache_position, position_embeddings=position_embeddings, **kwargs)
        if not res_o:
            hidden_states = residual + hidden_states
        residual = hidden_states
        gated = isinstance(self.mlp, QGatedMLP)
        norm2 = fusable_rmsnorm(self.post_attention_layernorm) if (self.fused and gated) else None
        res_d = self.fused and gated and self.mlp.tp == 1
        if norm2 is None:
            hidden_states = self.post_attention_layernorm(hidden_states)
        if res_d:
            hidden_states = self.mlp(hidden_states, norm2, residual=residual)
        elif norm2 is None:
            hidden_states = self.mlp(hidden_states)
        else:
            hidden_states = self.mlp(hidden_states, norm2)
        if isinstance(hidden_states, tuple):  # MoE blocks return (hidden_states, router_logits)
            hidden_states = hidden_states[0]
        if not res_d:
            hidden_states = residual + hidden_states
        outputs = (hidden_states,)
        if output_attentions:
            outputs += (attn_weights,)
        if use_cache:
            outputs += (present,)
        return outputs
