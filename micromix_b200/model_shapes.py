"""Random-init decoder layers of the shapes the BASELINE configs name (no checkpoints, no network): plain nn.Modules with
the attribute names of the HF layers the reference wraps, plus synthetic calibration dicts (reorder_index / p6 / p8)."""
from __future__ import annotations

from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F

LLAMA3_8B = dict(hidden_size=4096, intermediate_size=14336, num_attention_heads=32, num_key_value_heads=8, head_dim=128,
                 num_hidden_layers=32, rms_norm_eps=1e-5, qkv_bias=False, rope_theta=500000.0)
QWEN25_32B = dict(hidden_size=5120, intermediate_size=27648, num_attention_heads=40, num_key_value_heads=8, head_dim=128,
                  num_hidden_layers=64, rms_norm_eps=1e-6, qkv_bias=True, rope_theta=1000000.0)
MIXTRAL_8X7B = dict(hidden_size=4096, intermediate_size=14336, num_attention_heads=32, num_key_value_heads=8,
                    head_dim=128, num_hidden_layers=32, rms_norm_eps=1e-5, qkv_bias=False, rope_theta=1000000.0,
                    num_local_experts=8, num_experts_per_tok=2)


class RMSNorm(nn.Module):
    def __init__(self, dim, eps, device=None):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim, dtype=torch.bfloat16, device=device), requires_grad=False)
        self.variance_epsilon = eps

    def forward(self, x):
        v = x.float()
        v = v * torch.rsqrt(v.pow(2).mean(-1, keepdim=True) + self.variance_epsilon)
        return self.weight * v.to(x.dtype)


def _linear(i, o, bias, device, gen, std=0.02):
    lin = nn.Linear(i, o, bias=bias, device="meta", dtype=torch.bfloat16)
    w = (torch.randn(o, i, generator=gen, device=device, dtype=torch.float32) * std).to(torch.bfloat16)
    lin.weight = nn.Parameter(w, requires_grad=False)
    if bias:
        lin.bias = nn.Parameter((torch.randn(o, generator=gen, device=device) * 0.1).to(torch.bfloat16),
                                requires_grad=False)
    return lin


def make_layer(cfg: dict, device, seed=0, moe=False):
    """A random-init decoder layer with HF attribute names (self_attn.{q,k,v,o}_proj, mlp.{gate,up,down}_proj or
    block_sparse_moe.{gate, experts[j].{w1,w2,w3}}, input_layernorm, post_attention_layernorm)."""
    g = torch.Generator(device=device).manual_seed(seed)
    c = SimpleNamespace(**cfg)
    h, d = c.hidden_size, c.head_dim
    attn = nn.Module()
    attn.config = c
    attn.attention_dropout = 0.0
    attn.q_proj = _linear(h, c.num_attention_heads * d, c.qkv_bias, device, g)
    attn.k_proj = _linear(h, c.num_key_value_heads * d, c.qkv_bias, device, g)
    attn.v_proj = _linear(h, c.num_key_value_heads * d, c.qkv_bias, device, g)
    attn.o_proj = _linear(c.num_attention_heads * d, h, False, device, g)
    layer = nn.Module()
    layer.hidden_size = h
    layer.self_attn = attn
    if moe:
        blk = nn.Module()
        blk.num_experts, blk.top_k = c.num_local_experts, c.num_experts_per_tok
        blk.gate = _linear(h, c.num_local_experts, False, device, g)
        blk.experts = nn.ModuleList()
        for _ in range(c.num_local_experts):
            e = nn.Module()
            e.w1, e.w3 = _linear(h, c.intermediate_size, False, device, g), _linear(h, c.intermediate_size, False, device, g)
            e.w2 = _linear(c.intermediate_size, h, False, device, g)
            e.act_fn = F.silu
            blk.experts.append(e)
        layer.block_sparse_moe = blk
        layer.mlp = blk
    else:
        mlp = nn.Module()
        mlp.gate_proj = _linear(h, c.intermediate_size, False, device, g)
        mlp.up_proj = _linear(h, c.intermediate_size, False, device, g)
        mlp.down_proj = _linear(c.intermediate_size, h, False, device, g)
        mlp.act_fn = F.silu
        layer.mlp = mlp
    layer.input_layernorm = RMSNorm(h, c.rms_norm_eps, device)
    layer.post_attention_layernorm = RMSNorm(h, c.rms_norm_eps, device)
    return layer


def split_for(K):
    """5.0 average bits: p8 = K/8, p6 = K/4, rounded down to multiples of 128."""
    p8 = (K // 8) // 128 * 128
    p6 = (K // 4) // 128 * 128
    return K - p6 - p8, p6, p8


def make_calibration(cfg: dict, layer_idx: int, seed=0, moe=False):
    """Synthetic (reorder_index, p6_nums, p8_nums) dicts with the reference's keys ('layers.{i}.self_attn.q_proj.input',
    ...; reorder_indices.py:98-121): one permutation per distinct input tensor."""
    g = torch.Generator().manual_seed(1000 * seed + layer_idx)
    h, inter = cfg["hidden_size"], cfg["intermediate_size"]
    idx, p6, p8 = {}, {}, {}

    def put(keys, K):
        perm = torch.randperm(K, generator=g).to(torch.int16)
        _, s6, s8 = split_for(K)
        for k in keys:
            idx[k], p6[k], p8[k] = perm, s6, s8

    t = 'layers.{}.{}.{}.input'
    put([t.format(layer_idx, 'self_attn', n) for n in ('q_proj', 'k_proj', 'v_proj')], h)
    put([t.format(layer_idx, 'self_attn', 'o_proj')], cfg["num_attention_heads"] * cfg["head_dim"])
    if moe:
        te = 'layers.{}.block_sparse_moe.experts.{}.{}.input'
        for j in range(cfg["num_local_experts"]):
            put([te.format(layer_idx, j, 'w1'), te.format(layer_idx, j, 'w3')], h)
            put([te.format(layer_idx, j, 'w2')], inter)
    else:
        put([t.format(layer_idx, 'mlp', 'gate_proj'), t.format(layer_idx, 'mlp', 'up_proj')], h)
        put([t.format(layer_idx, 'mlp', 'down_proj')], inter)
    return idx, p6, p8


def rope_tables(cfg: dict, bsz, seq, device):
    d = cfg["head_dim"]
    inv = 1.0 / (cfg["rope_theta"] ** (torch.arange(0, d, 2, device=device, dtype=torch.float32) / d))
    t = torch.arange(seq, device=device, dtype=torch.float32)
    f = torch.outer(t, inv)
    emb = torch.cat((f, f), dim=-1)
    cos, sin = emb.cos().to(torch.bfloat16), emb.sin().to(torch.bfloat16)
    return cos[None].expand(bsz, -1, -1), sin[None].expand(bsz, -1, -1)
