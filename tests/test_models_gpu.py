"""GPU tests of the decoder-layer wrappers (qLlamaLayer / qQwenLayer / qMixtralLayer).

The check is against the REFERENCE'S structure rebuilt from per-projection QLinearLayers (one quantize + one GEMM per
linear, python loop over experts: qLlamaLayer.py:252-321,368-387, qMixtralLayer.py:420-450): fusing q/k/v, gate/up and
w1/w3 must not change a single bit, because every output column accumulates independently of the N tiling.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

TINY = dict(hidden_size=512, intermediate_size=1024, num_attention_heads=4, num_key_value_heads=2, head_dim=128,
            num_hidden_layers=2, rms_norm_eps=1e-5, qkv_bias=False, rope_theta=10000.0)
TINY_MOE = dict(TINY, num_local_experts=4, num_experts_per_tok=2)


def _unfused_layer(layer, cfg, idx, p6, p8, i, x, cos, sin, moe=False):
    from micromix_b200._qdecoder import apply_rope
    from micromix_b200.qLinearLayer import QLinearLayer

    def q(lin, key):
        return QLinearLayer(lin, p8[key], p6[key], idx[key])

    t = 'layers.{}.{}.{}.input'
    b, s, _ = x.shape
    a = layer.self_attn
    h = layer.input_layernorm(x)
    qs = q(a.q_proj, t.format(i, 'self_attn', 'q_proj'))(h).view(b, s, cfg['num_attention_heads'], -1).transpose(1, 2)
    ks = q(a.k_proj, t.format(i, 'self_attn', 'k_proj'))(h).view(b, s, cfg['num_key_value_heads'], -1).transpose(1, 2)
    vs = q(a.v_proj, t.format(i, 'self_attn', 'v_proj'))(h).view(b, s, cfg['num_key_value_heads'], -1).transpose(1, 2)
    qs, ks = apply_rope(qs, ks, cos, sin)
    o = F.scaled_dot_product_attention(qs, ks, vs, is_causal=True, enable_gqa=True)
    o = q(a.o_proj, t.format(i, 'self_attn', 'o_proj'))(o.transpose(1, 2).reshape(b, s, -1))
    x = x + o
    h = layer.post_attention_layernorm(x)
    if not moe:
        m = layer.mlp
        g = q(m.gate_proj, t.format(i, 'mlp', 'gate_proj'))(h)
        u = q(m.up_proj, t.format(i, 'mlp', 'up_proj'))(h)
        y = q(m.down_proj, t.format(i, 'mlp', 'down_proj'))(F.silu(g) * u)
    else:
        blk = layer.block_sparse_moe
        te = 'layers.{}.block_sparse_moe.experts.{}.{}.input'
        hx = h.view(-1, h.shape[-1])
        w = F.softmax(blk.gate(hx), dim=1, dtype=torch.float)
        w, sel = torch.topk(w, blk.top_k, dim=-1)
        w = (w / w.sum(dim=-1, keepdim=True)).to(hx.dtype)
        y = torch.zeros_like(hx)
        for j, e in enumerate(blk.experts):
            tok, slot = torch.where(sel == j)
            if tok.numel() == 0:
                continue
            cur = hx.index_select(0, tok).unsqueeze(0)
            g = q(e.w1, te.format(i, j, 'w1'))(cur)
            u = q(e.w3, te.format(i, j, 'w3'))(cur)
            yy = q(e.w2, te.format(i, j, 'w2'))(F.silu(g) * u).squeeze(0) * w[tok, slot, None]
            y.index_add_(0, tok, yy.to(hx.dtype))
        y = y.view_as(h)
    return x + y


def _run(cuda, cfg, cls_path, moe=False, bias=False, bsz=2, seq=96):
    import importlib
    from micromix_b200 import model_shapes as S
    cfg = dict(cfg, qkv_bias=bias)
    layer = S.make_layer(cfg, cuda, seed=3, moe=moe)
    idx, p6, p8 = S.make_calibration(cfg, 1, seed=5, moe=moe)
    mod, name = cls_path.rsplit('.', 1)
    cls = getattr(importlib.import_module(mod), name)
    qlayer = cls(layer, False, p8, p6, idx, 1)
    g = torch.Generator(device=cuda).manual_seed(11)
    x = torch.randn(bsz, seq, cfg['hidden_size'], generator=g, device=cuda, dtype=torch.float32).to(torch.bfloat16)
    cos, sin = S.rope_tables(cfg, bsz, seq, cuda)
    ref = _unfused_layer(layer, cfg, idx, p6, p8, 1, x, cos, sin, moe=moe)
    # (1) with RoPE left to the torch ops the layer must equal the reference structure bit for bit
    import micromix_b200._qdecoder as QD
    keep = QD.rope_tables_2d
    QD.rope_tables_2d = lambda *a, **k: None
    try:
        out = qlayer(x, position_embeddings=(cos, sin))
    finally:
        QD.rope_tables_2d = keep
    assert isinstance(out, tuple) and len(out) == 1
    torch.cuda.synchronize()
    assert out[0].shape == x.shape and out[0].dtype == torch.bfloat16
    assert torch.isfinite(out[0].float()).all()
    assert torch.equal(out[0], ref), float((out[0].float() - ref.float()).abs().max())
    # (2) the default path rotates q and k IN PLACE inside the fused qkv output (bit-identical to the torch ops:
    # test_rope_inplace_matches_hf_ops_bit_for_bit) and hands SDPA strided views of it; the attention library may then pick
    # another kernel / summation order, so the layer output is compared within the noise that leaves behind
    out2 = qlayer(x, position_embeddings=(cos, sin))[0]
    torch.cuda.synchronize()
    d = (out2.float() - ref.float()).abs()
    scale = ref.float().pow(2).mean().sqrt()
    assert float(d.mean() / scale) <= 5e-3 and float(d.max() / scale) <= 0.25, (float(d.mean() / scale), float(d.max() / scale))
    return qlayer, x, cos, sin


def test_llama_layer_matches_unfused_reference_structure(cuda):
    q, x, cos, sin = _run(cuda, TINY, 'micromix_b200.qLlamaLayer.QLlamaDecoderLayer')
    assert len(q.self_attn.qkv_proj) == 1 and len(q.mlp.gate_up_proj) == 1  # fused: one quantize + one GEMM each
    out = q(x, position_embeddings=(cos, sin), output_attentions=True, use_cache=True)
    assert len(out) == 3  # (hidden, attn_weights=None, present=None): the reference's tuple contract


def test_qwen_layer_with_qkv_bias(cuda):
    _run(cuda, TINY, 'micromix_b200.qQwenLayer.QQwen2DecoderLayer', bias=True)


def test_mixtral_layer_matches_expert_loop(cuda):
    _run(cuda, TINY_MOE, 'micromix_b200.qMixtralLayer.QMixtralDecoderLayer', moe=True)


def test_unshared_calibration_falls_back_to_separate_linears(cuda):
    from micromix_b200 import model_shapes as S
    from micromix_b200.qLlamaLayer import QLlamaDecoderLayer
    layer = S.make_layer(TINY, cuda, seed=3)
    idx, p6, p8 = S.make_calibration(TINY, 0, seed=5)
    k = 'layers.0.self_attn.k_proj.input'
    idx[k] = torch.randperm(TINY['hidden_size'], generator=torch.Generator().manual_seed(9)).to(torch.int16)
    q = QLlamaDecoderLayer(layer, False, p8, p6, idx, 0)
    assert len(q.self_attn.qkv_proj) == 3
    x = torch.randn(1, 40, TINY['hidden_size'], device=cuda).to(torch.bfloat16)
    cos, sin = S.rope_tables(TINY, 1, 40, cuda)
    ref = _unfused_layer(layer, TINY, idx, p6, p8, 0, x, cos, sin)
    assert torch.equal(q(x, position_embeddings=(cos, sin))[0], ref)


@pytest.mark.parametrize("cls_path,bias", [('micromix_b200.qLlamaLayer.QLlamaDecoderLayer', False),
                                           ('micromix_b200.qQwenLayer.QQwen2DecoderLayer', True)])
def test_fused_norm_and_activation_mode(cuda, cls_path, bias):
    """fused=True: RMSNorm inside the quantizer (rmsnorm_quantize_x) and SiLU(gate) * up inside down_proj's quantizer
    (activate_quantize_x on the halves of the gate_up output, gate / up rows stored in down_proj's channel order).
    Same function as the unfused layer up to where the roundings sit (norm rounded once instead of twice, fp32 product):
    compared against a float64 reference of the layer built from the DEQUANTISED op sequence is overkill here -- the two
    modes must agree within a few bf16 steps of the layer output's scale."""
    import importlib
    from micromix_b200 import mixedgemm
    from micromix_b200 import model_shapes as S
    cfg = dict(TINY, qkv_bias=bias)
    layer = S.make_layer(cfg, cuda, seed=3)
    idx, p6, p8 = S.make_calibration(cfg, 1, seed=5)
    mod, name = cls_path.rsplit('.', 1)
    cls = getattr(importlib.import_module(mod), name)
    plain = cls(layer, False, p8, p6, idx, 1)
    fused = cls(layer, False, p8, p6, idx, 1, fused=True)
    assert fused.mlp.fused_act and fused.mlp.down_proj is None
    g = torch.Generator(device=cuda).manual_seed(11)
    x = torch.randn(2, 96, cfg['hidden_size'], generator=g, device=cuda, dtype=torch.float32).to(torch.bfloat16)
    cos, sin = S.rope_tables(cfg, 2, 96, cuda)
    n0 = mixedgemm.launch_count()
    a = plain(x, position_embeddings=(cos, sin))[0]
    n1 = mixedgemm.launch_count()
    b = fused(x, position_embeddings=(cos, sin))[0]
    n2 = mixedgemm.launch_count()
    torch.cuda.synchronize()
    # plain: four quantize + four GEMM launches + the in-place RoPE kernel; fused: down_proj's quantizer is the gate_up
    # GEMM's epilogue, one launch less; the elementwise kernels are gone
    assert n1 - n0 == 9 and n2 - n1 == 8
    assert torch.isfinite(b.float()).all()
    # the epilogue form is bit-identical to gate_up GEMM -> activate_quantize_x (the reference's op pair)
    from micromix_b200._qdecoder import QGatedMLP, fusable_rmsnorm
    two_ops = QGatedMLP(layer.mlp, p8, p6, idx, 1, fused_act=True, act_epilogue=False)
    assert fused.mlp.act_epilogue and two_ops.fused_act and not two_ops.act_epilogue
    norm = fusable_rmsnorm(layer.post_attention_layernorm)
    assert norm is not None
    assert torch.equal(fused.mlp(x, norm), two_ops(x, norm))
    assert torch.equal(fused.mlp(x), two_ops(x))
    # ... and RoPE in the qkv GEMM's epilogue to GEMM -> rope_inplace; without tables the pair-adjacent columns are undone
    from micromix_b200._qdecoder import QAttention
    att = QAttention(layer.self_attn, False, p8, p6, idx, 1)
    att_e = QAttention(layer.self_attn, False, p8, p6, idx, 1, rope_epilogue=True)  # (opt-in: measured slower)
    assert att_e.rope_epilogue and not att.rope_epilogue and not fused.self_attn.rope_epilogue
    assert torch.equal(att_e(hidden_states=x, position_embeddings=(cos, sin))[0],
                       att(hidden_states=x, position_embeddings=(cos, sin))[0])
    assert torch.equal(att_e(hidden_states=x, position_embeddings=None)[0], att(hidden_states=x, position_embeddings=None)[0])
    d = (a.float() - b.float()).abs()
    scale = a.float().pow(2).mean().sqrt()
    assert float(d.max()) <= 0.05 * float(scale) and float(d.mean()) <= 0.005 * float(scale), (float(d.max()), float(d.mean()), float(scale))


@pytest.mark.parametrize("b,s,nh,nkv,d,extra", [(2, 96, 4, 2, 128, 256), (1, 33, 32, 8, 128, 1024), (3, 17, 2, 2, 64, 0)])
def test_rope_inplace_matches_hf_ops_bit_for_bit(cuda, b, s, nh, nkv, d, extra):
    """mixedgemm.rope_inplace on the q | k columns of a fused qkv output == HF's q * cos + rotate_half(q) * sin on the
    transposed views (the torch ops of qLlamaLayer.py:25-54), every bf16 rounding included; the v columns are untouched."""
    from micromix_b200 import mixedgemm
    from micromix_b200 import model_shapes as S
    from micromix_b200._qdecoder import apply_rope, rope_tables_2d
    g = torch.Generator(device=cuda).manual_seed(b * 100 + s)
    ld = (nh + nkv) * d + extra
    y = (torch.randn(b * s, ld, generator=g, device=cuda) * 3).to(torch.bfloat16)
    y0 = y.clone()
    cos, sin = S.rope_tables(dict(head_dim=d, rope_theta=10000.0), b, s, cuda)
    q = y0[:, :nh * d].view(b, s, nh, d).transpose(1, 2)
    k = y0[:, nh * d:(nh + nkv) * d].view(b, s, nkv, d).transpose(1, 2)
    qr, kr = apply_rope(q, k, cos, sin)
    tabs = rope_tables_2d((cos, sin), b, s, d)
    assert tabs is not None and tabs[0].shape == (s, d)
    mixedgemm.rope_inplace(y, nh + nkv, d, *tabs)
    torch.cuda.synchronize()
    assert torch.equal(y[:, :nh * d].view(b, s, nh, d).transpose(1, 2), qr)
    assert torch.equal(y[:, nh * d:(nh + nkv) * d].view(b, s, nkv, d).transpose(1, 2), kr)
    assert torch.equal(y[:, (nh + nkv) * d:], y0[:, (nh + nkv) * d:])
    # per-token tables (contiguous [b, s, d]) take the S = M form
    cos2, sin2 = cos.contiguous(), sin.contiguous()
    tabs2 = rope_tables_2d((cos2, sin2), b, s, d)
    y2 = y0.clone()
    mixedgemm.rope_inplace(y2, nh + nkv, d, *tabs2)
    assert torch.equal(y2, y)
