"""GPU tests of the layer-level drop-in: QLinearLayer (model/qLinearLayer.py:20-74) against the CPU fake-quant path."""
import numpy as np
import pytest
import torch
import torch.nn as nn

import helpers as H

O = H.O
pytestmark = pytest.mark.gpu


def _ref_linear(x, w, idx, split, bias=None):
    y = O.fake_quant_linear(H.bits(x), H.bits(w), idx.numpy(), *split, chain=False)
    if bias is not None:
        y = O.f32_to_bf16_bits(O.bf16_bits_to_f32(y) + O.bf16_bits_to_f32(H.bits(bias))[None, :])
    return y


@pytest.mark.parametrize("use_bias", [False, True])
def test_qlinear_forward_matches_fake_quant(cuda, use_bias):
    from micromix_b200.qLinearLayer import QLinearLayer
    K, N, b, s = 1024, 768, 2, 75
    split = (512, 256, 256)
    idx = H.make_index(K, seed=21)
    lin = nn.Linear(K, N, bias=use_bias).to(torch.bfloat16)
    lin.weight.data = H.make_weights(N, K)
    if use_bias:
        lin.bias.data = (torch.randn(N) * 0.1).to(torch.bfloat16)
    q = QLinearLayer(lin, p8_num=split[2], p6_num=split[1], reorder_index=idx)
    assert (q.p4_num, q.p6_num, q.p8_num) == split
    x = H.make_activations(b * s, K, idx).reshape(b, s, K)
    y = q(x.to(cuda))
    assert y.shape == (b, s, N) and y.dtype == torch.bfloat16
    ref = _ref_linear(x.reshape(b * s, K), lin.weight.data, idx, split, lin.bias.data if use_bias else None)
    mx, mean = H.rel_err(H.bits(y.reshape(b * s, N)), ref)
    assert mx <= 1e-2 and mean <= 1e-3, (mx, mean)


def test_qlinear_weight_buffers_match_reference_layout(cuda):
    from micromix_b200.qLinearLayer import QLinearLayer
    K, N = 512, 256
    idx = H.make_index(K, seed=22)
    lin = nn.Linear(K, N, bias=False).to(torch.bfloat16)
    lin.weight.data = H.make_weights(N, K)
    q = QLinearLayer(lin, 128, 128, idx)
    ref = O.reorder_quantize(H.bits(lin.weight.data), idx.numpy(), 256, 128, 128, "w4")
    for got, r in zip((q.BN, q.BS, q.BO, q.SFBN, q.SFBS, q.SFBO), ref):
        assert np.array_equal(H.u8(got), r)
    assert q.reorder_index.dtype == torch.int16 and q.reorder_index.is_cuda


def test_qlinear_rejects_bad_split(cuda):
    from micromix_b200.qLinearLayer import QLinearLayer
    lin = nn.Linear(512, 256, bias=False).to(torch.bfloat16)
    with pytest.raises(ValueError):
        QLinearLayer(lin, 100, 128, H.make_index(512))


@pytest.mark.parametrize("use_bias", [False, True])
def test_qlinear_packed_checkpoint_round_trip(cuda, tmp_path, use_bias):
    """the six packed-weight tensors are buffers (state_dict / .to()) and save_packed() /
    from_packed() rebuild the layer without re-quantizing -- same bytes, same forward bits."""
    from micromix_b200.qLinearLayer import PACKED_NAMES, QLinearLayer
    K, N = 1024, 512
    idx = H.make_index(K, seed=23)
    lin = nn.Linear(K, N, bias=use_bias).to(torch.bfloat16)
    lin.weight.data = H.make_weights(N, K)
    if use_bias:
        lin.bias.data = (torch.randn(N) * 0.1).to(torch.bfloat16)
    q = QLinearLayer(lin, 256, 256, idx)
    sd = q.state_dict()
    for name in PACKED_NAMES + ("reorder_index",):
        assert name in sd, f"{name} must be a persistent buffer"
    path = tmp_path / "lin.pt"
    q.save_packed(path)
    q2 = QLinearLayer.from_packed(str(path))
    for name in PACKED_NAMES:
        assert torch.equal(getattr(q, name), getattr(q2, name))
    x = H.make_activations(200, K, idx).reshape(1, 200, K).to(cuda)
    assert torch.equal(q(x), q2(x))
    bad = q.packed_state()
    bad["BN"] = bad["BN"][:, :-1]
    with pytest.raises(ValueError):
        QLinearLayer.from_packed(bad)
