"""Host-side weight layouts of the fused epilogues (pure tensor code, no GPU): interleave_gate_up (matmul_activate_quantize)
and pair_adjacent_rows (matmul(..., rope=))."""
import pytest
import torch

from micromix_b200 import mixedgemm


def test_interleave_gate_up_blocks_of_128():
    inter, K = 384, 8
    gate = torch.arange(inter * K, dtype=torch.float32).reshape(inter, K)
    up = -gate
    w = mixedgemm.interleave_gate_up(gate, up)
    assert w.shape == (2 * inter, K) and w.is_contiguous()
    for t in range(inter // 128):
        assert torch.equal(w[256 * t:256 * t + 128], gate[128 * t:128 * t + 128])
        assert torch.equal(w[256 * t + 128:256 * t + 256], up[128 * t:128 * t + 128])
    b = mixedgemm.interleave_gate_up(torch.arange(inter), torch.arange(inter) + 1000)  # 1-D (biases)
    assert b[:128].tolist() == list(range(128)) and b[128:256].tolist() == list(range(1000, 1128))
    with pytest.raises(ValueError):
        mixedgemm.interleave_gate_up(gate[:100], up[:100])
    with pytest.raises(ValueError):
        mixedgemm.interleave_gate_up(gate, up[:256])


def test_pair_adjacent_rows_and_inverse():
    heads, d = 3, 128
    w = torch.arange(heads * d * 2).reshape(heads * d, 2)
    p = mixedgemm.pair_adjacent_rows(w, heads, d)
    for h in range(heads):
        for j in range(d // 2):
            assert torch.equal(p[h * d + 2 * j], w[h * d + j])
            assert torch.equal(p[h * d + 2 * j + 1], w[h * d + j + d // 2])
    assert torch.equal(mixedgemm.pair_adjacent_rows(p, heads, d, inverse=True), w)
    v = torch.arange(heads * d)
    assert torch.equal(mixedgemm.pair_adjacent_rows(mixedgemm.pair_adjacent_rows(v, heads, d), heads, d, inverse=True), v)
    with pytest.raises(ValueError):
        mixedgemm.pair_adjacent_rows(w, heads + 1, d)
