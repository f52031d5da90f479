"""CPU tests: pin the oracle (oracle/mmx_oracle.c) to outputs of the reference's own code.

  * tests/golden/cutlass_convert_table.npz -- the vendored CUTLASS NumericConverter / CuTe SF layout run on the host
    (tools/make_golden_cutlass.py), i.e. the element and layout semantics reorder.cu:138-143,182-185 relies on;
  * tests/golden/ref_reorder_golden.npz    -- the reference's reorder.cu kernels run on a B200
    (tools/make_golden_ref.py);
  * the known-answer table of SURVEY.md section 8a-7.
"""
import os

import numpy as np
import pytest

import helpers as H

O = H.O
GOLD = os.path.join(H.ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def table():
    return np.load(os.path.join(GOLD, "cutlass_convert_table.npz"))


@pytest.mark.parametrize("fmt,key", [(4, "e2m1"), (6, "e3m2"), (8, "e4m3")])
def test_element_encode_matches_cutlass_for_every_bf16(table, fmt, key):
    L = O.lib()
    vals = O.bf16_bits_to_f32(np.arange(65536, dtype=np.uint16))
    finite = np.isfinite(vals)
    mine = np.array([L.mmxo_encode(float(v), fmt) for v in vals], dtype=np.uint8)
    assert np.array_equal(mine[finite], table[key][finite])


def test_ue8m0_matches_cutlass(table):
    L = O.lib()
    vals = O.bf16_bits_to_f32(np.arange(65536, dtype=np.uint16))
    nonneg = ~np.signbit(vals)
    mine = np.array([L.mmxo_ue8m0_from_float(float(v)) for v in vals], dtype=np.uint8)
    assert np.array_equal(mine[nonneg], table["ue8m0"][nonneg])


def test_sf_layout_matches_cute(table):
    r, g = np.arange(300)[:, None], np.arange(32)[None, :]
    assert np.array_equal(O.sf_offset(r, g, 1024), table["sfa_off_M300_K1024"])
    r, g = np.arange(384)[:, None], np.arange(20)[None, :]
    assert np.array_equal(O.sf_offset(r, g, 640), table["sfb_off_N384_K640"])
    L = O.lib()
    for (rr, gg, k) in [(0, 0, 128), (299, 31, 1024), (129, 5, 640), (31, 3, 128), (32, 4, 256)]:
        assert L.mmxo_sf_offset(rr, gg, k) == int(O.sf_offset(rr, gg, k))


def test_known_answers_survey_8a7():
    """SURVEY.md 8a-7 table (probed from the CUTLASS host converters)."""
    L = O.lib()
    kat4 = {0.25: 0x0, 0.75: 0x2, 1.25: 0x2, 1.75: 0x4, 2.5: 0x4, 3.5: 0x6, 5.0: 0x6, -0.25: 0x8, 6.0: 0x7, 100.0: 0x7}
    for v, c in kat4.items():
        assert L.mmxo_encode(v, 4) == c, (v, c)
    assert L.mmxo_encode(27.0, 6) == 0x1F and L.mmxo_encode(0.03, 6) == 0x00
    assert L.mmxo_encode(500.0, 8) == 0x7E
    assert L.mmxo_ue8m0_from_float(3.0) == 0x81 and L.mmxo_ue8m0_from_float(0.0) == 0x00
    assert L.mmxo_ue8m0_from_float(0.5) == 0x7E  # all-zero group scale (reorder.cu:179)
    # decode(encode(v)) is the identity on each format's grid
    for fmt, n in ((4, 16), (6, 64), (8, 256)):
        for code in range(n):
            v = L.mmxo_decode(code, fmt)
            if v == v:
                assert L.mmxo_encode(v, fmt) == code


def test_scale_rule_exhaustive():
    """ceil(log2(amax/QMAX)) in float (the reference recipe, reorder.cu:180,192,204) == integer exponent arithmetic
    (what the CUDA kernel does) for EVERY positive bf16 amax, subnormals included."""
    L = O.lib()
    for fmt in (4, 6, 8):
        for b in range(0x0001, 0x7F80):
            assert L.mmxo_scale_byte_float(b, fmt) == L.mmxo_scale_byte_int(b, fmt), (hex(b), fmt)
        assert L.mmxo_scale_byte_float(0, fmt) == 126 == L.mmxo_scale_byte_int(0, fmt)


def test_packing_layouts():
    K = 128
    idx = np.arange(K, dtype=np.int16)
    # one row whose reordered values are exactly representable: codes are predictable
    x = np.zeros((1, K), dtype=np.float32)
    x[0, :8] = [0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0, -6.0]
    xb = O.f32_to_bf16_bits(x)
    qn, _, _, sfn, _, _ = O.reorder_quantize(xb, idx, K, 0, 0, "x")
    assert sfn[0] == 127  # amax 6 -> scale 1
    assert list(qn[0, :4]) == [0x21, 0x43, 0x65, 0xF7]  # even element in the low nibble
    _, qs, _, _, sfs, _ = O.reorder_quantize(xb, idx, 0, K, 0, "x")
    assert sfs[0] == 125  # amax 6 -> 2^ceil(log2(6/28)) = 2^-2
    codes = [O.lib().mmxo_encode(float(v) * 4.0, 6) for v in x[0, :4]]
    w = codes[0] | codes[1] << 6 | codes[2] << 12 | codes[3] << 18
    assert list(qs[0, :3]) == [w & 0xFF, (w >> 8) & 0xFF, (w >> 16) & 0xFF]
    _, _, qo, _, _, sfo = O.reorder_quantize(xb, idx, 0, 0, K, "x")
    assert sfo[0] == 121  # 2^ceil(log2(6/448)) = 2^-6
    assert qo[0, 0] == O.lib().mmxo_encode(0.5 * 64, 8)


def test_zero_group_and_negative_zero():
    K = 128
    idx = np.arange(K, dtype=np.int16)
    xb = np.zeros((2, K), dtype=np.uint16)
    xb[1, 0] = 0x8000  # -0.0
    for split, i in (((K, 0, 0), 0), ((0, K, 0), 1), ((0, 0, K), 2)):
        out = O.reorder_quantize(xb, idx, *split, "x")
        assert out[3 + i][int(O.sf_offset(0, 0, K))] == 0x7E
        assert not out[i][0].any()
        sign = {0: 0x08, 1: 0x20, 2: 0x80}[i]
        assert out[i][1][0] == sign


@pytest.mark.parametrize("tag", list(H.GOLDEN_CASES))
def test_oracle_reproduces_reference_kernel_golden(tag):
    """Outputs of the reference's own reorder.cu on a B200 (committed) == oracle on the same seeded inputs."""
    g = H.load_golden()
    M, K, (KN, KS, KO) = H.GOLDEN_CASES[tag]
    x, idx = H.golden_inputs(tag)
    for mode in ("x", "w", "w4"):
        if f"{tag}_{mode}_q0" not in g.files:
            continue
        o = O.reorder_quantize(H.bits(x), idx.numpy(), KN, KS, KO, mode)
        for i, k in enumerate((KN, KS, KO)):
            assert np.array_equal(o[i], g[f"{tag}_{mode}_q{i}"]), (tag, mode, i)
            m = O.sf_valid_mask(M, k, o[3 + i].shape[0])
            assert np.array_equal(o[3 + i][m], g[f"{tag}_{mode}_sf{i}"][m]), (tag, mode, i)


def test_dequant_roundtrip_error_bounds():
    M, K = 64, 1024
    idx = H.make_index(K, seed=5)
    x = H.make_activations(M, K, idx)
    xr = O.bf16_bits_to_f32(H.bits(x))[:, idx.numpy().astype(np.int64)]
    for split, fmt, rel in (((K, 0, 0), 4, 0.25), ((0, K, 0), 6, 0.125), ((0, 0, K), 8, 0.0625)):
        out = O.reorder_quantize(H.bits(x), idx.numpy(), *split, "x")
        i = (4, 6, 8).index(fmt)
        d = O.dequant(out[i], out[3 + i], M, K, fmt)
        amax = np.abs(xr).reshape(M, K // 32, 32).max(-1, keepdims=True).repeat(32, -1).reshape(M, K)
        assert np.all(np.abs(d - xr) <= rel * amax + 1e-30)


def test_matmul_oracle_chain_vs_fused_and_identity_scale():
    M, N, K = 48, 128, 384
    idx = H.make_index(K, seed=9)
    x, w = H.make_activations(M, K, idx), H.make_weights(N, K)
    a = O.reorder_quantize(H.bits(x), idx.numpy(), 128, 128, 128, "x")
    b = O.reorder_quantize(H.bits(w), idx.numpy(), 128, 128, 128, "w4")
    args = (a[0], b[0], a[1], b[1], a[2], b[2], a[3], b[3], a[4], b[4], a[5], b[5])
    chain, fused = O.matmul(*args, chain=True), O.matmul(*args, chain=False)
    mx, mean = H.rel_err(chain, fused)
    assert mx <= 2e-2 and mean <= 1e-3  # inter-segment bf16 roundings of the reference: a couple of bf16 ulps
    f64 = O.matmul(*args, chain=False, f64=True)
    assert H.rel_err(fused, f64)[0] <= 8e-3  # at most one bf16 ulp from fp32 vs fp64 accumulation


# ---------------------------------------------------------------- activate.cu / rmsnorm.cu restatements
def test_cuda_log2f_restatement():
    """mmxo_cuda_log2f restates the PTX nvcc 12.9 emits for log2f: exact on powers of two, within 2 ulp elsewhere."""
    import math
    L = O.lib()
    for e in range(-126, 128):
        assert L.mmxo_cuda_log2f(float(2.0 ** e)) == float(e)
    rng = np.random.default_rng(0)
    xs = np.exp(rng.uniform(-60, 60, 4000)).astype(np.float32)
    for x in xs:
        got, want = L.mmxo_cuda_log2f(float(x)), math.log2(float(x))
        assert abs(got - want) <= 2.5 * np.spacing(np.float32(abs(want) + 1e-30)), (x, got, want)
    # KATs read off a B200 run of the reference's kernel are in tests/golden/ref_rowquant_golden.npz (scale bytes)


def test_scale_exponent_shortcut_is_exact():
    """The CUDA kernel reads ceil(log2(amax/QMAX)) off the exponent unless the ratio is within 2^-13 above a power of
    two; the polynomial must agree there for every exponent (65536 mantissas next to each end of the range)."""
    assert O.lib().mmxo_check_scale_shortcut(-100, 128) == 0
    L = O.lib()
    rng = np.random.default_rng(1)
    for fmt in (4, 6, 8):
        for a in np.exp(rng.uniform(-13, 80, 3000)).astype(np.float32):
            assert L.mmxo_act_scale_exp(float(a), fmt) == L.mmxo_act_scale_exp_fast(float(a), fmt)


def test_quantize_f32_agrees_with_reorder_oracle_on_bf16_inputs():
    """On bf16 inputs with an identity permutation the activate.cu recipe and the reorder.cu recipe coincide wherever
    the group maximum is > 1e-6 (they differ only in the all-zero / tiny rule: scale 1.0 vs 0.5)."""
    M, K, split = 64, 1024, (512, 256, 256)
    idx = H.make_index(K, identity=True)
    x = H.make_activations(M, K, idx)
    a = O.reorder_quantize(H.bits(x), idx.numpy(), *split, "x")
    b = O.quantize_f32(O.bf16_bits_to_f32(H.bits(x)), *split)
    for i in range(3):
        assert np.array_equal(a[i], b[i])
    for i, k in enumerate(split):
        m = O.sf_valid_mask(M, k, a[3 + i].shape[0])
        assert np.array_equal(a[3 + i][m], b[3 + i][m])
    z = O.quantize_f32(np.zeros((1, 128), dtype=np.float32), 128, 0, 0)
    assert z[3][0] == 0x7F and not z[0].any()


def test_rmsnorm_oracle_is_an_rmsnorm():
    import torch
    M, K = 16, 4096
    x = H.make_activations(M, K, H.make_index(K))
    g = torch.Generator().manual_seed(1)
    w = (1.0 + 0.1 * torch.randn(K, generator=g)).to(torch.bfloat16)
    y = O.bf16_bits_to_f32(O.rmsnorm(H.bits(x), H.bits(w), 1e-5))
    xf = x.float()
    t = (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5) * w.float()).numpy()
    assert np.all(np.abs(y - t) <= 2.0 ** -7 * np.abs(t) + 1e-30)
    # the tree sum is exact on data whose squares add without rounding
    ones = np.full((1, 256), 0x3F80, dtype=np.uint16)
    y1 = O.rmsnorm(ones, ones[0], 0.0)
    assert np.all(y1 == 0x3F80)
