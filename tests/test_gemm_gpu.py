"""GPU parity tests for the three-segment mixed-MX GEMM (mixedgemm.matmul -> C ABI -> persistent tcgen05 kernel).

Tolerance (BASELINE.json north_star): on bf16 outputs, max <= 1e-2 and mean <= 1e-3 of
|got-ref| / max(|ref|, rms(ref)) against the fake-quant dequant-matmul oracle (fp32 accumulate, one rounding).
The reference's own three-launch chain rounds to bf16 after every segment (w4a6.cu:178); against that chained
oracle the bound is a couple of bf16 ulps (max <= 2e-2).
"""
import numpy as np
import pytest
import torch

import helpers as H

O = H.O
pytestmark = pytest.mark.gpu
TOL_MAX, TOL_MEAN = 1e-2, 1e-3


def _quantize(cuda, x, w, idx, split, sym):
    from micromix_b200 import mixedgemm
    a = mixedgemm.reorder_quantize_x(x.to(cuda), idx.to(cuda), *split)
    fn = mixedgemm.reorder_quantize_w if sym else mixedgemm.reorder_quantize_w4
    b = fn(w.to(cuda), idx.to(cuda), *split)
    return a, b


def _mm(a, b, **kw):
    from micromix_b200 import mixedgemm
    return mixedgemm.matmul(a[0], b[0], a[1], b[1], a[2], b[2], a[3], b[3], a[4], b[4], a[5], b[5], **kw)


def _oracle(a, b, chain=False):
    an, bn = [H.u8(t) for t in a], [H.u8(t) for t in b]
    return O.matmul(an[0], bn[0], an[1], bn[1], an[2], bn[2], an[3], bn[3], an[4], bn[4], an[5], bn[5], chain=chain)


def _check(cuda, M, N, split, sym=False, seed=0):
    K = sum(split)
    idx = H.make_index(K, seed=seed)
    x, w = H.make_activations(M, K, idx, seed=721 + seed), H.make_weights(N, K, seed=1234 + seed)
    a, b = _quantize(cuda, x, w, idx, split, sym)
    c = _mm(a, b)
    torch.cuda.synchronize()
    assert c.shape == (M, N) and c.dtype == torch.bfloat16
    mx, mean = H.rel_err(H.bits(c), _oracle(a, b))
    assert mx <= TOL_MAX and mean <= TOL_MEAN, (mx, mean)
    return a, b, c


@pytest.mark.parametrize("split", [(256, 0, 0), (128, 0, 0), (0, 128, 0), (0, 0, 128), (384, 0, 0), (0, 384, 0),
                                   (0, 0, 384), (128, 128, 128), (640, 256, 128), (2560, 1024, 512)])
@pytest.mark.parametrize("sym", [False, True])
def test_segments_single_tile(cuda, split, sym):
    if sym and split[1] == 0 and split[2] == 0:
        pytest.skip("symmetric and w4 coincide without FP6/FP8 segments")
    _check(cuda, 128, 256, split, sym, seed=sum(split))


@pytest.mark.parametrize("M", [1, 2, 17, 64, 127, 128, 129, 200, 256, 300, 1000])
def test_ragged_m(cuda, M):
    _check(cuda, M, 512, (256, 128, 128), seed=M)


@pytest.mark.parametrize("N", [128, 256, 384, 512, 640, 1024, 1152])
def test_n_tails(cuda, N):
    _check(cuda, 130, N, (384, 128, 128), seed=N)


@pytest.mark.parametrize("M,N,K", [(512, 6144, 4096), (512, 4096, 4096), (256, 28672, 4096), (512, 4096, 14336),
                                   (256, 1024, 4096), (384, 5120, 5120)])
def test_llama_and_qwen_shapes(cuda, M, N, K):
    _check(cuda, M, N, H.SPLITS[K], seed=K + N)


def test_config1_q_proj_2048(cuda):
    """BASELINE config 1: q_proj 4096x4096, 2048 tokens, split (2560,1024,512); both oracle flavours."""
    a, b, c = _check(cuda, 2048, 4096, (2560, 1024, 512))
    mx, mean = H.rel_err(H.bits(c), _oracle(a, b, chain=True))
    assert mx <= 2e-2 and mean <= TOL_MEAN, (mx, mean)


def test_bias_epilogue_matches_separate_add(cuda):
    M, N, split = 200, 512, (256, 128, 128)
    K = sum(split)
    idx = H.make_index(K, seed=4)
    x, w = H.make_activations(M, K, idx), H.make_weights(N, K)
    a, b = _quantize(cuda, x, w, idx, split, False)
    bias = (torch.randn(N, device=cuda) * 0.5).to(torch.bfloat16)
    fused = _mm(a, b, bias=bias)
    separate = _mm(a, b) + bias  # model/qLinearLayer.py:70-71
    assert torch.equal(fused, separate)


def test_out_argument_and_no_prezero_needed(cuda):
    M, N, split = 150, 256, (0, 128, 128)  # KN == 0: the reference needs its zero-filled C here (gemm.cu:75-77)
    K = sum(split)
    idx = H.make_index(K, seed=6)
    a, b = _quantize(cuda, H.make_activations(M, K, idx), H.make_weights(N, K), idx, split, False)
    out = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device=cuda)
    r = _mm(a, b, out=out)
    assert r.data_ptr() == out.data_ptr() and not torch.isnan(out).any()
    assert torch.equal(out, _mm(a, b))


def test_linearity_in_scales_full_size(cuda):
    """Size-independent property at a BASELINE size: adding 1 to every activation scale byte doubles the output
    exactly (powers of two commute with the fp32 accumulation and the bf16 rounding)."""
    M, N, K = 8192, 4096, 4096
    split = H.SPLITS[K]
    idx = H.make_index(K).to(cuda)
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(M, K, generator=g, device=cuda).to(torch.bfloat16)
    w = (torch.randn(N, K, generator=g, device=cuda) * 0.02).to(torch.bfloat16)
    from micromix_b200 import mixedgemm
    a = list(mixedgemm.reorder_quantize_x(x, idx, *split))
    b = mixedgemm.reorder_quantize_w4(w, idx, *split)
    c1 = _mm(a, b)
    for i in (3, 4, 5):
        a[i] = a[i] + 1
    c2 = _mm(a, b)
    assert torch.equal(c2.float(), c1.float() * 2)
    # and a few rows against the oracle
    rows = torch.tensor([0, 1, 4097, 8191], device=cuda)
    xs = x[rows].contiguous()
    a_s = mixedgemm.reorder_quantize_x(xs, idx, *split)
    ref = _oracle(a_s, b)
    mx, mean = H.rel_err(H.bits(c1[rows]), ref)
    assert mx <= TOL_MAX and mean <= TOL_MEAN


def test_permutation_invariance(cuda):
    """A @ B^T is unchanged when both operands use another channel permutation inside a segment-preserving
    relabelling: quantising with idx and with idx composed with a within-group rotation gives the same product."""
    M, N, K = 96, 256, 512
    split = (256, 128, 128)
    idx = H.make_index(K, seed=8)
    x, w = H.make_activations(M, K, idx), H.make_weights(N, K)
    idx2 = idx.view(-1, 32).roll(5, dims=1).reshape(-1).contiguous()  # same groups, rotated inside each group
    a1, b1 = _quantize(cuda, x, w, idx, split, False)
    a2, b2 = _quantize(cuda, x, w, idx2, split, False)
    # same products, possibly summed in another order inside the tensor core: at most rare one-ulp flips
    mx, mean = H.rel_err(H.bits(_mm(a1, b1)), H.bits(_mm(a2, b2)))
    assert mx <= 8e-3 and mean <= 1e-4, (mx, mean)


def test_determinism_and_watchdog_build(cuda, mmx_lib):
    a, b, c = _check(cuda, 300, 1024, (2560, 1024, 512), seed=3)
    for _ in range(3):
        assert torch.equal(_mm(a, b), c)
    mmx_lib.mmx_set_option(b"gemm_watchdog", 1)
    try:
        c2 = _mm(a, b)
        torch.cuda.synchronize()
        import ctypes
        buf = (ctypes.c_uint32 * 8)()
        assert mmx_lib.mmx_gemm_debug_status(buf, 8) == 0
        assert all(v == 0 for v in buf)
        assert torch.equal(c2, c)
    finally:
        mmx_lib.mmx_set_option(b"gemm_watchdog", 0)


def test_cuda_graph_capture(cuda):
    from micromix_b200 import mixedgemm
    M, N, split = 256, 512, (256, 128, 128)
    K = sum(split)
    idx = H.make_index(K, seed=10).to(cuda)
    x = H.make_activations(M, K, idx.cpu()).to(cuda)
    w = H.make_weights(N, K).to(cuda)
    b = mixedgemm.reorder_quantize_w4(w, idx, *split)
    eager = _mm(mixedgemm.reorder_quantize_x(x, idx, *split), b)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        gph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gph, stream=s):
            a = mixedgemm.reorder_quantize_x(x, idx, *split)
            y = _mm(a, b)
    x.copy_(x * 2)
    gph.replay()
    torch.cuda.synchronize()
    assert torch.equal(y.float(), eager.float() * 2)


def test_errors(cuda):
    from micromix_b200 import mixedgemm
    u = lambda *s: torch.zeros(*s, dtype=torch.uint8, device=cuda)
    with pytest.raises(ValueError):  # N not a multiple of 128
        mixedgemm.matmul(u(4, 64), u(100, 64), u(4, 0), u(100, 0), u(4, 0), u(100, 0), u(512), u(512), u(0), u(0), u(0), u(0))
    with pytest.raises(ValueError):  # SF buffer too small
        mixedgemm.matmul(u(4, 64), u(128, 64), u(4, 0), u(128, 0), u(4, 0), u(128, 0), u(16), u(512), u(0), u(0), u(0), u(0))


@pytest.mark.parametrize("sk", [0, 2, 4, 8])
@pytest.mark.parametrize("M,N,split", [(1, 4096, (2560, 1024, 512)), (16, 1024, (384, 128, 128)), (128, 512, (640, 256, 128)),
                                       (77, 384, (2688, 0, 1408)), (5, 256, (0, 1024, 0)), (128, 4096, (8960, 3584, 1792)),
                                       (300, 512, (640, 256, 128)), (512, 1024, (2560, 1024, 512))])
def test_splitk_decode_shapes(cuda, mmx_lib, sk, M, N, split):
    """M <= 512: split-K over a CTA cluster (DSMEM reduction in CTA order).  Against the oracle within the GEMM tolerance,
    against the unsplit kernel within fp32 re-association (a couple of bf16 steps), deterministic run to run, with and
    without bias.  sk = 0 is the automatic choice."""
    K = sum(split)
    idx = H.make_index(K, seed=sk + M)
    x, w = H.make_activations(M, K, idx, seed=31 + M), H.make_weights(N, K, seed=47 + N)
    a, b = _quantize(cuda, x, w, idx, split, False)
    bias = (torch.randn(N, device=cuda) * 0.5).to(torch.bfloat16)
    try:
        mmx_lib.mmx_set_option(b"gemm_splitk", 1)
        c1, c1b = _mm(a, b), _mm(a, b, bias=bias)
        mmx_lib.mmx_set_option(b"gemm_splitk", sk)
        c, cb, c_again = _mm(a, b), _mm(a, b, bias=bias), _mm(a, b)
        torch.cuda.synchronize()
    finally:
        mmx_lib.mmx_set_option(b"gemm_splitk", 0)
    mx, mean = H.rel_err(H.bits(c), _oracle(a, b))
    assert mx <= TOL_MAX and mean <= TOL_MEAN, (mx, mean)
    assert torch.equal(c, c_again)
    mx1, _ = H.rel_err(H.bits(c), H.bits(c1))
    mxb, _ = H.rel_err(H.bits(cb), H.bits(c1b))
    assert mx1 <= TOL_MAX and mxb <= TOL_MAX, (mx1, mxb)


# ---------------------------------------------------------------------------------------------- bench-size oracle checks
def _edge_rows(M, n_extra=4, seed=0):
    """>= 8 rows including every kind of tile edge: first / last row of the first and last 128- and 256-row tiles."""
    base = {0, 127, 128, 255, 256, M - 257, M - 256, M - 129, M - 128, M - 1, M // 2}
    g = torch.Generator().manual_seed(seed)
    base |= set(torch.randint(0, M, (n_extra,), generator=g).tolist())
    return sorted(r for r in base if 0 <= r < M)


def _big_case(cuda, M, N, K, split, seed):
    """Full-size GEMM on the GPU, sampled rows against the CPU oracle (the rows' codes are re-quantized on their own:
    quantization is per row, so they are the same bytes)."""
    from micromix_b200 import mixedgemm
    idx = H.make_index(K, seed=seed).to(cuda)
    g = torch.Generator(device="cuda").manual_seed(100 + seed)
    gain = 1.0 + 31.0 * (torch.arange(K, device=cuda, dtype=torch.float32) / K) ** 8
    x = torch.randn(M, K, generator=g, device=cuda)
    xg = torch.empty_like(x)
    xg[:, idx.long()] = x * gain
    x = xg.to(torch.bfloat16)
    del xg
    w = (torch.randn(N, K, generator=g, device=cuda) * 0.02).to(torch.bfloat16)
    a = mixedgemm.reorder_quantize_x(x, idx, *split)
    b = mixedgemm.reorder_quantize_w4(w, idx, *split)
    c = _mm(a, b)
    rows = torch.tensor(_edge_rows(M, seed=seed), device=cuda)
    a_s = mixedgemm.reorder_quantize_x(x[rows].contiguous(), idx, *split)
    for i in range(3):  # the sampled rows' codes are exactly the rows of the full quantization
        assert torch.equal(a_s[i], a[i][rows])
    ref = _oracle(a_s, b)
    mx, mean = H.rel_err(H.bits(c[rows]), ref)
    assert mx <= TOL_MAX and mean <= TOL_MEAN, (M, N, K, mx, mean)
    return c


LLAMA_LINEARS = [("qkv", 6144, 4096), ("o", 4096, 4096), ("gate_up", 28672, 4096), ("down", 4096, 14336)]


@pytest.mark.parametrize("M", [8192, 16384])
@pytest.mark.parametrize("name,N,K", LLAMA_LINEARS)
def test_llama_linears_at_bench_sizes_vs_oracle(cuda, M, name, N, K):
    """all four Llama-3-8B linears at the benchmarked M = 8192 and the prefill M = 16384, >= 8 sampled rows
    including tile edges, against the fake-quant oracle within the north_star tolerance."""
    _big_case(cuda, M, N, K, H.SPLITS[K], seed=M // 8192 + N % 7)


def _shard_split(K_local):
    p8 = max(128, (K_local // 8 + 127) // 128 * 128)
    p6 = max(128, (K_local // 4 + 127) // 128 * 128)
    return K_local - p6 - p8, p6, p8


@pytest.mark.parametrize("N,K", [(4096, 512), (4096, 1792), (768, 4096), (3584, 4096),   # Llama-3-8B at tp = 8
                                 (5120, 3456), (5120, 6912), (5120, 13824), (5120, 640),  # Qwen2.5-32B down / o at tp 8/4/2
                                 (6912, 5120), (896, 5120)])                              # ... gate_up / qkv at tp = 8
def test_tp_shard_shapes_vs_oracle(cuda, N, K):
    """The per-rank GEMM shapes of the tensor-parallel runs (K_local 512 / 1792, N 768 / 3584, the Qwen K = 27648 shards)
    at M = 8192 against the oracle: short-K and narrow-N tiles take other code paths than the square prefill shapes."""
    _big_case(cuda, 8192, N, K, _shard_split(K), seed=K % 11 + N % 5)


# ---- fused SiLU(gate) * up + MX quantize in the GEMM epilogue (mmx_matmul_activate_quantize): bit-identical to the two ops
@pytest.mark.parametrize("M,K,dsplit", [(128, 512, (128, 0, 0)), (64, 512, (256, 128, 128)), (300, 1024, (256, 128, 128)),
                                        (129, 512, (384, 0, 128)), (1000, 1024, (1024, 512, 256)),
                                        (2048, 4096, (9216, 3584, 1536)), (515, 640, (640, 256, 128))])
def test_matmul_activate_quantize_matches_two_ops(cuda, M, K, dsplit):
    from micromix_b200 import mixedgemm
    inter = sum(dsplit)
    p8 = (K // 8) // 128 * 128
    p6 = (K // 4) // 128 * 128
    split = (K - p6 - p8, p6, p8)
    idx = H.make_index(K, seed=M + K)
    x = H.make_activations(M, K, idx, seed=5 + M)
    wg, wu = H.make_weights(inter, K, seed=77 + M), H.make_weights(inter, K, seed=78 + M)
    a = mixedgemm.reorder_quantize_x(x.to(cuda), idx.to(cuda), *split)
    # the two separate ops on the plain [gate; up] weight
    b = mixedgemm.reorder_quantize_w4(torch.cat([wg, wu]).to(cuda), idx.to(cuda), *split)
    y = _mm(a, b)
    ref = mixedgemm.activate_quantize_x(y[:, :inter], y[:, inter:], *dsplit)
    # the fused kernel on the interleaved weight
    bi = mixedgemm.reorder_quantize_w4(mixedgemm.interleave_gate_up(wg, wu).to(cuda), idx.to(cuda), *split)
    got = mixedgemm.matmul_activate_quantize(a[0], bi[0], a[1], bi[1], a[2], bi[2], a[3], bi[3], a[4], bi[4], a[5], bi[5],
                                             *dsplit)
    torch.cuda.synchronize()
    for i in range(3):
        assert torch.equal(got[i], ref[i]), f"codes of segment {i} differ"
    for i, k in enumerate(dsplit):
        g, r = H.u8(got[3 + i]), H.u8(ref[3 + i])
        assert g.shape == r.shape
        valid = -(-M // 128) * (k // 128) * 512  # whole row blocks, padding rows included (0x7F); the rest is never written
        assert np.array_equal(g[:valid], r[:valid]), f"scale bytes of segment {i} differ"


def test_matmul_activate_quantize_special_values_and_errors(cuda):
    """Zero rows of X (all-zero groups -> scale byte 0x7F), large gate magnitudes (silu saturation on both sides)."""
    from micromix_b200 import mixedgemm
    M, K, dsplit = 256, 512, (256, 128, 128)
    inter = sum(dsplit)
    split = (256, 128, 128)
    idx = H.make_index(K, seed=3)
    x = H.make_activations(M, K, idx, seed=9)
    x[5] = 0
    x[200:210] *= 64.0
    wg, wu = H.make_weights(inter, K, seed=1) * 8.0, H.make_weights(inter, K, seed=2)
    a = mixedgemm.reorder_quantize_x(x.to(cuda), idx.to(cuda), *split)
    b = mixedgemm.reorder_quantize_w(torch.cat([wg, wu]).to(cuda), idx.to(cuda), *split)
    y = _mm(a, b)
    ref = mixedgemm.activate_quantize_x(y[:, :inter].contiguous(), y[:, inter:].contiguous(), *dsplit)
    bi = mixedgemm.reorder_quantize_w(mixedgemm.interleave_gate_up(wg, wu).to(cuda), idx.to(cuda), *split)
    got = mixedgemm.matmul_activate_quantize(a[0], bi[0], a[1], bi[1], a[2], bi[2], a[3], bi[3], a[4], bi[4], a[5], bi[5],
                                             *dsplit)
    torch.cuda.synchronize()
    for i in range(3):
        assert torch.equal(got[i], ref[i]), i
    for i, k in enumerate(dsplit):
        valid = -(-M // 128) * (k // 128) * 512
        assert torch.equal(got[3 + i][:valid], ref[3 + i][:valid]), i
    with pytest.raises(ValueError):
        mixedgemm.matmul_activate_quantize(a[0], bi[0], a[1], bi[1], a[2], bi[2], a[3], bi[3], a[4], bi[4], a[5], bi[5],
                                           256, 128, 0)


def test_fast_silu_equals_reference_sequence_on_every_bf16_in_range(cuda, mmx_lib):
    """The epilogue's branch-free silu against the reference instruction sequence, exhaustively: all bf16 gate values with
    2^-60 <= |x| <= 32 (outside that range the epilogue itself falls back to the reference sequence)."""
    fast = torch.empty(65536, dtype=torch.int32, device=cuda)
    ref = torch.empty(65536, dtype=torch.int32, device=cuda)
    rc = mmx_lib.mmx_debug_silu_table(fast.data_ptr(), ref.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    bits = torch.arange(65536, device=cuda)
    mag = bits & 0x7fff
    inr = (mag >= 0x2180) & (mag <= 0x4200)
    assert int(inr.sum()) == 2 * (0x4200 - 0x2180 + 1)
    bad = inr & (fast != ref)
    assert int(bad.sum()) == 0, [hex(int(b)) for b in bits[bad][:8]]


@pytest.mark.parametrize("M,N,split,bias", [(1, 512, (256, 128, 128), False), (64, 1152, (640, 256, 128), True),
                                            (200, 512, (256, 128, 128), False), (1000, 1024, (640, 256, 128), True),
                                            (515, 4096, (2560, 1024, 512), False),
                                            # N % 256 == 128: the last tile's upper column blocks do not exist (single-CTA
                                            # and pair kernels; several tiles per CTA so both accumulators are exercised)
                                            (200, 1152, (256, 128, 128), False), (1000, 640, (256, 128, 128), True),
                                            (5000, 384, (256, 0, 0), False)])
def test_residual_epilogue_matches_separate_add(cuda, M, N, split, bias):
    """matmul(..., residual=r) == r + matmul(...) (torch's bf16 add), bit for bit: the decoder layer's residual adds in the
    GEMM epilogue, split-K (M <= 128), single-CTA and pair kernels, with and without the bias epilogue, in place too."""
    K = sum(split)
    idx = H.make_index(K, seed=N)
    x, w = H.make_activations(M, K, idx, seed=3 + M), H.make_weights(N, K, seed=4 + M)
    a, b = _quantize(cuda, x, w, idx, split, False)
    g = torch.Generator(device=cuda).manual_seed(M)
    r = torch.randn(M, N, generator=g, device=cuda, dtype=torch.float32).to(torch.bfloat16)
    bv = (torch.randn(N, generator=g, device=cuda, dtype=torch.float32) * 0.1).to(torch.bfloat16) if bias else None
    want = r + _mm(a, b, bias=bv)
    got = _mm(a, b, bias=bv, residual=r)
    torch.cuda.synchronize()
    assert torch.equal(got, want)
    r2 = r.clone()
    _mm(a, b, bias=bv, residual=r2, out=r2)
    torch.cuda.synchronize()
    assert torch.equal(r2, want)


@pytest.mark.parametrize("M,S,nh,nkv,extra,bias", [(1, 1, 4, 2, 256, False), (96, 48, 4, 2, 256, True), (300, 300, 2, 2, 128, False),
                                                   (1000, 250, 6, 2, 1024, True), (515, 103, 3, 1, 0, False)])
def test_rope_epilogue_matches_gemm_then_rope_inplace(cuda, M, S, nh, nkv, extra, bias):
    """matmul(..., rope=) on pair-adjacent q / k weight rows == matmul on the plain rows followed by rope_inplace (itself
    bit-identical to HF's apply_rotary_pos_emb, tests/test_models_gpu.py): split-K (M <= 128), single-CTA and pair kernels,
    with and without a bias, v columns untouched, table rows reused modulo S."""
    from micromix_b200 import mixedgemm
    d, K, split = 128, 512, (256, 128, 128)
    nrope = (nh + nkv) * d
    N = nrope + extra
    idx = H.make_index(K, seed=M)
    x, w = H.make_activations(M, K, idx, seed=11 + M), H.make_weights(N, K, seed=12 + M)
    g = torch.Generator(device=cuda).manual_seed(M + 1)
    bv = (torch.randn(N, generator=g, device=cuda, dtype=torch.float32) * 0.1).to(torch.bfloat16) if bias else None
    ang = torch.rand(S, d // 2, generator=g, device=cuda) * 6.28
    cos = torch.cat([ang.cos(), ang.cos()], -1).to(torch.bfloat16).contiguous()
    sin = torch.cat([ang.sin(), (ang * 1.01).sin()], -1).to(torch.bfloat16).contiguous()  # (not symmetric: both halves are read)
    a, b = _quantize(cuda, x, w, idx, split, False)
    want = _mm(a, b, bias=bv)
    mixedgemm.rope_inplace(want, nh + nkv, d, cos, sin)
    wp = torch.cat([mixedgemm.pair_adjacent_rows(w[:nrope], nh + nkv, d), w[nrope:]])
    bp = None if bv is None else torch.cat([mixedgemm.pair_adjacent_rows(bv[:nrope], nh + nkv, d), bv[nrope:]])
    _, b2 = _quantize(cuda, x, wp, idx, split, False)
    got = _mm(a, b2, bias=bp, rope=(cos, sin, nrope))
    torch.cuda.synchronize()
    assert torch.equal(got[:, nrope:], want[:, nrope:])
    assert torch.equal(got, want), f"{(got.float() - want.float()).abs().max().item()}"
