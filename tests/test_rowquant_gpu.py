"""GPU parity tests for the ops that quantize without a permutation -- activate_quantize_x, downproj_quantize_w,
downproj_quantize_w4 (activate.cu) -- and for rmsnorm_quantize_x (rmsnorm.cu), through mixedgemm -> ctypes -> C ABI.

Bars
  downproj_quantize_w / _w4   codes + scale bytes BIT-EXACT against the oracle (its log2f is CUDA's, restated)
  rmsnorm_quantize_x          BIT-EXACT against the oracle (fixed-order sum of squares, IEEE sqrt / reciprocal)
  activate_quantize_x         BIT-EXACT against the reference's own kernel (committed golden outputs from a B200, and the
                              kernel live when oracle/_ref/libref_activate.so travelled with the snapshot).  Against the
                              CPU oracle the comparison brackets the fp32 product by +-1e-6 relative, because CUDA's
                              expf ends in the hardware ex2.approx and cannot be restated on a CPU bit for bit.
"""
import ctypes
import os

import numpy as np
import pytest
import torch

import helpers as H

O = H.O
pytestmark = pytest.mark.gpu
FMT_X, FMT_W4 = (4, 6, 8), (4, 4, 4)


def _mg():
    from micromix_b200 import mixedgemm
    return mixedgemm


def _assert_exact(got, ref, M, ks, what=""):
    for i in range(3):
        g = H.u8(got[i])
        assert g.shape == ref[i].shape, (what, i, g.shape, ref[i].shape)
        assert np.array_equal(g, ref[i]), f"{what}: codes of segment {i} differ"
    for i, k in enumerate(ks):
        g = H.u8(got[3 + i])
        assert g.shape == ref[3 + i].shape
        m = O.sf_valid_mask(M, k, g.shape[0])
        assert np.array_equal(g[m], ref[3 + i][m]), f"{what}: scale bytes of segment {i} differ"


# ---------------------------------------------------------------------------------------------- downproj (exact)
@pytest.mark.parametrize("N,K,split", [
    (128, 128, (128, 0, 0)), (128, 128, (0, 128, 0)), (128, 128, (0, 0, 128)),
    (1, 384, (128, 128, 128)), (130, 640, (384, 128, 128)), (257, 4096, (2560, 1024, 512)),
    (300, 14336, (8960, 3584, 1792)), (64, 4224, (2688, 1024, 512)), (1000, 1024, (0, 512, 512)),
    (96, 1792, (1152, 384, 256)), (40, 27648, (17280, 6912, 3456)),
])
def test_downproj_bit_exact(cuda, N, K, split):
    w = H.make_weights(N, K, seed=N + K)
    wb = H.bits(w)
    for w4, op in ((False, _mg().downproj_quantize_w), (True, _mg().downproj_quantize_w4)):
        got = op(w.to(cuda), *split)
        torch.cuda.synchronize()
        _assert_exact(got, O.downproj_quantize(wb, *split, w4), N, split, f"w4={w4}")


def test_downproj_zero_tiny_negative_zero(cuda):
    K = 256
    wb = np.zeros((6, K), dtype=np.uint16)
    wb[1, :] = 0x8000          # -0.0 everywhere: scale 1.0 (0x7F), sign-bit codes
    wb[2, ::3] = 0x3580        # 9.5e-7 < 1e-6: the "tiny" rule gives scale 1.0 and all-zero codes
    wb[3, 5] = 0x3590          # 1.07e-6 > 1e-6: the group is scaled normally
    wb[4, :] = 0x3F80
    wb[5, 17] = 0x7F00         # 1.7e38
    w = H.from_bits(wb)
    for split in ((K, 0, 0), (0, K, 0), (0, 0, K), (128, 128, 0)):
        for w4, op in ((False, _mg().downproj_quantize_w), (True, _mg().downproj_quantize_w4)):
            got = op(w.to(cuda), *split)
            ref = O.downproj_quantize(wb, *split, w4)
            _assert_exact(got, ref, 6, split)
            k0 = next(k for k in split if k)
            first = [t for t, k in zip(got[3:], split) if k][0]
            assert int(first[O.sf_offset(0, 0, k0)]) == 0x7F  # all-zero group: scale 1.0, unlike reorder.cu's 0.5


def test_downproj_every_bf16_value(cuda):
    """Every finite bf16 value in the range where the recipe is defined, in shuffled and in magnitude-sorted groups."""
    vals = np.arange(65536, dtype=np.uint16)
    f = O.bf16_bits_to_f32(vals)
    vals = vals[np.isfinite(f) & ((np.abs(f) >= 2.0 ** -100) | (f == 0)) & (np.abs(f) < 2.0 ** 100)]
    rng = np.random.default_rng(3)
    K = 128
    rows = []
    for _ in range(2):
        v = vals.copy()
        rng.shuffle(v)
        rows.append(np.concatenate([v, np.zeros((-v.size) % K, dtype=np.uint16)]).reshape(-1, K))
    s = vals[np.argsort(np.abs(O.bf16_bits_to_f32(vals)), kind="stable")]
    rows.append(np.concatenate([s, np.zeros((-s.size) % K, dtype=np.uint16)]).reshape(-1, K))
    wb = np.concatenate(rows, axis=0)
    w = H.from_bits(wb).to(cuda)
    for split in ((K, 0, 0), (0, K, 0), (0, 0, K)):
        got = _mg().downproj_quantize_w(w, *split)
        _assert_exact(got, O.downproj_quantize(wb, *split, False), wb.shape[0], split)


# ---------------------------------------------------------------------------------------------- activate
_gate_up = H.make_gate_up


def _check_activate_bracket(got, gate, up, split, M):
    """got must equal the oracle quantizer applied to SOME fp32 product within 1e-6 (relative) of the oracle's."""
    v = O.silu_mul(H.bits(gate), H.bits(up))
    d = np.float32(1e-6)
    mid = O.quantize_f32(v, *split)
    lo = O.quantize_f32(v * (1 - d), *split)
    hi = O.quantize_f32(v * (1 + d), *split)
    n_groups = n_unstable = n_elems = n_off_mid = 0
    for i, (k, fmt) in enumerate(zip(split, FMT_X)):
        if k == 0:
            continue
        sg, slo, shi, smid = H.u8(got[3 + i]), lo[3 + i], hi[3 + i], mid[3 + i]
        r, g = np.arange(M)[:, None], np.arange(k // 32)[None, :]
        off = O.sf_offset(r, g, k)
        assert np.all((sg[off] == slo[off]) | (sg[off] == shi[off])), f"segment {i}: a scale byte is outside the bracket"
        stable = (slo[off] == shi[off])  # [M, groups]
        n_groups += stable.size
        n_unstable += int((~stable).sum())
        dg = O.dequant(H.u8(got[i]), sg, M, k, fmt)
        dlo, dhi, dmid = (O.dequant(t[i], t[3 + i], M, k, fmt) for t in (lo, hi, mid))
        el = np.repeat(stable, 32, axis=1)
        inside = (dg >= np.minimum(dlo, dhi)) & (dg <= np.maximum(dlo, dhi))
        assert np.all(inside[el]), f"segment {i}: a code is outside the bracket"
        n_elems += int(el.sum())
        n_off_mid += int((dg != dmid)[el].sum())
        assert np.array_equal(smid[off][stable], sg[off][stable])
    assert n_unstable <= max(2, n_groups // 1000), (n_unstable, n_groups)
    assert n_off_mid <= max(4, n_elems // 2000), (n_off_mid, n_elems)


@pytest.mark.parametrize("M,K,split", [(1, 128, (128, 0, 0)), (33, 384, (128, 128, 128)), (300, 4096, (2560, 1024, 512)),
                                       (129, 14336, (8960, 3584, 1792)), (70, 4224, (2688, 1024, 512)),
                                       (256, 1024, (0, 0, 1024)), (256, 1024, (0, 1024, 0))])
def test_activate_against_oracle_bracket(cuda, M, K, split):
    gate, up = _gate_up(M, K, seed=M + K)
    got = _mg().activate_quantize_x(gate.to(cuda), up.to(cuda), *split)
    torch.cuda.synchronize()
    assert got[0].shape == (M, split[0] // 2) and got[1].shape == (M, split[1] // 4 * 3) and got[2].shape == (M, split[2])
    _check_activate_bracket(got, gate, up, split, M)


def test_activate_equals_unfused_pipeline_semantics(cuda):
    """Dequantised activate_quantize_x(gate, up) ~= silu(gate) * up (fp32): the MX rounding error bound per format."""
    M, K, split = 200, 1024, (512, 256, 256)
    gate, up = _gate_up(M, K, seed=9)
    got = _mg().activate_quantize_x(gate.to(cuda), up.to(cuda), *split)
    ref = (torch.nn.functional.silu(gate.float()) * up.float()).numpy()
    c0 = 0
    for i, (k, fmt, rel) in enumerate(zip(split, FMT_X, (0.25, 0.125, 0.0625))):
        d = O.dequant(H.u8(got[i]), H.u8(got[3 + i]), M, k, fmt)
        r = ref[:, c0:c0 + k]
        gmax = np.abs(r).reshape(M, k // 32, 32).max(axis=2, keepdims=True)
        err = np.abs(d - r).reshape(M, k // 32, 32)
        assert np.all(err <= rel * gmax + 1e-6), (i, float((err / (gmax + 1e-30)).max()))
        c0 += k


# ---------------------------------------------------------------------------------------------- reference kernel
ROWQ_GOLDEN = H.ROWQ_GOLDEN
rowq_golden_inputs = H.rowq_golden_inputs


def _our_rowq(mode):
    mg = _mg()
    return (mg.activate_quantize_x, mg.downproj_quantize_w, mg.downproj_quantize_w4)[mode]


@pytest.mark.parametrize("tag", list(ROWQ_GOLDEN))
def test_committed_reference_kernel_outputs(cuda, tag):
    """tests/golden/ref_rowquant_golden.npz: outputs of the reference's own activate.cu kernels on a B200
    (tools/make_golden_rowquant.py)."""
    path = os.path.join(H.ROOT, "tests", "golden", "ref_rowquant_golden.npz")
    if not os.path.exists(path):
        pytest.fail("tests/golden/ref_rowquant_golden.npz is missing")
    g = np.load(path)
    mode, M, split = ROWQ_GOLDEN[tag]
    ins = rowq_golden_inputs(tag)
    got = _our_rowq(mode)(*[t.to(cuda) for t in ins], *split)
    for i, k in enumerate(split):
        assert np.array_equal(H.u8(got[i]), g[f"{tag}_q{i}"]), f"{tag}: codes of segment {i} differ from the reference"
        sfg = H.u8(got[3 + i])
        m = O.sf_valid_mask(M, k, sfg.shape[0])
        assert np.array_equal(sfg[m], g[f"{tag}_sf{i}"][m]), f"{tag}: scales of segment {i} differ from the reference"


def load_ref_rowq():
    so = os.path.join(H.ROOT, "oracle", "_ref", "libref_activate.so")
    if not os.path.exists(so):
        return None
    R = ctypes.CDLL(so)
    R.ref_rowwise_quantize.argtypes = ([ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int] + [ctypes.c_int] * 3 +
                                       [ctypes.c_void_p] * 6)
    return R


def run_ref_rowq(R, mode, ins, M, split, dev):
    fm = FMT_W4 if mode == 2 else FMT_X
    q = [torch.zeros((M, k * f // 8), dtype=torch.uint8, device=dev) for k, f in zip(split, fm)]
    sf = [torch.zeros((O.sf_bytes(M, k, True),), dtype=torch.uint8, device=dev) for k in split]
    torch.cuda.synchronize()
    a = ins[0].data_ptr()
    b = ins[1].data_ptr() if len(ins) > 1 else None
    rc = R.ref_rowwise_quantize(mode, a, b, M, *split, *[t.data_ptr() for t in q], *[t.data_ptr() for t in sf])
    torch.cuda.synchronize()
    assert rc == 0, rc
    return q, sf


def test_live_reference_kernel_if_present(cuda):
    """oracle/_ref/libref_activate.so = the reference's activate.cu compiled in place for sm_100a (git-ignored)."""
    R = load_ref_rowq()
    if R is None:
        pytest.skip("reference kernel library not present (needs /root/reference at build time)")
    for mode, M, split in [(0, 2048, (2560, 1024, 512)), (0, 515, (8704, 3584, 2048)), (0, 1024, (0, 0, 1024)),
                           (1, 513, (2560, 1024, 512)), (2, 513, (8704, 3584, 2048)), (1, 100, (3072, 1536, 512))]:
        K = sum(split)
        ins = _gate_up(M, K, seed=31 + M) if mode == 0 else (H.make_weights(M, K, seed=M),)
        ins = [t.to(cuda) for t in ins]
        q, sf = run_ref_rowq(R, mode, ins, M, split, cuda)
        ours = _our_rowq(mode)(*ins, *split)
        for i, k in enumerate(split):
            assert torch.equal(q[i], ours[i]), (mode, M, split, i)
            m = torch.from_numpy(O.sf_valid_mask(M, k, sf[i].numel())).to(cuda)
            assert torch.equal(sf[i][m], ours[3 + i][m]), (mode, M, split, i)


def test_live_reference_scale_boundaries_if_present(cuda):
    """amax / QMAX a few ulps above a power of two is where log2f decides the scale: products silu(a) * b whose group
    maxima sweep those neighbourhoods must still match the reference kernel bit for bit."""
    R = load_ref_rowq()
    if R is None:
        pytest.skip("reference kernel library not present (needs /root/reference at build time)")
    M, split = 4096, (512, 512, 512)
    K = sum(split)
    g = torch.Generator().manual_seed(5)
    # gate large (silu(a) ~ a) and up = exact powers of two times QMAX-ish factors: many maxima land near 2^k * QMAX
    gate = (6.0 + torch.rand(M, K, generator=g) * 0.05).to(torch.bfloat16)
    up = torch.ones(M, K).to(torch.bfloat16)
    up[:, 512:1024] = 28.0 / 6.0
    up[:, 1024:] = 448.0 / 6.0
    ins = [gate.to(cuda), up.to(cuda)]
    q, sf = run_ref_rowq(R, 0, ins, M, split, cuda)
    ours = _mg().activate_quantize_x(*ins, *split)
    for i, k in enumerate(split):
        assert torch.equal(q[i], ours[i])
        m = torch.from_numpy(O.sf_valid_mask(M, k, sf[i].numel())).to(cuda)
        assert torch.equal(sf[i][m], ours[3 + i][m])


# ---------------------------------------------------------------------------------------------- rmsnorm + quantize
def _norm_weight(K, seed):
    g = torch.Generator().manual_seed(seed)
    return (1.0 + 0.1 * torch.randn(K, generator=g)).to(torch.bfloat16)


@pytest.mark.parametrize("M,K,split,eps", [
    (1, 128, (128, 0, 0), 1e-5), (130, 384, (128, 128, 128), 1e-6), (257, 3072, (2048, 512, 512), 1e-5),
    (300, 4096, (2560, 1024, 512), 1e-5), (100, 5120, (3200, 1280, 640), 1e-6), (96, 4224, (2688, 1024, 512), 1e-5),
    (70, 8192, (5120, 2048, 1024), 1e-5), (40, 14336, (8960, 3584, 1792), 1e-5), (33, 3584, (2304, 896, 384), 1e-6),
])
def test_rmsnorm_quantize_bit_exact(cuda, M, K, split, eps):
    idx = H.make_index(K, seed=K + 1)
    x = H.make_activations(M, K, idx, seed=M)
    w = _norm_weight(K, seed=K)
    got = _mg().rmsnorm_quantize_x(x.to(cuda), w.to(cuda), eps, idx.to(cuda), *split)
    torch.cuda.synchronize()
    ref = O.rmsnorm_quantize(H.bits(x), H.bits(w), eps, idx.numpy(), *split)
    _assert_exact(got, ref, M, split)


def test_rmsnorm_quantize_equals_norm_then_quantize(cuda):
    """The fused op == the oracle's normalised bf16 rows pushed through the plain reorder_quantize_x kernel."""
    M, K, split = 515, 4096, (2560, 1024, 512)
    idx = H.make_index(K, seed=4)
    x = H.make_activations(M, K, idx, seed=8)
    w = _norm_weight(K, seed=2)
    fused = _mg().rmsnorm_quantize_x(x.to(cuda), w.to(cuda), 1e-5, idx.to(cuda), *split)
    y = H.from_bits(O.rmsnorm(H.bits(x), H.bits(w), 1e-5)).to(cuda)
    two = _mg().reorder_quantize_x(y, idx.to(cuda), *split)
    for i, k in enumerate(split):
        assert torch.equal(fused[i], two[i])
        m = torch.from_numpy(O.sf_valid_mask(M, k, fused[3 + i].numel())).to(cuda)
        assert torch.equal(fused[3 + i][m], two[3 + i][m])
    # and the oracle's norm is an RMSNorm: close to the plain fp32 formula
    xf = x.float()
    t = (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5) * w.float()).numpy()
    yo = O.bf16_bits_to_f32(O.rmsnorm(H.bits(x), H.bits(w), 1e-5))
    assert np.all(np.abs(yo - t) <= 2.0 ** -7 * np.abs(t) + 1e-30)


def test_full_size_properties(cuda):
    """BASELINE sizes: row independence + a sampled oracle check at M=16384 (activate K=14336, rmsnorm K=4096)."""
    mg = _mg()
    rows = torch.tensor([0, 1, 127, 128, 5000, 16383], device=cuda)
    g = torch.Generator(device="cuda").manual_seed(3)
    K, split = 14336, H.SPLITS[14336]
    gate = torch.randn(16384, K, generator=g, device=cuda).to(torch.bfloat16)
    up = torch.randn(16384, K, generator=g, device=cuda).to(torch.bfloat16)
    full = mg.activate_quantize_x(gate, up, *split)
    sub = mg.activate_quantize_x(gate[rows].contiguous(), up[rows].contiguous(), *split)
    for i in range(3):
        assert torch.equal(full[i][rows], sub[i])
    _check_activate_bracket(sub, gate[rows].cpu(), up[rows].cpu(), split, rows.numel())
    K, split = 4096, H.SPLITS[4096]
    idx = H.make_index(K).to(cuda)
    x = torch.randn(16384, K, generator=g, device=cuda).to(torch.bfloat16)
    w = _norm_weight(K, 1).to(cuda)
    full = mg.rmsnorm_quantize_x(x, w, 1e-5, idx, *split)
    sub = mg.rmsnorm_quantize_x(x[rows].contiguous(), w, 1e-5, idx, *split)
    for i in range(3):
        assert torch.equal(full[i][rows], sub[i])
    ref = O.rmsnorm_quantize(H.bits(x[rows]), H.bits(w), 1e-5, idx.cpu().numpy(), *split)
    _assert_exact(sub, ref, rows.numel(), split)


def test_errors(cuda):
    mg = _mg()
    a = torch.zeros(4, 256, dtype=torch.bfloat16, device=cuda)
    idx = torch.arange(256, dtype=torch.int16, device=cuda)
    w = torch.ones(256, dtype=torch.bfloat16, device=cuda)
    with pytest.raises(ValueError):
        mg.activate_quantize_x(a, a, 100, 100, 56)
    with pytest.raises(ValueError):
        mg.activate_quantize_x(a, a[:2], 256, 0, 0)
    with pytest.raises(ValueError):
        mg.downproj_quantize_w(a.float(), 256, 0, 0)
    with pytest.raises(RuntimeError):
        mg.downproj_quantize_w4(a.cpu(), 256, 0, 0)
    with pytest.raises(ValueError):
        mg.rmsnorm_quantize_x(a, w[:128], 1e-5, idx, 256, 0, 0)
    with pytest.raises(ValueError):
        mg.rmsnorm_quantize_x(a, w, 1e-5, idx, 128, 0, 0)
    out = mg.activate_quantize_x(a[:0], a[:0], 256, 0, 0)
    assert out[0].shape == (0, 128)
