"""GPU parity tests for reorder+quantize (through mixedgemm -> ctypes -> C ABI -> sm_100a kernel).

Bar: codes, packing and E8M0 scale bytes BIT-EXACT against the oracle, against the committed outputs of the
reference's own kernel, and (when oracle/_ref/libref_reorder.so travelled with the snapshot) against that kernel live.
"""
import ctypes
import os

import numpy as np
import pytest
import torch

import helpers as H

O = H.O
pytestmark = pytest.mark.gpu

REF_K = (3072, 3584, 4096, 5120, 8192, 11008, 12288, 13824, 14336, 18944)  # bindings.cpp:134-144


def _ops():
    from micromix_b200 import mixedgemm
    return {"x": mixedgemm.reorder_quantize_x, "w": mixedgemm.reorder_quantize_w, "w4": mixedgemm.reorder_quantize_w4}


def _assert_parity(got, ref, M, ks):
    for i in range(3):
        g = H.u8(got[i])
        assert g.shape == ref[i].shape
        assert np.array_equal(g, ref[i]), f"codes of segment {i} differ"
    for i, k in enumerate(ks):
        g = H.u8(got[3 + i])
        assert g.shape == ref[3 + i].shape
        m = O.sf_valid_mask(M, k, g.shape[0])
        assert np.array_equal(g[m], ref[3 + i][m]), f"scale bytes of segment {i} differ"


def _run_case(cuda, M, K, split, mode, seed=0, x=None, idx=None):
    idx = H.make_index(K, seed=seed) if idx is None else idx
    x = H.make_activations(M, K, idx, seed=721 + seed) if x is None else x
    got = _ops()[mode](x.to(cuda), idx.to(cuda), *split)
    torch.cuda.synchronize()
    ref = O.reorder_quantize(H.bits(x), idx.numpy(), *split, mode)
    _assert_parity(got, ref, M, split)
    return got


@pytest.mark.parametrize("K", REF_K)
def test_every_reference_hidden_size(cuda, K):
    p8 = 128 * max(1, K // 128 // 8)
    p6 = 128 * max(1, K // 128 // 4)
    for mode in ("x", "w4"):
        _run_case(cuda, 160, K, (K - p6 - p8, p6, p8), mode, seed=K)


@pytest.mark.parametrize("M", [1, 7, 31, 32, 33, 127, 128, 129, 255, 256, 257, 1000])
def test_ragged_row_counts(cuda, M):
    for mode in ("x", "w", "w4"):
        _run_case(cuda, M, 1024, (512, 256, 256), mode, seed=M)


@pytest.mark.parametrize("split", [(4096, 0, 0), (0, 4096, 0), (0, 0, 4096), (2560, 1024, 512), (2048, 128, 1920),
                                   (0, 3968, 128), (128, 0, 3968), (3968, 128, 0)])
def test_segment_extremes(cuda, split):
    for mode in ("x", "w", "w4"):
        _run_case(cuda, 130, 4096, split, mode, seed=sum(split) + split[0])


@pytest.mark.parametrize("K,split", [(128, (128, 0, 0)), (128, (0, 128, 0)), (128, (0, 0, 128)),
                                     (384, (128, 128, 128)), (27648, (17280, 6912, 3456)), (30720, (30080, 512, 128)),
                                     (6912, (4352, 1792, 768)), (1792, (1152, 384, 256))])
def test_k_not_in_reference_list(cuda, K, split):
    """TP shards (6912, 1792, ...) and the K the reference cannot run (27648) must work here."""
    _run_case(cuda, 70, K, split, "x", seed=K)


@pytest.mark.parametrize("rows", [4, 2])
def test_both_row_group_variants(cuda, mmx_lib, rows):
    mmx_lib.mmx_set_option(b"quant_rows", rows)
    try:
        _run_case(cuda, 300, 4096, (2560, 1024, 512), "x", seed=rows)
        _run_case(cuda, 140, 14336, (8960, 3584, 1792), "w4", seed=rows)
    finally:
        mmx_lib.mmx_set_option(b"quant_rows", 0)


def test_exhaustive_bf16_values_each_format(cuda):
    """Every finite bf16 value as an element, under many group maxima, in each format: the hardware
    cvt.rn.satfinite path must agree with the reference's software RNE-satfinite path bit for bit."""
    vals = np.arange(65536, dtype=np.uint16)
    f = O.bf16_bits_to_f32(vals)
    vals = vals[np.isfinite(f) & ((np.abs(f) >= 2.0 ** -100) | (f == 0)) & (np.abs(f) < 2.0 ** 100)]
    rng = np.random.default_rng(1)
    K = 128
    n = vals.size
    rows = []
    for rep in range(3):
        v = vals.copy()
        rng.shuffle(v)
        pad = (-n) % K
        v = np.concatenate([v, np.zeros(pad, dtype=np.uint16)])
        rows.append(v.reshape(-1, K))
    # sorted rows: groups of neighbouring magnitudes exercise every rounding boundary below each group maximum
    s = vals[np.argsort(np.abs(O.bf16_bits_to_f32(vals)), kind="stable")]
    s = np.concatenate([s, np.zeros((-n) % K, dtype=np.uint16)]).reshape(-1, K)
    xb = np.concatenate(rows + [s], axis=0)
    x = H.from_bits(xb)
    idx = H.make_index(K, identity=True)
    for split in ((K, 0, 0), (0, K, 0), (0, 0, K)):
        _run_case(cuda, x.shape[0], K, split, "x", x=x, idx=idx)


def test_zero_rows_negative_zero_and_tiny(cuda):
    K = 256
    idx = H.make_index(K, seed=2)
    xb = np.zeros((5, K), dtype=np.uint16)
    xb[1, :] = 0x8000                     # all -0.0
    xb[2, ::2] = 0x0480                   # 2^-118: tiny, but above the range where the reference recipe degenerates
    xb[3, 7] = 0x7F7F                     # largest finite
    xb[4, :] = 0x3F80                     # all ones
    x = H.from_bits(xb)
    for split in ((K, 0, 0), (0, K, 0), (0, 0, K), (128, 128, 0)):
        got = _run_case(cuda, 5, K, split, "x", x=x, idx=idx)
    assert got is not None


@pytest.mark.parametrize("tag", list(H.GOLDEN_CASES))
def test_committed_reference_kernel_outputs(cuda, tag):
    """tests/golden/ref_reorder_golden.npz: outputs of the reference's own reorder.cu on a B200."""
    g = H.load_golden()
    M, K, split = H.GOLDEN_CASES[tag]
    x, idx = H.golden_inputs(tag)
    for mode in ("x", "w", "w4"):
        if f"{tag}_{mode}_q0" not in g.files:
            continue
        got = _ops()[mode](x.to(cuda), idx.to(cuda), *split)
        for i, k in enumerate(split):
            assert np.array_equal(H.u8(got[i]), g[f"{tag}_{mode}_q{i}"])
            sfg = H.u8(got[3 + i])
            m = O.sf_valid_mask(M, k, sfg.shape[0])
            assert np.array_equal(sfg[m], g[f"{tag}_{mode}_sf{i}"][:sfg.shape[0]][m])


def test_live_reference_kernel_if_present(cuda):
    """oracle/_ref/libref_reorder.so = the reference's reorder.cu compiled in place for sm_100a (git-ignored)."""
    so = os.path.join(H.ROOT, "oracle", "_ref", "libref_reorder.so")
    if not os.path.exists(so):
        pytest.skip("reference kernel library not present (needs /root/reference at build time)")
    R = ctypes.CDLL(so)
    R.ref_reorder_quantize.argtypes = ([ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p] +
                                       [ctypes.c_int] * 3 + [ctypes.c_void_p] * 6)
    for (M, K, split) in [(2048, 4096, (2560, 1024, 512)), (513, 14336, (8960, 3584, 1792)), (100, 5120, (3200, 1280, 640))]:
        idx = H.make_index(K, seed=11)
        x = H.make_activations(M, K, idx, seed=5)
        xd, idd = x.to(cuda), idx.to(cuda)
        for mode_i, mode in enumerate(("x", "w", "w4")):
            fm = (4, 4, 4) if mode == "w4" else (4, 6, 8)
            q = [torch.zeros((M, k * f // 8), dtype=torch.uint8, device=cuda) for k, f in zip(split, fm)]
            sf = [torch.zeros((O.sf_bytes(M, k, mode == "x"),), dtype=torch.uint8, device=cuda) for k in split]
            torch.cuda.synchronize()
            rc = R.ref_reorder_quantize(mode_i, xd.data_ptr(), M, idd.data_ptr(), *split,
                                        *[t.data_ptr() for t in q], *[t.data_ptr() for t in sf])
            torch.cuda.synchronize()
            assert rc == 0
            ours = _ops()[mode](xd, idd, *split)
            for i, k in enumerate(split):
                assert torch.equal(q[i], ours[i])
                m = torch.from_numpy(O.sf_valid_mask(M, k, sf[i].numel())).to(cuda)
                assert torch.equal(sf[i][m], ours[3 + i][:sf[i].numel()][m])


def test_full_size_properties(cuda):
    """BASELINE sizes (M=16384, K=4096 / 14336): size-independent properties instead of the (slow) oracle:
    row independence (a row's codes do not depend on which batch it is in) and idempotence of re-quantising the
    dequantised values within each format's grid."""
    from micromix_b200 import mixedgemm
    for K in (4096, 14336):
        split = H.SPLITS[K]
        idx = H.make_index(K).to(cuda)
        g = torch.Generator(device="cuda").manual_seed(3)
        x = torch.randn(16384, K, generator=g, device=cuda, dtype=torch.float32).to(torch.bfloat16)
        full = mixedgemm.reorder_quantize_x(x, idx, *split)
        rows = torch.tensor([0, 1, 127, 128, 5000, 16383], device=cuda)
        sub = mixedgemm.reorder_quantize_x(x[rows].contiguous(), idx, *split)
        for i in range(3):
            assert torch.equal(full[i][rows], sub[i])
        ref = O.reorder_quantize(H.bits(x[rows]), idx.cpu().numpy(), *split, "x")
        for i in range(3):
            assert np.array_equal(H.u8(sub[i]), ref[i])
        # checksum of checksums is launch-invariant
        again = mixedgemm.reorder_quantize_x(x, idx, *split)
        for i, k in enumerate(split):
            assert int(full[i].to(torch.int64).sum()) == int(again[i].to(torch.int64).sum())
            nb = 16384 * k // 32  # M % 128 == 0: the written scale bytes are exactly the first M*k/32 (rest is padding)
            assert int(full[3 + i][:nb].to(torch.int64).sum()) == int(again[3 + i][:nb].to(torch.int64).sum())


def test_errors(cuda):
    from micromix_b200 import mixedgemm
    x = torch.zeros(4, 256, dtype=torch.bfloat16, device=cuda)
    idx = torch.arange(256, dtype=torch.int16, device=cuda)
    with pytest.raises(ValueError):
        mixedgemm.reorder_quantize_x(x, idx, 100, 100, 56)
    with pytest.raises(ValueError):
        mixedgemm.reorder_quantize_x(x, idx, 128, 128, 128)
    with pytest.raises(ValueError):
        mixedgemm.reorder_quantize_x(x.float(), idx, 256, 0, 0)
    with pytest.raises(ValueError):
        mixedgemm.reorder_quantize_x(x, idx.to(torch.int32), 256, 0, 0)
    with pytest.raises(ValueError):
        mixedgemm.reorder_quantize_x(x[:, ::2], idx[:128], 128, 0, 0)
    out = mixedgemm.reorder_quantize_x(x[:0], idx, 256, 0, 0)
    assert out[0].shape == (0, 128)
