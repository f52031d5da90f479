"""Grouped expert path of the Mixtral block (BASELINE config 5): grouped quantize + grouped GEMM + combine kernel against
the per-expert loop -- the reference's op sequence, /root/reference/model/qMixtralLayer.py:437-450 -- bit for bit."""
import pytest
import torch
import torch.nn.functional as F

import helpers as H

pytestmark = pytest.mark.gpu


def _sorted_problem(cuda, groups, rows_per_group, K, seed=0):
    """Rows sorted by group, each group padded to 128 rows; returns (x_all, row_src, grp_rowblk, spans)."""
    g = torch.Generator().manual_seed(seed)
    T = 3 * sum(rows_per_group)
    x = (torch.randn(T, K, generator=g) * 2).to(torch.bfloat16).to(cuda)
    row_src, grp, spans = [], [], []
    off = 0
    for gi, n in enumerate(rows_per_group):
        tok = torch.randperm(T, generator=g)[:n].sort().values
        pad = (n + 127) // 128 * 128
        row_src += tok.tolist() + [0] * (pad - n)
        grp += [gi] * (pad // 128)
        spans.append((off, n, tok))
        off += pad
    return x, torch.tensor(row_src, dtype=torch.int32, device=cuda), torch.tensor(grp, dtype=torch.int32, device=cuda), spans, off


@pytest.mark.parametrize("K,split", [(1024, (512, 256, 256)), (4096, (2560, 1024, 512)), (2048, (2048, 0, 0))])
def test_grouped_quantize_matches_per_group(cuda, K, split):
    from micromix_b200 import mixedgemm
    rows = [200, 1, 128, 300]
    idx = torch.stack([H.make_index(K, seed=10 + i) for i in range(len(rows))]).to(cuda)
    x, row_src, grp, spans, Mp = _sorted_problem(cuda, len(rows), rows, K)
    got = mixedgemm.reorder_quantize_x_grouped(x, idx, grp, *split, row_src=row_src, rows=Mp)
    torch.cuda.synchronize()
    for gi, (off, n, tok) in enumerate(spans):
        xe = torch.zeros(((n + 127) // 128 * 128, K), dtype=torch.bfloat16, device=cuda)
        xe[:n] = x[tok.to(cuda)]
        ref = mixedgemm.reorder_quantize_x(xe, idx[gi].contiguous(), *split)
        for s in range(3):
            assert torch.equal(got[s][off:off + n], ref[s][:n]), f"group {gi} segment {s} codes"
            if split[s]:
                # scale atoms of the group's 128-row blocks: rows < n
                per_blk = 128 * split[s] // 32
                gsf = got[3 + s][off // 128 * per_blk:][: ref[3 + s].numel()]
                m = torch.from_numpy(H.O.sf_valid_mask(n, split[s], ref[3 + s].numel())).to(cuda)
                assert torch.equal(gsf[: m.numel()][m], ref[3 + s][m]), f"group {gi} segment {s} scales"


@pytest.mark.parametrize("tile", [128, 256])
def test_grouped_gemm_matches_per_group(cuda, tile):
    from micromix_b200 import mixedgemm
    K, N, split = 1024, 512, (512, 256, 256)
    rows = [300, 0, 70, 513]
    G = len(rows)
    idx = H.make_index(K, seed=3).to(cuda)
    Ws = [mixedgemm.reorder_quantize_w4(H.make_weights(N, K, seed=50 + i).to(cuda), idx, *split) for i in range(G)]
    Wst = tuple(torch.cat([w[c] for w in Ws], 0).contiguous() for c in range(6))
    pads = [(n + tile - 1) // tile * tile for n in rows]
    Mp = sum(pads) + 2 * tile  # two unused m-tiles at the end
    x = H.make_activations(Mp, K, idx.cpu(), seed=9).to(cuda)
    A = mixedgemm.reorder_quantize_x(x, idx, *split)
    gm, off = [], 0
    for gi, p in enumerate(pads):
        gm += [gi] * (p // tile)
    gm += [-1, -1]
    gm = torch.tensor(gm, dtype=torch.int32, device=cuda)
    out = torch.full((Mp, N), 7.0, dtype=torch.bfloat16, device=cuda)
    mixedgemm.matmul_grouped(A, Wst, gm, G, tile, out=out)
    full = [mixedgemm.matmul(A[0], w[0], A[1], w[1], A[2], w[2], A[3], w[3], A[4], w[4], A[5], w[5]) for w in Ws]
    torch.cuda.synchronize()
    off = 0
    for gi, p in enumerate(pads):
        assert torch.equal(out[off:off + p], full[gi][off:off + p]), f"group {gi}"
        off += p
    assert bool((out[off:] == 7.0).all()), "padding m-tiles must not be written"


def test_moe_combine_matches_index_add_loop(cuda):
    from micromix_b200 import mixedgemm
    T, k, E, Hd = 333, 2, 8, 1024
    g = torch.Generator().manual_seed(1)
    sel = torch.stack([torch.randperm(E, generator=g)[:k] for _ in range(T)]).to(cuda)
    w = torch.rand(T, k, generator=g).to(torch.bfloat16).to(cuda)
    rows = torch.randperm(T * k, generator=g).view(T, k).to(torch.int32).to(cuda)
    y = torch.randn(T * k, Hd, generator=g).to(torch.bfloat16).to(cuda)
    local = sel % 2 == 0  # pretend odd experts live elsewhere
    rows_l = torch.where(local, rows, torch.full_like(rows, -1))
    got = mixedgemm.moe_combine(y, rows_l, sel.to(torch.int32), w)
    want = torch.zeros(T, Hd, dtype=torch.bfloat16, device=cuda)
    for e in range(E):
        if e % 2:
            continue
        tok, slot = torch.where(sel == e)
        cur = y[rows[tok, slot].long()] * w[tok, slot, None]
        want.index_add_(0, tok, cur)
    torch.cuda.synchronize()
    assert torch.equal(got, want)


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("tokens", [37, 700])
def test_moe_block_grouped_equals_expert_loop(cuda, fused, tokens):
    from micromix_b200 import model_shapes as S
    from micromix_b200.qMixtralLayer import QMixtralSparseMoeBlock
    cfg = dict(S.MIXTRAL_8X7B, hidden_size=1024, intermediate_size=2048, num_attention_heads=8, num_key_value_heads=2,
               num_local_experts=4)
    layer = S.make_layer(cfg, cuda, seed=3, moe=True)
    idx, p6, p8 = S.make_calibration(cfg, 0, moe=True)
    loop = QMixtralSparseMoeBlock(layer.block_sparse_moe, p8, p6, idx, 0, fused=fused, grouped=False)
    grp = QMixtralSparseMoeBlock(layer.block_sparse_moe, p8, p6, idx, 0, fused=fused)
    assert grp.grouped
    g = torch.Generator(device=cuda).manual_seed(4)
    x = torch.randn(1, tokens, cfg["hidden_size"], generator=g, device=cuda).to(torch.bfloat16)
    a, la = loop(x)
    b, lb = grp(x)
    torch.cuda.synchronize()
    assert torch.equal(la, lb)
    assert torch.isfinite(b.float()).all()
    assert torch.equal(a, b), f"max diff {(a.float() - b.float()).abs().max().item()}"
    if fused:
        # the grouped GEMM's SiLU * up + quantize epilogue against the two separate kernels (grouped GEMM -> activate_quantize_x)
        two = QMixtralSparseMoeBlock(layer.block_sparse_moe, p8, p6, idx, 0, fused=True, act_epilogue=False)
        assert grp.act_epilogue and not two.act_epilogue
        c, _ = two(x)
        torch.cuda.synchronize()
        assert torch.equal(c, b)


@pytest.mark.parametrize("T,k,E,local,tile", [(1, 2, 8, [0, 1, 2, 3, 4, 5, 6, 7], 128), (37, 2, 8, [3], 128),
                                              (700, 2, 8, [1, 5], 256), (16384, 2, 8, [0, 1, 2, 3, 4, 5, 6, 7], 256),
                                              (5000, 3, 16, [2, 3, 9, 15], 128), (300, 1, 4, [0, 2], 128)])
def test_moe_route_kernel_equals_the_torch_tables(cuda, T, k, E, local, tile):
    """mixedgemm.moe_route (one kernel) == qMixtralLayer.route_tables (stable argsort + scatters + cumsums), table by table."""
    from micromix_b200 import mixedgemm
    from micromix_b200.qMixtralLayer import route_tables
    g = torch.Generator(device=cuda).manual_seed(T + k)
    logits = torch.randn(T, E, generator=g, device=cuda)
    sel = torch.topk(logits, k, dim=-1).indices
    slot = torch.full((E,), len(local), dtype=torch.int64)
    for s_, j in enumerate(local):
        slot[j] = s_
    slot = slot.to(cuda)
    want = route_tables(sel, slot, len(local), tile)
    got = mixedgemm.moe_route(sel, slot, len(local), tile)
    torch.cuda.synchronize()
    assert got[4] == want[4]
    used = int(want[5].item())
    assert int(got[5].item()) == used
    assert torch.equal(got[1], want[1])                                  # pair_row
    assert torch.equal(got[0][:used], want[0][:used])                    # row_src (rows in use; the rest is never read)
    assert torch.equal(got[2][: used // 128], want[2][: used // 128])    # grp_rowblk
    assert torch.equal(got[3], want[3])                                  # grp_mtile (-1 past the rows in use)
