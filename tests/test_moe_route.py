"""CPU test of the grouped Mixtral path's device-side routing tables (micromix_b200/qMixtralLayer.py::route_tables):
pure tensor code, checked against the per-expert `torch.where` of the reference's loop (model/qMixtralLayer.py:437-450)."""
import pytest
import torch


@pytest.mark.parametrize("T,k,E,ep,rank,tile", [(1000, 2, 8, 1, 0, 256), (777, 2, 8, 4, 1, 128), (64, 2, 8, 8, 3, 128),
                                                (5, 2, 8, 1, 0, 128), (300, 1, 4, 2, 1, 256), (16, 2, 8, 8, 7, 128)])
def test_route_tables_match_per_expert_where(T, k, E, ep, rank, tile):
    from micromix_b200.qMixtralLayer import route_tables
    g = torch.Generator().manual_seed(T + E)
    sel = torch.stack([torch.randperm(E, generator=g)[:k] for _ in range(T)])
    local = [j for j in range(E) if j % ep == rank]
    slot = torch.full((E,), len(local), dtype=torch.int64)
    for s_, j in enumerate(local):
        slot[j] = s_
    row_src, pair_row, grp_rowblk, grp_mtile, Mp, used = route_tables(sel, slot, len(local), tile)
    assert Mp % tile == 0 and row_src.numel() == Mp and row_src.dtype == torch.int32
    assert grp_rowblk.numel() == Mp // 128 and grp_mtile.numel() == Mp // tile
    assert int(row_src.min()) >= 0 and int(row_src.max()) < T  # padding rows point at a valid token
    off = 0
    for s_, j in enumerate(local):
        tok, sl = torch.where(sel == j)  # tokens ascending: the order of the reference's loop
        n = tok.numel()
        assert torch.equal(row_src[off:off + n].long(), tok)
        assert torch.equal(pair_row[tok, sl].long(), torch.arange(off, off + n))
        pad = (n + tile - 1) // tile * tile
        assert bool((grp_rowblk[off // 128:(off + pad) // 128] == s_).all())
        assert bool((grp_mtile[off // tile:(off + pad) // tile] == s_).all())
        off += pad
    assert off <= Mp and bool((grp_mtile[off // tile:] == -1).all()) and int(used) == off
    assert int(grp_rowblk.min()) >= 0 and int(grp_rowblk.max()) < len(local)
    assert bool((pair_row[slot[sel] == len(local)] == -1).all())  # experts of other ranks
