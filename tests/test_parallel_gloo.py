"""CPU tests of the N>1 path: shard planning + the reduce step under a world_size-2 gloo group.

The per-rank compute stand-in is the oracle's fake-quant linear (the CUDA layers cannot run here); what is under
test is the host logic of micromix_b200/parallel_utils.py: which rows/channels a rank owns, its rank-local
permutation and split, and that summing the partials over the group reproduces the unsharded sharded-math result.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as H

O = H.O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_row_shard_plan_properties():
    from micromix_b200.parallel_utils import row_shard_plan, column_shard_range
    for K, (p4, p6, p8), tp in ((14336, (8960, 3584, 1792), 8), (4096, (2560, 1024, 512), 4), (27648, (17280, 6912, 3456), 2)):
        idx = H.make_index(K, seed=K)
        seen = []
        for r in range(tp):
            k0, k1, lidx, q4, q6, q8 = row_shard_plan(idx, p6, p8, tp, r)
            assert (k0, k1) == (r * K // tp, (r + 1) * K // tp)
            assert q4 + q6 + q8 == K // tp and min(q4, q6, q8) >= 0
            assert q4 % 128 == 0 and q6 % 128 == 0 and q8 % 128 == 0
            li = lidx.to(torch.int64)
            assert sorted(li.tolist()) == list(range(K // tp))  # a permutation of the local slice
            # importance order is preserved: local order is the global order filtered to the slice
            glob = idx.to(torch.int64)
            assert torch.equal(li + k0, glob[(glob >= k0) & (glob < k1)])
            # the split follows the global assignment (random permutation -> close to p/tp)
            assert abs(q8 - p8 / tp) <= 256 and abs(q6 - p6 / tp) <= 256
            seen += (li + k0).tolist()
        assert sorted(seen) == list(range(K))
    assert column_shard_range(6144, 4, 1) == (1536, 3072)
    with pytest.raises(ValueError):
        column_shard_range(1024, 16, 0)
    with pytest.raises(ValueError):
        row_shard_plan(H.make_index(4096), 1024, 512, 64, 0)


def test_token_parallel_plan_is_the_rank_blocked_concatenation():
    """token_parallel_plan: ONE global permutation / split under which a full-K GEMM sees exactly the ranks' local quantization
    groups -- the FP4 blocks of all ranks, then their FP6, then their FP8 blocks, each at the offset the plan reports."""
    from micromix_b200.parallel_utils import row_shard_plan, token_parallel_plan
    for K, (p6, p8), tp in ((4096, (1024, 512), 8), (14336, (3584, 1792), 4), (5120, (1280, 640), 2)):
        idx = H.make_index(K, seed=K + tp)
        perm, tot, shards = token_parallel_plan(idx, p6, p8, tp)
        assert perm.dtype == torch.int16 and sorted(perm.tolist()) == list(range(K))
        assert sum(tot) == K and all(t % 128 == 0 for t in tot) and len(shards) == tp
        seg0 = (0, tot[0], tot[0] + tot[1])
        acc = [0, 0, 0]
        for r, sh in enumerate(shards):
            k0, k1, lidx, q4, q6, q8 = row_shard_plan(idx, p6, p8, tp, r)
            assert (sh["k0"], sh["k1"]) == (k0, k1) and sh["split"] == (q4, q6, q8) and torch.equal(sh["index"], lidx)
            assert sh["offset"] == tuple(acc)
            g = lidx.to(torch.int64) + k0
            parts = (g[:q4], g[q4:q4 + q6], g[q4 + q6:])
            for i in range(3):  # rank r's block of total segment i is its local segment i, in local order
                lo = seg0[i] + sh["offset"][i]
                assert torch.equal(perm[lo:lo + parts[i].numel()].to(torch.int64), parts[i])
                acc[i] += parts[i].numel()
        assert tuple(acc) == tot


def test_row_shard_plan_never_demotes():
    """the rank-local counts are rounded UP -- no globally-FP8 channel may land in a local FP6 / FP4 segment and
    no globally-FP6 channel in the local FP4 segment, at any tp (the stock 5:2:1 split at K=4096, tp=8 has ~64 FP8
    channels per rank: nearest-multiple rounding gave p8 = 0 there)."""
    from micromix_b200.parallel_utils import row_shard_plan
    for K, (p4, p6, p8) in ((4096, (2560, 1024, 512)), (14336, (8960, 3584, 1792)), (5120, (3200, 1280, 640)),
                            (4096, (3968, 0, 128)), (4096, (3840, 128, 128))):
        for tp in (2, 4, 8):
            for seed in (0, 1):
                idx = H.make_index(K, seed=seed)
                glob = idx.to(torch.int64)
                prec = torch.empty(K, dtype=torch.int64)  # bits each ORIGINAL channel gets in the global split
                prec[glob[:p4]] = 4
                prec[glob[p4:p4 + p6]] = 6
                prec[glob[p4 + p6:]] = 8
                for r in range(tp):
                    k0, k1, lidx, q4, q6, q8 = row_shard_plan(idx, p6, p8, tp, r)
                    li = lidx.to(torch.int64) + k0
                    local_bits = torch.cat([torch.full((q4,), 4), torch.full((q6,), 6), torch.full((q8,), 8)])
                    assert bool((local_bits >= prec[li]).all()), (K, tp, r, q4, q6, q8)
                    # ... and the promotion stays below one 128-group per format
                    n8, n6 = int((prec[li] == 8).sum()), int((prec[li] == 6).sum())
                    assert q8 - n8 < 128 and q6 + q8 - (n6 + n8) < 128


def _worker(rank, world, port, K, N, M, split, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from micromix_b200.parallel_utils import row_shard_plan, column_shard_range, all_reduce_sum
        idx = H.make_index(K, seed=77)
        x = H.make_activations(M, K, idx)
        w = H.make_weights(N, K)
        # ---- row parallel: local quantize + local mixed GEMM -> partial; partials summed over the group
        k0, k1, lidx, p4, p6, p8 = row_shard_plan(idx, split[1], split[2], world, rank)
        part = O.fake_quant_linear(H.bits(x[:, k0:k1].contiguous()), H.bits(w[:, k0:k1].contiguous()),
                                   lidx.numpy(), p4, p6, p8, chain=False)
        y = H.from_bits(part).clone()          # bf16 partial, as the GPU layer produces
        all_reduce_sum(y)
        # ---- column parallel: each rank's slice equals the slice of the unsharded result
        n0, n1 = column_shard_range(N, world, rank)
        col = O.fake_quant_linear(H.bits(x), H.bits(w[n0:n1].contiguous()), idx.numpy(), *split, chain=False)
        gathered = [torch.empty(M, n1 - n0, dtype=torch.bfloat16) for _ in range(world)]
        dist.all_gather(gathered, H.from_bits(col))
        if rank == 0:
            ret["row"] = H.bits(y)
            ret["col"] = H.bits(torch.cat(gathered, dim=1))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_row_and_column_parallel():
    from micromix_b200.parallel_utils import row_shard_plan
    K, N, M, split = 1024, 256, 24, (512, 256, 256)
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), K, N, M, split, ret), nprocs=world, join=True)
    idx = H.make_index(K, seed=77)
    x = H.make_activations(M, K, idx)
    w = H.make_weights(N, K)
    # expected row-parallel result: same sharded math in one process, partials rounded to bf16 then summed in bf16
    acc = None
    for r in range(world):
        k0, k1, lidx, p4, p6, p8 = row_shard_plan(idx, split[1], split[2], world, r)
        part = O.fake_quant_linear(H.bits(x[:, k0:k1].contiguous()), H.bits(w[:, k0:k1].contiguous()),
                                   lidx.numpy(), p4, p6, p8, chain=False)
        t = H.from_bits(part)
        acc = t.clone() if acc is None else acc + t
    assert np.array_equal(ret["row"], H.bits(acc))
    # and it approximates the full-precision product as well as the unsharded quantized path does
    full = O.fake_quant_linear(H.bits(x), H.bits(w), idx.numpy(), *split, chain=False)
    exact = O.f32_to_bf16_bits((x.float() @ w.float().T).numpy())
    e_tp, e_1 = H.rel_err(ret["row"], exact)[1], H.rel_err(full, exact)[1]
    assert e_tp <= 1.5 * e_1 + 1e-3
    assert np.array_equal(ret["col"], full)  # column parallel is exactly the unsharded result


def test_peer_workspace_argument_errors_need_no_gpu():
    """Constructor checks that run before any device work."""
    import pytest
    from micromix_b200.parallel_utils import PeerWorkspace
    with pytest.raises(ValueError, match="mode"):
        PeerWorkspace(128, 128, mode="bogus")
    import torch.distributed as dist
    if not dist.is_initialized():
        with pytest.raises(RuntimeError, match="process group"):
            PeerWorkspace(128, 128, mode="push")
