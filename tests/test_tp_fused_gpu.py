"""Fused row-parallel GEMM -> all-reduce over peer memory (csrc/tp_reduce.cu, mmx_matmul_allreduce).

Single-GPU parity: `tp` ranks are simulated inside ONE process (PeerWorkspace.simulate: the same kernels, counters and
parity protocol with local pointers instead of cudaIpc mappings, one stream per "rank").  The expectation is exact:
every rank's partial is what mmx_matmul writes for its K shard (bf16), the fused path must return
bf16(sum in fp32, in rank order) of those partials -- bit for bit, on every rank.
The real multi-process run (cudaIpc over NVLink, vs NCCL) is tools/tp_fused_check.py under torchrun.
"""
import pytest
import torch

import helpers as H

pytestmark = pytest.mark.gpu


def _shards(tp, M, N, K, seed):
    from micromix_b200 import mixedgemm
    from micromix_b200.parallel_utils import row_shard_plan
    dev = torch.device("cuda:0")
    idx = H.make_index(K, seed=seed)
    x = H.make_activations(M, K, idx, seed=721 + seed).to(dev)
    w = H.make_weights(N, K, seed=1234 + seed).to(dev)
    p8 = (K // 8) // 128 * 128
    p6 = (K // 4) // 128 * 128
    out = []
    for r in range(tp):
        k0, k1, lidx, p4, l6, l8 = row_shard_plan(idx, p6, p8, tp, r)
        lidx = lidx.to(dev)
        a = mixedgemm.reorder_quantize_x(x[:, k0:k1].contiguous(), lidx, p4, l6, l8)
        b = mixedgemm.reorder_quantize_w4(w[:, k0:k1].contiguous(), lidx, p4, l6, l8)
        out.append((a, b))
    return out


def _expected(shards, bias=None):
    from micromix_b200 import mixedgemm
    acc = None
    for r, (a, b) in enumerate(shards):
        part = mixedgemm.matmul(a[0], b[0], a[1], b[1], a[2], b[2], a[3], b[3], a[4], b[4], a[5], b[5],
                                bias=bias if r == 0 else None).float()
        acc = part if acc is None else acc + part
    return acc.to(torch.bfloat16)


CASES = [
    (1, 200, 256, 256),     # degenerate: GEMM -> staging -> reducer -> C
    (2, 128, 256, 512),     # single-CTA kernel (M <= 128), 128-row tiles
    (2, 300, 512, 1024),    # CTA pairs, ragged M
    (4, 1000, 1152, 2048),  # N tail of 128 columns, uneven tile ownership
    (8, 520, 768, 4096),    # every rank of an 8-GPU box
    (2, 2048, 4096, 1024),  # more tiles than reducer CTAs
]


@pytest.mark.parametrize("tp,M,N,K", CASES)
def test_fused_allreduce_matches_sum_of_partials(cuda, mmx_lib, tp, M, N, K):
    from micromix_b200.parallel_utils import PeerWorkspace
    mmx_lib.mmx_set_option(b"tp_reduce_ctas", 8)    # tp reducer grids must be co-resident on the one GPU
    mmx_lib.mmx_set_option(b"tp_timeout_ms", 4000)  # a protocol bug costs a timeout, not the GPU
    try:
        shards = _shards(tp, M, N, K, seed=tp)
        bias = (torch.randn(N, device=cuda) * 0.1).to(torch.bfloat16)
        want = _expected(shards, bias)
        works = PeerWorkspace.simulate(tp, M, N)
        streams = [torch.cuda.Stream() for _ in range(tp)]
        torch.cuda.synchronize()
        for call in range(3):  # both parities, and the counters handed back by call 0 are reused by call 2
            ys = []
            for r in range(tp):
                with torch.cuda.stream(streams[r]):
                    ys.append(works[r].matmul_allreduce(*shards[r], bias=bias if r == 0 else None))
            torch.cuda.synchronize()
            for r in range(tp):
                assert works[r].status() == 0, f"rank {r}: cross-rank wait timed out (call {call})"
                assert torch.equal(ys[r], want), f"rank {r} call {call}: fused result differs from the sum of partials"
        for w in works:
            w.close()
    finally:
        mmx_lib.mmx_set_option(b"tp_reduce_ctas", 0)
        mmx_lib.mmx_set_option(b"tp_timeout_ms", 10000)


def test_fused_allreduce_alternating_shapes(cuda, mmx_lib):
    """o_proj and down_proj share one workspace: different shapes on alternating parities."""
    from micromix_b200.parallel_utils import PeerWorkspace
    mmx_lib.mmx_set_option(b"tp_reduce_ctas", 8)
    mmx_lib.mmx_set_option(b"tp_timeout_ms", 4000)
    try:
        tp = 2
        sa = _shards(tp, 600, 512, 1024, seed=11)
        sb = _shards(tp, 300, 1024, 2048, seed=12)
        wa, wb = _expected(sa), _expected(sb)
        works = PeerWorkspace.simulate(tp, 600, 1024)
        streams = [torch.cuda.Stream() for _ in range(tp)]
        torch.cuda.synchronize()
        for call in range(5):
            sh, want = (sa, wa) if call % 2 == 0 else (sb, wb)
            ys = []
            for r in range(tp):
                with torch.cuda.stream(streams[r]):
                    ys.append(works[r].matmul_allreduce(*sh[r]))
            torch.cuda.synchronize()
            for r in range(tp):
                assert works[r].status() == 0
                assert torch.equal(ys[r], want), f"rank {r} call {call}"
        for w in works:
            w.close()
    finally:
        mmx_lib.mmx_set_option(b"tp_reduce_ctas", 0)
        mmx_lib.mmx_set_option(b"tp_timeout_ms", 10000)


@pytest.mark.parametrize("mode", ["push", "switch"])
def test_fused_allreduce_multiprocess(cuda, mode):
    """Real ranks, real NVLink: tools/tp_fused_check.py under torchrun on two GPUs of this box (skipped on a one-GPU box).
    push: bit-exact against the rank-ordered fp32 sum.  switch: the NVSwitch sums (multimem.ld_reduce) -- identical bits on
    every rank, within one bf16 rounding step of the rank-ordered sum (the tool exits non-zero otherwise)."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, MMX_TP_MODE=mode)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(29650 + (mode == "switch")), os.path.join(root, "tools", "tp_fused_check.py"),
           "--tokens", "1000", "--iters", "3", "--shapes", "512x1024,1280x4096"]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert f'"mode": "{mode}"' in r.stdout and '"bit_exact_all_ranks": true' in r.stdout


# ---------------------------------------------------------------------------------------------- round 2 additions
@pytest.mark.parametrize("tp,M,N,K", [(2, 300, 512, 1024), (2, 1024, 1024, 1024), (4, 1000, 1152, 2048), (8, 2100, 768, 4096),
                                      (2, 128, 256, 512)])
def test_fused_reduce_scatter_matches_rows_of_the_sum(cuda, mmx_lib, tp, M, N, K):
    """mmx_matmul_reduce_scatter (push data path, simulated ranks): rank r must hold exactly the rows
    [r * shard, (r+1) * shard) of bf16(sum in fp32, rank order) of the partials -- the same bits the all-reduce gives."""
    from micromix_b200.parallel_utils import PeerWorkspace
    mmx_lib.mmx_set_option(b"tp_reduce_ctas", 8)
    mmx_lib.mmx_set_option(b"tp_timeout_ms", 4000)
    try:
        shards = _shards(tp, M, N, K, seed=tp + 40)
        want = _expected(shards)
        works = PeerWorkspace.simulate(tp, M, N)
        streams = [torch.cuda.Stream() for _ in range(tp)]
        torch.cuda.synchronize()
        per = works[0].shard_rows(M)
        assert per % 256 == 0 and per * tp >= M
        for call in range(3):
            ys = []
            for r in range(tp):
                with torch.cuda.stream(streams[r]):
                    ys.append(works[r].matmul_reduce_scatter(*shards[r]))
            torch.cuda.synchronize()
            seen = 0
            for r in range(tp):
                y, row0 = ys[r]
                assert works[r].status() == 0, f"rank {r}: cross-rank wait timed out (call {call})"
                lo, hi = min(M, per * r), min(M, per * (r + 1))
                assert row0 == lo and y.shape == (hi - lo, N)
                assert torch.equal(y, want[lo:hi]), f"rank {r} call {call}: shard differs from the rows of the sum"
                seen += hi - lo
            assert seen == M
        for w in works:
            w.close()
    finally:
        mmx_lib.mmx_set_option(b"tp_reduce_ctas", 0)
        mmx_lib.mmx_set_option(b"tp_timeout_ms", 10000)


def test_fused_allreduce_back_to_back_small(cuda, mmx_lib):
    """many decode-sized fused calls back to back with NOTHING between them (the grids of several calls fit
    on the GPU at once), odd call counts, all-reduce and reduce-scatter interleaved: the reducer releases its dependents
    only at its end and `done` is a monotonic counter, so no call can see another call's counters."""
    from micromix_b200.parallel_utils import PeerWorkspace
    mmx_lib.mmx_set_option(b"tp_reduce_ctas", 4)
    mmx_lib.mmx_set_option(b"tp_timeout_ms", 4000)
    try:
        tp = 2
        sa = _shards(tp, 16, 256, 512, seed=31)
        sb = _shards(tp, 96, 512, 512, seed=32)
        wa, wb = _expected(sa), _expected(sb)
        works = PeerWorkspace.simulate(tp, 128, 512)
        streams = [torch.cuda.Stream() for _ in range(tp)]
        torch.cuda.synchronize()
        plan = [("a", 0), ("a", 0), ("b", 1), ("a", 0), ("b", 0), ("b", 1), ("a", 1)] * 3  # (shape, reduce-scatter?)
        outs = [[] for _ in range(tp)]
        for r in range(tp):
            with torch.cuda.stream(streams[r]):
                for which, rs in plan:
                    sh = sa if which == "a" else sb
                    if rs:
                        y, row0 = works[r].matmul_reduce_scatter(*sh[r])
                        outs[r].append((which, rs, y.clone(), row0))
                    else:
                        outs[r].append((which, rs, works[r].matmul_allreduce(*sh[r]).clone(), 0))
        torch.cuda.synchronize()
        for r in range(tp):
            assert works[r].status() == 0
            for i, (which, rs, y, row0) in enumerate(outs[r]):
                want = wa if which == "a" else wb
                ref = want[row0:row0 + y.shape[0]] if rs else want
                assert torch.equal(y, ref), f"rank {r} call {i} ({which}, rs={rs})"
        for w in works:
            w.close()
    finally:
        mmx_lib.mmx_set_option(b"tp_reduce_ctas", 0)
        mmx_lib.mmx_set_option(b"tp_timeout_ms", 10000)


@pytest.mark.parametrize("M,K,norm", [(512, 1024, False), (300, 2048, False), (1024, 4096, True), (256, 5120, True)])
def test_gather_channel_single_rank(cuda, mmx_lib, M, K, norm):
    """The sequence-parallel hand-over at tp = 1 (the workspace's own address stands in for the multicast mapping): the
    multicast-store quantizer must write exactly the bytes of mmx_reorder_quantize_x / mmx_rmsnorm_quantize_x into the gather
    channel, and the gathered GEMM must equal mmx_matmul on them -- three rounds, so the channel's counters are re-used."""
    from micromix_b200 import mixedgemm
    from micromix_b200.parallel_utils import PeerWorkspace
    N = 512
    split = (K // 2, K // 4, K // 4)
    idx = H.make_index(K, seed=5).to(cuda)
    w = H.make_weights(N, K).to(cuda)
    W = mixedgemm.reorder_quantize_w4(w, idx, *split)
    nw = (1.0 + 0.1 * torch.randn(K, device=cuda)).to(torch.bfloat16)
    ws = PeerWorkspace.simulate(1, M, N, gather=(M, K))[0]
    try:
        for rnd in range(3):
            x = H.make_activations(M, K, idx.cpu(), seed=100 + rnd).to(cuda)
            if norm:
                ref = mixedgemm.rmsnorm_quantize_x(x, nw, 1e-5, idx, *split)
            else:
                ref = mixedgemm.reorder_quantize_x(x, idx, *split)
            got = ws.quantize_allgather(x, M, idx, *split, norm=(nw, 1e-5) if norm else None)
            y = ws.matmul_gathered(M, W, *split)
            torch.cuda.synchronize()
            for i in range(3):
                assert torch.equal(got[i], ref[i]), f"codes of segment {i} (round {rnd})"
                m = torch.from_numpy(H.O.sf_valid_mask(M, split[i], ref[3 + i].numel())).to(cuda)
                assert torch.equal(got[3 + i][: ref[3 + i].numel()][m], ref[3 + i][m]), f"scales of segment {i} (round {rnd})"
            want = mixedgemm.matmul(ref[0], W[0], ref[1], W[1], ref[2], W[2], ref[3], W[3], ref[4], W[4], ref[5], W[5])
            assert torch.equal(y, want), f"gathered GEMM differs (round {rnd})"
    finally:
        ws.close()


def test_tp_layers_multiprocess(cuda):
    """Real ranks: a tensor-parallel decoder layer (NCCL, fused all-reduce, sequence parallel) and an expert-parallel
    Mixtral block against the 1-GPU layer -- tools/tp_layer_check.py under torchrun on two GPUs (skipped on a one-GPU box;
    the builder's multi-GPU logs are under profiles/)."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29671", os.path.join(root, "tools", "tp_layer_check.py"), "--tokens", "512"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert '"ok": true' in r.stdout


@pytest.mark.parametrize("tp,M,N,K", [(1, 300, 512, 1024), (2, 1024, 512, 2048), (4, 1500, 768, 4096), (8, 2100, 512, 4096),
                                      (2, 200, 256, 1024)])
def test_token_parallel_exchange_equals_single_gpu_linear(cuda, mmx_lib, tp, M, N, K):
    """The all-to-all of packed codes (mmx_tp_quantize_alltoall -> mmx_tp_matmul_exchanged), `tp` simulated ranks on one GPU
    (the exchange uses unicast peer pointers only): rank r's rows must equal, bit for bit, the rows of ONE QLinearLayer over
    the full K with the rank-blocked permutation of token_parallel_plan -- three rounds, so the channel counters are re-used,
    ranks with no rows included (M = 200 at tp = 2)."""
    import torch.nn as nn
    from micromix_b200.parallel_utils import PeerWorkspace, TokenParallelQLinear, token_parallel_plan
    from micromix_b200.qLinearLayer import QLinearLayer
    mmx_lib.mmx_set_option(b"tp_timeout_ms", 4000)
    try:
        idx = H.make_index(K, seed=tp)
        p8, p6 = (K // 8) // 128 * 128, (K // 4) // 128 * 128
        lin = nn.Linear(K, N, bias=False).to(torch.bfloat16)
        lin.weight.data = H.make_weights(N, K, seed=7)
        perm, tot, shards = token_parallel_plan(idx, p6, p8, tp)
        single = QLinearLayer(lin, tot[2], tot[1], perm)
        works = PeerWorkspace.simulate(tp, M, N, gather=(M, K))
        layers = [TokenParallelQLinear(lin, p8, p6, idx, None, works[r]) for r in range(tp)]
        streams = [torch.cuda.Stream() for _ in range(tp)]
        per = works[0].shard_rows(M)
        for rnd in range(3):
            x = H.make_activations(M, K, idx, seed=300 + rnd).to(cuda)
            want = single(x.unsqueeze(0)).squeeze(0)
            torch.cuda.synchronize()
            outs = []
            for r in range(tp):
                k0, k1 = layers[r].k_range
                with torch.cuda.stream(streams[r]):
                    outs.append(layers[r](x[:, k0:k1].unsqueeze(0)))
            torch.cuda.synchronize()
            for r in range(tp):
                y, row0 = outs[r]
                lo, hi = min(M, per * r), min(M, per * (r + 1))
                assert works[r].status() == 0, f"rank {r}: a wait timed out (round {rnd})"
                assert row0 == lo and y.shape == (hi - lo, N)
                assert torch.equal(y, want[lo:hi]), f"rank {r} round {rnd}"
        for w in works:
            w.close()
    finally:
        mmx_lib.mmx_set_option(b"tp_timeout_ms", 10000)
