#!/usr/bin/env python
"""Multi-process check + timing of the fused row-parallel GEMM -> all-reduce (run under torchrun on N GPUs of one box).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
      tools/tp_fused_check.py [--tokens 8192]

Parity: every rank's partial (mmx_matmul on its K shard, bf16) is all-gathered; the fused op must return
bf16(sum in fp32, rank order) of the partials bit for bit on every rank.  Timing: quantize + GEMM + reduction per
row-parallel linear, fused (mmx_matmul_allreduce) vs plain (mmx_matmul + NCCL all_reduce), CUDA events, max over ranks.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.nn as nn  # noqa: E402

import helpers as H  # noqa: E402
from micromix_b200 import mixedgemm  # noqa: E402
from micromix_b200.parallel_utils import PeerWorkspace, RowParallelQLinear, init_tensor_parallel  # noqa: E402


def timed(fn, iters, warmup, dev):
    for _ in range(warmup):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) * 1e3  # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=int, default=8192)
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--shapes", default="4096x4096,4096x14336,5120x27648")
    args = ap.parse_args()
    rank, world, dev = init_tensor_parallel("nccl")
    mixedgemm._lib.load().mmx_set_option(b"tp_timeout_ms", 5000)
    M = args.tokens
    shapes = [tuple(int(v) for v in s.split("x")) for s in args.shapes.split(",")]
    shapes = [(n, k) for n, k in shapes if (k // world) % 128 == 0]
    ws = PeerWorkspace(M, max(n for n, _ in shapes), device=dev)
    ok_all = True
    for N, K in shapes:
        p8 = (K // 8) // 128 * 128
        p6 = (K // 4) // 128 * 128
        idx = H.make_index(K, seed=3)
        lin = nn.Linear(K, N, bias=False, device="meta", dtype=torch.bfloat16)
        g = torch.Generator(device=dev).manual_seed(1234)
        lin.weight = nn.Parameter((torch.randn(N, K, generator=g, device=dev) * 0.02).to(torch.bfloat16), requires_grad=False)
        fused = RowParallelQLinear(lin, p8, p6, idx, workspace=ws)
        plain = RowParallelQLinear(lin, p8, p6, idx)
        plain.linear = fused.linear  # same quantized shard
        k0, k1 = fused.k_range
        gx = torch.Generator(device=dev).manual_seed(721)
        x = torch.randn(M, K, generator=gx, device=dev).to(torch.bfloat16)[:, k0:k1].contiguous().view(1, M, -1)
        # ---- parity: exact against the rank-ordered fp32 sum of the gathered bf16 partials
        part = fused.linear(x).view(M, N)
        parts = [torch.empty_like(part) for _ in range(world)]
        dist.all_gather(parts, part)
        want = torch.zeros((M, N), dtype=torch.float32, device=dev)
        for p_ in parts:
            want += p_.float()
        want = want.to(torch.bfloat16)
        ok = True
        ulp_off = 0.0
        for call in range(3):
            y = fused(x).view(M, N)
            torch.cuda.synchronize()
            if ws.mode == "switch":
                # the switch sums in its own order (fp32 accumulation): identical bits on every rank, and within one
                # bf16 rounding step of the rank-ordered sum
                y0 = y.clone()
                dist.broadcast(y0, src=0)
                d = (y.float() - want.float()).abs()
                tol = want.float().abs() * 2.0 ** -7 + 1e-30
                ulp_off = max(ulp_off, float((d > 0).float().mean()))
                ok = ok and bool(torch.equal(y, y0)) and bool((d <= tol).all()) and ws.status() == 0
            else:
                ok = ok and bool(torch.equal(y, want)) and ws.status() == 0
        y_nccl = plain(x).view(M, N).float()
        dn = float((y_nccl - want.float()).abs().max() / want.float().abs().max())
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok_all = ok_all and bool(flag.item())
        # ---- timing
        t_fused = timed(lambda: fused(x), args.iters, 5, dev)
        t_plain = timed(lambda: plain(x), args.iters, 5, dev)
        t_local = timed(lambda: fused.linear(x), args.iters, 5, dev)
        if rank == 0:
            print(json.dumps({"tp": world, "mode": ws.mode, "M": M, "N": N, "K": K, "K_local": k1 - k0, "bit_exact_all_ranks": bool(flag.item()),
                              "frac_elems_off_by_one_rounding": ulp_off,
                              "nccl_vs_exact_max_rel": dn, "fused_us": round(t_fused, 1), "gemm_plus_nccl_us": round(t_plain, 1),
                              "local_quant_gemm_us": round(t_local, 1),
                              "exposed_reduce_us_fused": round(t_fused - t_local, 1),
                              "exposed_reduce_us_nccl": round(t_plain - t_local, 1),
                              "allreduce_bytes": M * N * 2}), flush=True)
    st = ws.status()
    ws.close()
    dist.barrier()
    dist.destroy_process_group()
    if not ok_all or st:
        sys.exit(1)


if __name__ == "__main__":
    main()
