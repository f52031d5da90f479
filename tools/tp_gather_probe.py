#!/usr/bin/env python
"""Where does the time of the sequence-parallel gather go?  (torchrun, N >= 2 GPUs)

Times, per variant, `iters` rounds of [quantize_allgather -> matmul_gathered] and of the pieces alone:
  plain      mmx_reorder_quantize_x on this rank's rows (local stores): the floor
  gather     the product path (multicast stores + arrival protocol)
  local      the same kernel with its stores redirected to this rank's own buffer (tp_debug 16): protocol without the wire
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

import helpers as H  # noqa: E402
from micromix_b200 import mixedgemm  # noqa: E402
from micromix_b200.parallel_utils import PeerWorkspace, init_tensor_parallel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=int, default=8192)
    ap.add_argument("--K", type=int, default=4096)
    ap.add_argument("--N", type=int, default=768)
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    rank, world, dev = init_tensor_parallel("nccl")
    lib = mixedgemm._lib.load()
    lib.mmx_set_option(b"tp_timeout_ms", 3000)
    M, K, N = args.tokens, args.K, args.N
    split = (K * 5 // 8, K // 4, K // 8)
    idx = H.make_index(K).to(dev)
    ws = PeerWorkspace(M, 4096, device=dev, gather=(M, K))
    lo, hi = ws.shard_range(M)
    x = torch.randn(hi - lo, K, device=dev).to(torch.bfloat16)
    W = mixedgemm.reorder_quantize_w4(H.make_weights(N, K).to(dev), idx, *split)
    out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)

    def timed(fn, n):
        for _ in range(3):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / n * 1e3], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return round(float(t.item()), 1)

    res = {"world": world, "M": M, "K": K, "N": N, "rows_per_rank": hi - lo, "mode": ws.mode}
    res["plain_quantize_us"] = timed(lambda: mixedgemm.reorder_quantize_x(x, idx, *split), args.iters)
    full = torch.randn(M, K, device=dev).to(torch.bfloat16)
    a = mixedgemm.reorder_quantize_x(full, idx, *split)
    res["plain_quantize_all_rows_us"] = timed(lambda: mixedgemm.reorder_quantize_x(full, idx, *split), args.iters)
    res["plain_matmul_us"] = timed(lambda: mixedgemm.matmul(a[0], W[0], a[1], W[1], a[2], W[2], a[3], W[3], a[4], W[4], a[5],
                                                            W[5], out=out), args.iters)

    def pair():
        ws.quantize_allgather(x, M, idx, *split)
        ws.matmul_gathered(M, W, *split, out=out)

    res["gather_plus_gemm_us"] = timed(pair, args.iters)
    lib.mmx_set_option(b"tp_debug", 16)
    res["local_stores_plus_gemm_us"] = timed(pair, args.iters)
    lib.mmx_set_option(b"tp_debug", 0)
    # per-kernel events of the product path
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.iters)]
    dist.barrier()
    torch.cuda.synchronize()
    for e in evs:
        e[0].record()
        ws.quantize_allgather(x, M, idx, *split)
        e[1].record()
        ws.matmul_gathered(M, W, *split, out=out)
        e[2].record()
    torch.cuda.synchronize()
    res["gather_kernel_us_evented"] = round(sum(e[0].elapsed_time(e[1]) for e in evs) / len(evs) * 1e3, 1)
    res["gathered_gemm_us_evented"] = round(sum(e[1].elapsed_time(e[2]) for e in evs) / len(evs) * 1e3, 1)
    res["status"] = ws.status()
    if rank == 0:
        print(json.dumps(res), flush=True)
    ws.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
