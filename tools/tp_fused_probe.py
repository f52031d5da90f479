#!/usr/bin/env python
"""Timing decomposition of the fused GEMM -> all-reduce (under torchrun): the same call with parts of the exchange
switched off by debug options (results are wrong in those modes; timing only).  One JSON line per shape on rank 0."""
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
from micromix_b200.parallel_utils import PeerWorkspace, init_tensor_parallel, row_shard_plan  # noqa: E402


def timed(fn, iters, dev):
    for _ in range(4):
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
    return round(float(t.item()) * 1e3, 1)


def main():
    rank, world, dev = init_tensor_parallel("nccl")
    lib = mixedgemm._lib.load()
    lib.mmx_set_option(b"tp_timeout_ms", 3000)
    M = int(os.environ.get("TOKENS", "8192"))
    ws = PeerWorkspace(M, 5120, device=dev)
    for N, K in ((4096, 4096), (4096, 14336)):
        idx = H.make_index(K, seed=3)
        p8, p6 = (K // 8) // 128 * 128, (K // 4) // 128 * 128
        k0, k1, lidx, p4, p6, p8 = row_shard_plan(idx, p6, p8, world, rank)
        g = torch.Generator(device=dev).manual_seed(5)
        w = (torch.randn(N, k1 - k0, generator=g, device=dev) * 0.02).to(torch.bfloat16)
        lidx = lidx.to(dev)
        W = mixedgemm.reorder_quantize_w4(w, lidx, p4, p6, p8)
        x = torch.randn(M, k1 - k0, generator=g, device=dev).to(torch.bfloat16)
        A = mixedgemm.reorder_quantize_x(x, lidx, p4, p6, p8)
        out = torch.empty((M, N), dtype=torch.bfloat16, device=dev)
        res = {"tp": world, "M": M, "N": N, "K_local": k1 - k0, "mode": ws.mode}
        res["gemm_local_us"] = timed(lambda: mixedgemm.matmul(A[0], W[0], A[1], W[1], A[2], W[2], A[3], W[3], A[4], W[4],
                                                               A[5], W[5], out=out), 20, dev)
        res["nccl_allreduce_us"] = timed(lambda: dist.all_reduce(out), 20, dev)
        for name, tpd, gd in (("fused", 0, 0), ("no_peer_result_stores", 1, 0), ("no_reduce_work", 2, 0),
                              ("local_push", 0, 8), ("local_push_no_peer_stores", 1, 8), ("relaxed_arrivals", 0, 16),
                              ("no_arrivals_no_waits_no_work", 6, 32), ("no_arrivals_no_waits_no_work_local", 6, 40)):
            lib.mmx_set_option(b"tp_debug", tpd)
            lib.mmx_set_option(b"gemm_debug_flags", gd)
            res[name + "_us"] = timed(lambda: ws.matmul_allreduce(A, W), 20, dev)
        lib.mmx_set_option(b"tp_debug", 0)
        lib.mmx_set_option(b"gemm_debug_flags", 0)
        # timeline of ONE isolated call (ns after the GEMM's first epilogue wait began, this rank's clocks)
        import ctypes
        for rep in range(2):
            dist.barrier()
            torch.cuda.synchronize()
            ws.matmul_allreduce(A, W)
            torch.cuda.synchronize()
        gd = (ctypes.c_uint32 * 48)()
        lib.mmx_gemm_debug_status(gd, 48)
        tt = (ctypes.c_uint64 * 8)()
        lib.mmx_tp_debug_times(tt, 8)
        g0 = gd[40] | (gd[41] << 32)
        g1 = gd[42] | (gd[43] << 32)
        res["timeline_us"] = {"gemm_epilogue_end": round((g1 - g0) / 1e3, 1),
                              "reducer_entry": round((tt[0] - g0) / 1e3, 1), "first_tile_ready": round((tt[1] - g0) / 1e3, 1),
                              "last_unit_done": round((tt[2] - g0) / 1e3, 1), "all_ranks_done": round((tt[3] - g0) / 1e3, 1)}
        res["status"] = ws.status()
        if rank == 0:
            print(json.dumps(res), flush=True)
    ws.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
