#!/usr/bin/env python
"""Timing of the fused quantizers on the B200 box (CUDA events, rotating buffers larger than L2):
activate_quantize_x (SiLU(gate)*up -> MX), downproj_quantize_w4, rmsnorm_quantize_x, next to what they replace:
the unfused torch SiLU*mul / RMSNorm followed by reorder_quantize_x, and (with --ref) the reference's own activate.cu.

  python tools/rowquant_sweep.py [--shapes 8192x14336,16384x14336] [--norm-shapes 8192x4096,16384x4096] [--ref]
"""
import argparse
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from micromix_b200 import mixedgemm  # noqa: E402


def split_for(K, align=128):
    p8 = (K // 8) // align * align
    p6 = (K // 4) // align * align
    return K - p6 - p8, p6, p8


def timeit(fn, nbuf, iters=20):
    for i in range(3):
        fn(i % nbuf)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i % nbuf)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="2048x14336,8192x14336,16384x14336,8192x4096")
    ap.add_argument("--norm-shapes", default="2048x4096,8192x4096,16384x4096,8192x5120,8192x8192")
    ap.add_argument("--ref", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    R = None
    so = os.path.join(ROOT, "oracle", "_ref", "libref_activate.so")
    if args.ref and os.path.exists(so):
        R = ctypes.CDLL(so)
        R.ref_rowwise_quantize.argtypes = ([ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int] +
                                           [ctypes.c_int] * 3 + [ctypes.c_void_p] * 6)
    for M, K in [tuple(int(v) for v in s.split("x")) for s in args.shapes.split(",")]:
        sp = split_for(K, 512)
        nbuf = min(6, max(2, int(600e6 // (M * K * 4)) + 1))
        gs = [torch.randn(M, K, device=dev).to(torch.bfloat16) for _ in range(nbuf)]
        us_ = [torch.randn(M, K, device=dev).to(torch.bfloat16) for _ in range(nbuf)]
        idx = torch.randperm(K, device=dev).to(torch.int16)
        out_b = M * (sp[0] / 2 + sp[1] * 3 / 4 + sp[2]) + M * K / 32
        t_act = timeit(lambda i: mixedgemm.activate_quantize_x(gs[i], us_[i], *sp), nbuf)
        t_w4 = timeit(lambda i: mixedgemm.downproj_quantize_w4(gs[i], *sp), nbuf)
        t_unf = timeit(lambda i: mixedgemm.reorder_quantize_x(F.silu(gs[i]) * us_[i], idx, *sp), nbuf)
        t_rq = timeit(lambda i: mixedgemm.reorder_quantize_x(gs[i], idx, *sp), nbuf)
        print(f"M={M} K={K} split={sp}", flush=True)
        print(f"  activate_quantize_x        {t_act:8.1f} us  {(4.0 * M * K + out_b) / t_act / 1e3:7.1f} GB/s (algorithmic 4+{out_b / M / K:.3f} B/elem)")
        print(f"  downproj_quantize_w4       {t_w4:8.1f} us  {(2.0 * M * K + M * K / 2 + M * K / 32) / t_w4 / 1e3:7.1f} GB/s")
        print(f"  torch silu*mul + reorder_q {t_unf:8.1f} us   (reorder_quantize_x alone {t_rq:.1f} us)")
        if R is not None:
            u8 = dict(dtype=torch.uint8, device=dev)
            q = [torch.empty((M, w), **u8) for w in (sp[0] // 2, sp[1] // 4 * 3, sp[2])]
            sf = [torch.empty(((M // 128 + 1) * 128 * k // 32,), **u8) for k in sp]
            t_ref = timeit(lambda i: R.ref_rowwise_quantize(0, gs[i].data_ptr(), us_[i].data_ptr(), M, *sp,
                                                            *[t.data_ptr() for t in q], *[t.data_ptr() for t in sf]), nbuf)
            print(f"  reference activate.cu      {t_ref:8.1f} us")
        del gs, us_
    for M, K in [tuple(int(v) for v in s.split("x")) for s in args.norm_shapes.split(",")]:
        sp = split_for(K)
        nbuf = min(8, max(2, int(400e6 // (M * K * 2)) + 1))
        xs = [torch.randn(M, K, device=dev).to(torch.bfloat16) for _ in range(nbuf)]
        w = torch.ones(K, device=dev, dtype=torch.bfloat16)
        idx = torch.randperm(K, device=dev).to(torch.int16)

        def torch_norm(x):
            v = x.float()
            v = v * torch.rsqrt(v.pow(2).mean(-1, keepdim=True) + 1e-5)
            return w * v.to(x.dtype)

        nbytes = 2.0 * M * K + M * (sp[0] / 2 + sp[1] * 3 / 4 + sp[2]) + M * K / 32
        t_f = timeit(lambda i: mixedgemm.rmsnorm_quantize_x(xs[i], w, 1e-5, idx, *sp), nbuf)
        t_q = timeit(lambda i: mixedgemm.reorder_quantize_x(xs[i], idx, *sp), nbuf)
        t_u = timeit(lambda i: mixedgemm.reorder_quantize_x(torch_norm(xs[i]), idx, *sp), nbuf)
        t_n = timeit(lambda i: mixedgemm.reorder_quantize_x(F.rms_norm(xs[i], (K,), w, 1e-5), idx, *sp), nbuf)
        print(f"M={M} K={K} split={sp}", flush=True)
        print(f"  rmsnorm_quantize_x         {t_f:8.1f} us  {nbytes / t_f / 1e3:7.1f} GB/s")
        print(f"  reorder_quantize_x         {t_q:8.1f} us  {nbytes / t_q / 1e3:7.1f} GB/s")
        print(f"  HF-style torch norm + reorder_q {t_u:8.1f} us ;  F.rms_norm + reorder_q {t_n:8.1f} us")
        del xs


if __name__ == "__main__":
    main()
