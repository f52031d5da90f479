#!/usr/bin/env python
"""Mixed GEMM sweep on the B200 box.  Stage 1 (always first): small parity cases with the watchdog build, so a wrong
barrier cannot hang the GPU.  Stage 2: parity without the watchdog.  Stage 3: CUDA-event timing of the bench shapes
for every option setting given.

  python tools/gemm_sweep.py [--opts gemm_sf_ahead=1,gemm_sf_ahead=0] [--shapes 8192x6144x4096,...] [--stage N]
"""
import argparse
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import helpers as H  # noqa: E402
from micromix_b200 import _lib, mixedgemm  # noqa: E402

O = H.O
L = _lib.load()
dev = torch.device("cuda:0")


def split_for(K):
    p8 = (K // 8) // 128 * 128
    p6 = (K // 4) // 128 * 128
    return K - p6 - p8, p6, p8


def quant(M, N, K, split, seed=0, sym=False):
    idx = H.make_index(K, seed=seed)
    x, w = H.make_activations(M, K, idx, seed=721 + seed), H.make_weights(N, K, seed=1234 + seed)
    a = mixedgemm.reorder_quantize_x(x.to(dev), idx.to(dev), *split)
    b = (mixedgemm.reorder_quantize_w if sym else mixedgemm.reorder_quantize_w4)(w.to(dev), idx.to(dev), *split)
    return a, b


def mm(a, b, **kw):
    return mixedgemm.matmul(a[0], b[0], a[1], b[1], a[2], b[2], a[3], b[3], a[4], b[4], a[5], b[5], **kw)


def parity(watchdog):
    L.mmx_set_option(b"gemm_watchdog", 1 if watchdog else 0)
    ok_all = True
    cases = [(128, 256, (256, 0, 0), False), (128, 256, (0, 128, 0), False), (128, 256, (0, 0, 128), True),
             (128, 256, (128, 128, 128), False), (300, 512, (384, 128, 128), False), (1, 128, (128, 128, 128), False),
             (129, 384, (640, 256, 128), True), (1000, 1152, (256, 128, 128), False),
             (600, 1024, (2560, 1024, 512), False), (2048, 4096, (2560, 1024, 512), False),
             (515, 2048, (1792, 0, 0), False), (777, 768, (0, 512, 0), False)]
    for M, N, split, sym in cases:
        K = sum(split)
        a, b = quant(M, N, K, split, seed=M + N, sym=sym)
        c = mm(a, b)
        torch.cuda.synchronize()
        st = [0, 0, 0]
        if watchdog:
            buf = (ctypes.c_uint32 * 8)()
            L.mmx_gemm_debug_status(buf, 8)
            st = [int(v) for v in buf[:3]]
        an, bn = [H.u8(t) for t in a], [H.u8(t) for t in b]
        ref = O.matmul(an[0], bn[0], an[1], bn[1], an[2], bn[2], an[3], bn[3], an[4], bn[4], an[5], bn[5], chain=False)
        mx, mean = H.rel_err(H.bits(c), ref)
        exact = float((H.bits(c) == ref).mean())
        ok = mx <= 1e-2 and mean <= 1e-3 and not any(st)
        ok_all &= ok
        print(f"parity wd={int(watchdog)} M={M} N={N} split={split} sym={sym}: max={mx:.2e} mean={mean:.2e} "
              f"exact={exact:.4f} status={[hex(v) for v in st]} {'OK' if ok else 'FAIL'}", flush=True)
        if any(st):
            break
    L.mmx_set_option(b"gemm_watchdog", 0)
    return ok_all


def timing(M, N, K, iters=20):
    split = split_for(K)
    a, b = quant(M, N, K, split)
    nrot = 3  # rotate outputs (and keep >L2 traffic per call at the big shapes)
    outs = [torch.empty((M, N), dtype=torch.bfloat16, device=dev) for _ in range(nrot)]
    for i in range(3):
        mm(a, b, out=outs[i % nrot])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        mm(a, b, out=outs[i % nrot])
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    return us, 2.0 * M * N * K / us / 1e6


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--opts", default="gemm_ctas=0")
    ap.add_argument("--shapes", default="8192x6144x4096,8192x4096x4096,8192x28672x4096,8192x4096x14336,2048x4096x4096")
    ap.add_argument("--stage", type=int, default=3)
    args = ap.parse_args()
    for opt in args.opts.split(","):
        kvs = [kv.split("=") for kv in opt.split("+")]
        for k, v in kvs:
            L.mmx_set_option(k.encode(), int(v))
        print(f"=== options {opt}", flush=True)
        if not parity(True):
            print("WATCHDOG_PARITY_FAILED -- not running without the watchdog", flush=True)
            return 1
        if args.stage >= 2:
            if not parity(False):
                print("PARITY_FAILED", flush=True)
                return 1
            print("PARITY_ALL_OK", flush=True)
        if args.stage >= 3:
            for sh in args.shapes.split(","):
                M, N, K = (int(v) for v in sh.split("x"))
                us, tf = timing(M, N, K)
                print(f"time {opt} M={M} N={N} K={K}: {us:8.1f} us  {tf:7.1f} TFLOP/s", flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
