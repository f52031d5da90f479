#!/usr/bin/env python
"""Quantize kernel sweep on the B200 box: quick parity against the oracle, then CUDA-event timing over rotating
buffers larger than L2, for every (M, K) and every option setting given.  Optionally times the reference's own
reorder.cu (oracle/_ref/libref_reorder.so) on the same inputs.

  python tools/quant_sweep.py [--opts quant_rows=0,quant_rows=2] [--ref] [--shapes 8192x4096,8192x14336]
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


def split_for(K):
    p8 = (K // 8) // 128 * 128
    p6 = (K // 4) // 128 * 128
    return K - p6 - p8, p6, p8


def parity(dev):
    ok_all = True
    cases = [(1, 4096), (7, 1024), (127, 4096), (129, 2048), (300, 1024), (515, 4096), (256, 14336), (130, 128),
             (200, 27648), (160, 8192), (96, 6144), (1000, 5120), (64, 11008 // 128 * 128)]
    for M, K in cases:
        sp = split_for(K)
        idx = H.make_index(K, seed=K + M)
        x = H.make_activations(M, K, idx)
        for mode, fn in (("x", mixedgemm.reorder_quantize_x), ("w4", mixedgemm.reorder_quantize_w4)):
            got = fn(x.to(dev), idx.to(dev), *sp)
            torch.cuda.synchronize()
            ref = O.reorder_quantize(H.bits(x), idx.numpy(), *sp, mode)
            ok = all(np.array_equal(H.u8(got[i]), ref[i]) for i in range(3))
            for i, k in enumerate(sp):
                g = H.u8(got[3 + i])
                m = O.sf_valid_mask(M, k, g.shape[0])
                ok &= np.array_equal(g[m], ref[3 + i][m])
            ok_all &= ok
            if not ok:
                for i in range(3):
                    g = H.u8(got[i])
                    bad = np.argwhere(g != ref[i])
                    if bad.shape[0]:
                        print(f"   q{i}: {bad.shape[0]} bytes differ, first {bad[0].tolist()} got {g[tuple(bad[0])]:#x} "
                              f"ref {ref[i][tuple(bad[0])]:#x}")
            print(f"parity M={M} K={K} mode={mode}: {'OK' if ok else 'MISMATCH'}", flush=True)
    return ok_all


def timing(dev, M, K, lib, iters=20, ref=None):
    sp = split_for(K)
    idx = H.make_index(K, seed=0).to(dev)
    nbuf = max(2, int(400e6 // (M * K * 2)) + 1)
    xs = [torch.randn(M, K, device=dev, dtype=torch.float32).to(torch.bfloat16) for _ in range(min(nbuf, 8))]
    u8 = dict(dtype=torch.uint8, device=dev)
    outs = [([torch.empty((M, w), **u8) for w in (sp[0] // 2, sp[1] // 4 * 3, sp[2])],
             [torch.empty((int(lib.mmx_sf_bytes_act(M, k)),), **u8) for k in sp]) for _ in range(len(xs))]
    st = torch.cuda.current_stream().cuda_stream
    p = lambda t: t.data_ptr() if t.numel() else None

    def run(i):
        x = xs[i % len(xs)]
        q, sf = outs[i % len(xs)]
        if ref is None:
            rc = lib.mmx_reorder_quantize_x(p(x), M, K, p(idx), *sp, p(q[0]), p(q[1]), p(q[2]), p(sf[0]), p(sf[1]),
                                            p(sf[2]), st)
        else:
            rc = ref.ref_reorder_quantize(0, p(x), M, p(idx), *sp, p(q[0]), p(q[1]), p(q[2]), p(sf[0]), p(sf[1]), p(sf[2]))
        if rc:
            raise RuntimeError(f"rc={rc} {lib.mmx_last_error().decode()}")

    for i in range(5):
        run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    nbytes = 2.0 * M * K + M * (sp[0] / 2 + sp[1] * 3 / 4 + sp[2]) + M * K / 32
    return us, nbytes / us / 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--opts", default="quant_rows=0")
    ap.add_argument("--shapes", default="2048x4096,8192x4096,16384x4096,8192x14336,16384x14336,8192x5120,8192x8192")
    ap.add_argument("--ref", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    lib = _lib.load()
    shapes = [tuple(int(v) for v in s.split("x")) for s in args.shapes.split(",")]
    for opt in args.opts.split(","):
        kvs = [kv.split("=") for kv in opt.split("+")]
        for k, v in kvs:
            lib.mmx_set_option(k.encode(), int(v))
        print(f"=== options {opt}", flush=True)
        if not args.no_parity:
            print("PARITY_ALL_OK" if parity(dev) else "PARITY_FAILED", flush=True)
        for M, K in shapes:
            us, gbs = timing(dev, M, K, lib)
            print(f"time {opt} M={M} K={K}: {us:8.1f} us  {gbs:7.1f} GB/s", flush=True)
        for k, v in kvs:
            lib.mmx_set_option(k.encode(), 0)
    refso = os.path.join(ROOT, "oracle", "_ref", "libref_reorder.so")
    if args.ref and os.path.exists(refso):
        R = ctypes.CDLL(refso)
        R.ref_reorder_quantize.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p] + \
            [ctypes.c_int] * 3 + [ctypes.c_void_p] * 6
        for M, K in shapes:
            if K not in (3072, 3584, 4096, 5120, 8192, 11008, 12288, 13824, 14336, 18944):
                continue
            us, gbs = timing(dev, M, K, lib, ref=R)
            print(f"time reference reorder.cu M={M} K={K}: {us:8.1f} us  {gbs:7.1f} GB/s", flush=True)


if __name__ == "__main__":
    main()
