#!/usr/bin/env python
"""gate_up GEMM + SiLU*up quantize: two kernels (matmul -> activate_quantize_x on the bf16 halves) against the fused
epilogue (matmul_activate_quantize), CUDA events, on Llama-3-8B / Qwen2.5-32B MLP shapes.  Prints one JSON line per case."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from micromix_b200 import mixedgemm  # noqa: E402


def split_for(K):
    p8 = (K // 8) // 128 * 128
    p6 = (K // 4) // 128 * 128
    return K - p6 - p8, p6, p8


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    dev = torch.device("cuda:0")
    for M, K, inter in [(8192, 4096, 14336), (16384, 4096, 14336), (2048, 4096, 14336), (8192, 5120, 27648)]:
        split, dsplit = split_for(K), split_for(inter)
        g = torch.Generator(device=dev).manual_seed(1)
        x = torch.randn(M, K, generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
        wg = (torch.randn(inter, K, generator=g, device=dev, dtype=torch.float32) * 0.02).to(torch.bfloat16)
        wu = (torch.randn(inter, K, generator=g, device=dev, dtype=torch.float32) * 0.02).to(torch.bfloat16)
        idx = torch.randperm(K, generator=g, device=dev).to(torch.int16)
        a = mixedgemm.reorder_quantize_x(x, idx, *split)
        b = mixedgemm.reorder_quantize_w4(torch.cat([wg, wu]), idx, *split)
        bi = mixedgemm.reorder_quantize_w4(mixedgemm.interleave_gate_up(wg, wu), idx, *split)
        del wg, wu
        y = torch.empty((M, 2 * inter), dtype=torch.bfloat16, device=dev)

        def two():
            mixedgemm.matmul(a[0], b[0], a[1], b[1], a[2], b[2], a[3], b[3], a[4], b[4], a[5], b[5], out=y)
            return mixedgemm.activate_quantize_x(y[:, :inter], y[:, inter:], *dsplit)

        def gemm_only():
            mixedgemm.matmul(a[0], b[0], a[1], b[1], a[2], b[2], a[3], b[3], a[4], b[4], a[5], b[5], out=y)

        def fused():
            return mixedgemm.matmul_activate_quantize(a[0], bi[0], a[1], bi[1], a[2], bi[2], a[3], bi[3], a[4], bi[4], a[5],
                                                      bi[5], *dsplit)

        r, f = two(), fused()
        torch.cuda.synchronize()
        nb = -(-M // 128)
        same = all(torch.equal(r[i], f[i]) for i in range(3)) and all(
            torch.equal(r[3 + i][: nb * (k // 128) * 512], f[3 + i][: nb * (k // 128) * 512]) for i, k in enumerate(dsplit))
        t_two, t_gemm, t_fused = timed(two), timed(gemm_only), timed(fused)
        flops = 2.0 * M * 2 * inter * K
        print(json.dumps({"M": M, "K": K, "inter": inter, "bit_identical": same, "gemm_us": round(t_gemm, 1),
                          "gemm_plus_activate_us": round(t_two, 1), "fused_us": round(t_fused, 1),
                          "fused_tflops": round(flops / t_fused / 1e6, 1), "gemm_tflops": round(flops / t_gemm / 1e6, 1)}),
              flush=True)
        del a, b, bi, y, r, f
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
