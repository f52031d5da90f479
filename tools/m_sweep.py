#!/usr/bin/env python
"""BASELINE config 2: Llama-3-8B linear shapes, M = 1 .. 8192 on one B200 -- quantize GB/s, GEMM TFLOP/s and the
quantize + GEMM time per linear.  Kernels are replayed from CUDA graphs (8 launches per replay over rotating buffers) so
that small-M points show device time, not Python launch overhead.  One line per point; the bound that applies is named:
`tensor` (split-weighted MX peak) for large M, `weights/hbm` (packed weight bytes over HBM bandwidth) for M <= 64."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import helpers as H  # noqa: E402
from micromix_b200 import mixedgemm  # noqa: E402

SHAPES = [("qkv", 6144, 4096), ("o", 4096, 4096), ("gate_up", 28672, 4096), ("down", 4096, 14336)]
MS = [int(v) for v in os.environ.get("MS", "1,16,64,128,256,512,1024,2048,4096,8192").split(",")]
ROT = 4


def split_for(K):
    p8, p6 = (K // 8) // 128 * 128, (K // 4) // 128 * 128
    return K - p6 - p8, p6, p8


def graph_time(fn, reps=20):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.graph(g, stream=s):
        fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3  # us per fn()


def main():
    dev = torch.device("cuda:0")
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) \
        else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
    hbm, pb = float(peaks["hbm_gbs"]), float(peaks["bf16_tflops"])
    for name, N, K in SHAPES:
        p4, p6, p8 = split_for(K)
        idx = H.make_index(K, seed=1).to(dev)
        w = (torch.randn(N, K, device=dev) * 0.02).to(torch.bfloat16)
        W = mixedgemm.reorder_quantize_w4(w, idx, p4, p6, p8)
        del w
        wbytes = N * K / 2 + N * K / 32
        for M in MS:
            xs = [torch.randn(M, K, device=dev).to(torch.bfloat16) for _ in range(ROT)]
            As = [mixedgemm.reorder_quantize_x(x, idx, p4, p6, p8) for x in xs]
            outs = [torch.empty((M, N), dtype=torch.bfloat16, device=dev) for _ in range(ROT)]

            def q():
                for i in range(8):
                    mixedgemm.reorder_quantize_x(xs[i % ROT], idx, p4, p6, p8)

            def g():
                for i in range(8):
                    A = As[i % ROT]
                    mixedgemm.matmul(A[0], W[0], A[1], W[1], A[2], W[2], A[3], W[3], A[4], W[4], A[5], W[5], out=outs[i % ROT])

            def both():
                for i in range(8):
                    A = mixedgemm.reorder_quantize_x(xs[i % ROT], idx, p4, p6, p8)
                    mixedgemm.matmul(A[0], W[0], A[1], W[1], A[2], W[2], A[3], W[3], A[4], W[4], A[5], W[5], out=outs[i % ROT])

            tq, tg, tb = graph_time(q) / 8, graph_time(g) / 8, graph_time(both) / 8
            qbytes = 2.0 * M * K + M * (p4 / 2 + p6 * 3 / 4 + p8) + M * K / 32
            flops = 2.0 * M * N * K
            t_tensor = 2.0 * M * N * (p4 / (4 * pb) + (p6 + p8) / (2 * pb)) / 1e6  # us
            t_weights = wbytes / hbm / 1e3  # us
            bound = "tensor" if t_tensor >= t_weights else "weights/hbm"
            print(json.dumps({"linear": name, "N": N, "K": K, "M": M, "quant_us": round(tq, 2), "quant_gbs": round(qbytes / tq / 1e3, 1),
                              "gemm_us": round(tg, 2), "gemm_tflops": round(flops / tg / 1e6, 1), "quant_plus_gemm_us": round(tb, 2),
                              "bound": bound, "gemm_frac_of_bound": round(max(t_tensor, t_weights) / tg, 3)}), flush=True)
            del xs, As, outs
        del W
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
