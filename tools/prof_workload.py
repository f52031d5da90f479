#!/usr/bin/env python
"""Small fixed workload for ncu: quantize + GEMM of the four Llama-3-8B linears at M tokens, a few iterations."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from micromix_b200 import _lib  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
lib = _lib.load()
dev = torch.device("cuda:0")
lins = [bench.HotLinear(n, N, K, mode, M, 0, 1, dev, lib, seed=i) for i, (n, N, K, mode) in enumerate(bench.LINEARS)]
st = torch.cuda.current_stream().cuda_stream
for _ in range(iters):
    for l in lins:
        l.run(st)
torch.cuda.synchronize()
print("done")
