#!/usr/bin/env python
"""Kernel durations of the sequence-parallel gather quantizer on ONE GPU (tp = 1 simulated channel): run under
ncu --metrics gpu__time_duration.sum to compare reorder_quantize_kernel<..., MC=false> and <..., MC=true> on the same rows."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import helpers as H  # noqa: E402
from micromix_b200 import mixedgemm  # noqa: E402
from micromix_b200.parallel_utils import PeerWorkspace  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
M, K, N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 4096, 768
split = (2560, 1024, 512)
idx = H.make_index(K).to(dev)
x = torch.randn(M, K, device=dev).to(torch.bfloat16)
W = mixedgemm.reorder_quantize_w4(H.make_weights(N, K).to(dev), idx, *split)
ws = PeerWorkspace.simulate(1, M, N, gather=(M, K))[0]
if os.environ.get("MMX_TP_DEBUG"):
    mixedgemm._lib.load().mmx_set_option(b"tp_debug", int(os.environ["MMX_TP_DEBUG"]))
out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
for _ in range(4):
    mixedgemm.reorder_quantize_x(x, idx, *split)
    ws.quantize_allgather(x, M, idx, *split)
    ws.matmul_gathered(M, W, *split, out=out)
torch.cuda.synchronize()
print("status", ws.status())
