#!/usr/bin/env python
"""Device-time breakdown of the Mixtral MoE block on ONE GPU: grouped path vs the per-expert loop (torch.profiler)."""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from micromix_b200 import model_shapes as S  # noqa: E402
from micromix_b200.qMixtralLayer import QMixtralSparseMoeBlock  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
cfg = S.MIXTRAL_8X7B
tokens = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
layer = S.make_layer(cfg, dev, seed=0, moe=True)
idx, p6, p8 = S.make_calibration(cfg, 0, moe=True)
x0 = torch.randn(1, tokens, cfg["hidden_size"], device=dev).to(torch.bfloat16)
ep = int(sys.argv[2]) if len(sys.argv) > 2 else 1  # > 1: this process plays rank 0 of an ep-way expert-parallel block
for name, kw in (("grouped", {}), ("loop", {"grouped": False})):
    if ep > 1:
        kw = dict(kw, _emulate_ep=(ep, 0))
    blk = QMixtralSparseMoeBlock(layer.block_sparse_moe, p8, p6, idx, 0, fused=True, **kw)
    for _ in range(2):
        blk(x0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        blk(x0)
    e1.record()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        blk(x0)
        torch.cuda.synchronize()
    agg, cnt = collections.Counter(), collections.Counter()
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            agg[e.name[:80]] += e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total
            cnt[e.name[:80]] += 1
    tot = sum(agg.values())
    print(f"== {name}: {e0.elapsed_time(e1) / 3:.3f} ms per block (events), device kernel time {tot / 1e3:.3f} ms, {sum(cnt.values())} kernels")
    for n, t in agg.most_common(14):
        print(f"{t:9.1f} us {100 * t / tot:5.1f}%  x{cnt[n]:<3d} {n}")
    del blk
