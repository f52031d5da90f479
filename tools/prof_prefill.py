#!/usr/bin/env python
"""Where does a Llama-3-8B prefill layer spend its device time?  torch.profiler over a few fused=True layers
(batch 8 x seq 2048), kernels grouped by name: our hot-path kernels vs attention vs the torch glue around them."""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from micromix_b200 import model_shapes as S  # noqa: E402
from micromix_b200.qLlamaLayer import QLlamaDecoderLayer  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
cfg = S.LLAMA3_8B
n_layers = int(sys.argv[1]) if len(sys.argv) > 1 else 4
layers = []
for i in range(n_layers):
    layer = S.make_layer(cfg, dev, seed=i)
    idx, p6, p8 = S.make_calibration(cfg, i)
    layers.append(QLlamaDecoderLayer(layer, False, p8, p6, idx, i, fused=True))
    del layer
b, s = 8, 2048
x0 = torch.randn(b, s, cfg["hidden_size"], device=dev).to(torch.bfloat16)
pos = S.rope_tables(cfg, b, s, dev)


@torch.no_grad()
def fwd():
    x = x0
    for l in layers:
        x = l(x, position_embeddings=pos)[0]
    return x


for _ in range(2):
    fwd()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    fwd()
    torch.cuda.synchronize()
agg = collections.Counter()
cnt = collections.Counter()
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        name = e.name[:70]
        agg[name] += e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total
        cnt[name] += 1
tot = sum(agg.values())
print(f"layers={n_layers} total device time {tot / 1e3:.2f} ms  ({tot / 1e3 / n_layers:.3f} ms per layer)")
for name, t in agg.most_common(25):
    print(f"{t / n_layers:9.1f} us/layer {100 * t / tot:5.1f}%  x{cnt[name] // n_layers:<3d} {name}")
