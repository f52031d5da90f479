#!/usr/bin/env python
"""NVLink peer bandwidth between cuda:0 and cuda:1 of one box (single process, copy engines): the wire-rate reference
for the fused GEMM -> all-reduce numbers in profiles/."""
import json
import torch

n = 256 << 20
a = torch.empty(n, dtype=torch.uint8, device="cuda:0")
b = torch.empty(n, dtype=torch.uint8, device="cuda:1")
c = torch.empty(n, dtype=torch.uint8, device="cuda:1")
d = torch.empty(n, dtype=torch.uint8, device="cuda:0")
out = {"can_access_peer": torch.cuda.can_device_access_peer(0, 1)}


def timed(fn, dev, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(0), torch.cuda.synchronize(1)
    with torch.cuda.device(dev):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
    torch.cuda.synchronize(0), torch.cuda.synchronize(1)
    return e0.elapsed_time(e1) / iters


with torch.cuda.device(0):
    ms = timed(lambda: b.copy_(a, non_blocking=True), 0)
out["push_0_to_1_GBs"] = round(n / ms / 1e6, 1)
s0, s1 = torch.cuda.Stream(0), torch.cuda.Stream(1)


def both():
    with torch.cuda.stream(s0):
        b.copy_(a, non_blocking=True)
    with torch.cuda.stream(s1):
        d.copy_(c, non_blocking=True)


import time
for _ in range(3):
    both()
torch.cuda.synchronize(0), torch.cuda.synchronize(1)
t0 = time.perf_counter()
for _ in range(10):
    both()
torch.cuda.synchronize(0), torch.cuda.synchronize(1)
out["bidir_each_direction_GBs"] = round(n * 10 / (time.perf_counter() - t0) / 1e9, 1)
print(json.dumps(out))
