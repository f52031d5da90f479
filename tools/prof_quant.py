#!/usr/bin/env python
"""One quantize shape, a few launches, for ncu.  usage: prof_quant.py M K [opt=val ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import quant_sweep as Q
from micromix_b200 import _lib
lib = _lib.load()
M, K = int(sys.argv[1]), int(sys.argv[2])
for kv in sys.argv[3:]:
    k, v = kv.split("="); lib.mmx_set_option(k.encode(), int(v))
us, gbs = Q.timing(torch.device("cuda:0"), M, K, lib, iters=6)
print(f"M={M} K={K}: {us:.1f} us {gbs:.1f} GB/s")
