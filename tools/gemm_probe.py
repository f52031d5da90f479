#!/usr/bin/env python
"""Per-role cycle accounting of the mixed GEMM (watchdog build): where does CTA 0 spend its time?"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import helpers as H  # noqa: E402
from micromix_b200 import mixedgemm, _lib  # noqa: E402

L = _lib.load()
dev = torch.device("cuda:0")


def probe(M, N, K, cg, label="", flags=0):
    KN, KS, KO = H.SPLITS[K]
    idx = H.make_index(K).to(dev)
    x = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
    w = (torch.randn(N, K, device=dev, dtype=torch.float32) * 0.02).to(torch.bfloat16)
    a = mixedgemm.reorder_quantize_x(x, idx, KN, KS, KO)
    b = mixedgemm.reorder_quantize_w4(w, idx, KN, KS, KO)
    out = torch.empty((M, N), dtype=torch.bfloat16, device=dev)
    L.mmx_set_option(b"gemm_cta_group", 1 if cg == 1 else 0)
    L.mmx_set_option(b"gemm_watchdog", 1)
    L.mmx_set_option(b"gemm_debug_flags", flags)
    for _ in range(3):
        mixedgemm.matmul(a[0], b[0], a[1], b[1], a[2], b[2], a[3], b[3], a[4], b[4], a[5], b[5], out=out)
    torch.cuda.synchronize()
    buf = (ctypes.c_uint32 * 32)()
    L.mmx_gemm_debug_status(buf, 32)
    d = [int(v) * 16 for v in buf]
    nst, ntl = buf[21], buf[24]
    print(f"[{label} cg={cg}] M={M} N={N} K={K}: status={[hex(v) for v in buf[:3]]}")
    print(f"   producer: total {d[16]} cyc, waiting-for-slot {d[17]} ({100.0 * d[17] / max(d[16], 1):.0f}%)")
    print(f"   mma     : total {d[18]} cyc, wait-data {d[19]} ({100.0 * d[19] / max(d[18], 1):.0f}%), wait-tmem {d[20]} "
          f"({100.0 * d[20] / max(d[18], 1):.0f}%), stages {nst}, cyc/stage {d[18] / max(nst, 1):.0f}")
    print(f"   epilogue: wait {d[22]} work {d[23]} tiles {ntl}  work/tile {d[23] / max(ntl, 1):.0f}")
    L.mmx_set_option(b"gemm_watchdog", 0)
    L.mmx_set_option(b"gemm_debug_flags", 0)
    L.mmx_set_option(b"gemm_cta_group", 0)


for (M, N, K) in [(8192, 4096, 4096), (8192, 4096, 14336)]:
    for cg in (1, 2):
        for flags, label in ((0, "baseline"), (1, "no SF copies"), (4, "no C stores")):
            probe(M, N, K, cg, label, flags)
