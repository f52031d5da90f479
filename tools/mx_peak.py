#!/usr/bin/env python
"""Measured dense MX tensor-pipe peak of this B200 (the roofline denominator of mixed_gemm_kernel).

Every SM issues back-to-back block-scaled tcgen05 MMAs (M=128, N=256) on operands resident in shared memory
(mmx_debug_mma_peak, csrc/gemm.cu): kind::mxf4 (K=64) and kind::mxf8f6f4 (K=32, E3M2 / E4M3 x E2M1).
burst = a 0.3-0.6 ms kernel (best of 5); sustained = 40 back-to-back ~11 ms kernels timed as one interval (~0.45 s: long
enough for the power management to settle -- a single 10 ms kernel still runs at the boost clock).
Prints one JSON line; bench.py calls the same entry point at start-up.
"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def measure(lib, kind, stages, sf_copies, reps):
    t, ms = ctypes.c_double(), ctypes.c_double()
    rc = lib.mmx_debug_mma_peak(kind, stages, sf_copies, reps, ctypes.byref(t), ctypes.byref(ms))
    if rc:
        raise RuntimeError(lib.mmx_last_error().decode())
    return t.value, ms.value


def main():
    import torch
    from micromix_b200 import _lib
    torch.cuda.init()
    torch.cuda.set_device(0)
    lib = _lib.load()
    names = {0: "mxf4_e2m1xe2m1", 1: "mxf8f6f4_e3m2xe2m1", 2: "mxf8f6f4_e4m3xe2m1"}
    out = {}
    for kind, name in names.items():
        burst, ms_b = measure(lib, kind, 2000, 0, 5)
        sust, ms_s = measure(lib, kind, 40000, 0, -40)
        with_sf, _ = measure(lib, kind, 2000, 1, 5)
        out[name] = {"burst_tflops": round(burst, 1), "burst_ms": round(ms_b, 3), "sustained_tflops": round(sust, 1),
                     "sustained_ms": round(ms_s, 3), "burst_with_sf_copies_tflops": round(with_sf, 1)}
    print(json.dumps({"mx_peak": out, "how": "mmx_debug_mma_peak: 148 CTAs x back-to-back tcgen05.mma 128x256xK, smem-resident "
                      "operands, cudaEvent timing"}))


if __name__ == "__main__":
    main()
