#!/usr/bin/env python
"""Generate tests/golden/ref_reorder_golden.npz from the REFERENCE's own quantize kernels run on a B200.

Needs a GPU and oracle/_ref/libref_reorder.so (the reference's mgemm/src/reorder.cu compiled in place for sm_100a
by `make -C oracle ref_reorder`, which only works where /root/reference exists; the built .so travels to the GPU
box with the gpurun snapshot).  Typical use from the build container:

    make -C oracle ref_reorder
    gpurun -- 'python tools/make_golden_ref.py gpurun_out/ref_reorder_golden.npz'
    cp gpurun_out/ref_reorder_golden.npz tests/golden/

Inputs are regenerated from seeds by tests/helpers.py (GOLDEN_CASES), so only outputs are stored.
"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402


def main(out_path):
    so = os.path.join(ROOT, "oracle", "_ref", "libref_reorder.so")
    R = ctypes.CDLL(so)
    R.ref_reorder_quantize.argtypes = ([ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p] +
                                       [ctypes.c_int] * 3 + [ctypes.c_void_p] * 6)
    dev = torch.device("cuda:0")
    golden = {}
    for tag, (M, K, (KN, KS, KO)) in H.GOLDEN_CASES.items():
        x, idx = H.golden_inputs(tag)
        xd, idd = x.to(dev), idx.to(dev)
        for mode_i, mode in enumerate(("x", "w", "w4")):
            if tag == "testpy" and mode != "x":
                continue
            fm = (4, 4, 4) if mode == "w4" else (4, 6, 8)
            q = [torch.zeros((M, k * f // 8), dtype=torch.uint8, device=dev) for k, f in zip((KN, KS, KO), fm)]
            sf = [torch.zeros((H.O.sf_bytes(M, k, mode == "x"),), dtype=torch.uint8, device=dev)
                  for k in (KN, KS, KO)]
            torch.cuda.synchronize()
            rc = R.ref_reorder_quantize(mode_i, xd.data_ptr(), M, idd.data_ptr(), KN, KS, KO,
                                        *[t.data_ptr() for t in q], *[t.data_ptr() for t in sf])
            torch.cuda.synchronize()
            assert rc == 0, rc
            for i in range(3):
                golden[f"{tag}_{mode}_q{i}"] = H.u8(q[i])
                golden[f"{tag}_{mode}_sf{i}"] = H.u8(sf[i])
    np.savez_compressed(out_path, **golden)
    print("wrote", out_path)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "ref_reorder_golden.npz"))
