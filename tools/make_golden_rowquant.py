#!/usr/bin/env python
"""Generate tests/golden/ref_rowquant_golden.npz from the REFERENCE's own activate / downproj quantize kernels
(/root/reference/mgemm/src/activate.cu) run on a B200.

Needs a GPU and oracle/_ref/libref_activate.so (`make -C oracle ref_activate`, only where /root/reference exists; the
built .so travels to the GPU box with the gpurun snapshot):

    make -C oracle ref_activate
    gpurun -- 'python tools/make_golden_rowquant.py gpurun_out/ref_rowquant_golden.npz'
    cp gpurun_out/ref_rowquant_golden.npz tests/golden/

Inputs are regenerated from seeds by tests/helpers.py (ROWQ_GOLDEN), so only outputs are stored.
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
    R = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_activate.so"))
    R.ref_rowwise_quantize.argtypes = ([ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int] + [ctypes.c_int] * 3 +
                                       [ctypes.c_void_p] * 6)
    dev = torch.device("cuda:0")
    golden = {}
    for tag, (mode, M, split) in H.ROWQ_GOLDEN.items():
        ins = [t.to(dev) for t in H.rowq_golden_inputs(tag)]
        fm = (4, 4, 4) if mode == 2 else (4, 6, 8)
        q = [torch.zeros((M, k * f // 8), dtype=torch.uint8, device=dev) for k, f in zip(split, fm)]
        sf = [torch.zeros((H.O.sf_bytes(M, k, True),), dtype=torch.uint8, device=dev) for k in split]
        torch.cuda.synchronize()
        rc = R.ref_rowwise_quantize(mode, ins[0].data_ptr(), ins[1].data_ptr() if len(ins) > 1 else None, M, *split,
                                    *[t.data_ptr() for t in q], *[t.data_ptr() for t in sf])
        torch.cuda.synchronize()
        assert rc == 0, rc
        for i in range(3):
            golden[f"{tag}_q{i}"] = H.u8(q[i])
            golden[f"{tag}_sf{i}"] = H.u8(sf[i])
    np.savez_compressed(out_path, **golden)
    print("wrote", out_path)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "ref_rowquant_golden.npz"))
