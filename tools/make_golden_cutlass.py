#!/usr/bin/env python
"""Generate tests/golden/cutlass_convert_table.npz from the REFERENCE's own code run on this host.

Runs oracle/_ref/ref_convert (oracle/ref_convert.cpp compiled against the vendored CUTLASS under
/root/reference/cutlass -- `make -C oracle ref_convert`), which evaluates, for every one of the 65536 bf16 bit
patterns, cutlass::NumericConverter<float_e2m1_t|float_e3m2_t|float_e4m3_t|float_ue8m0_t, float, RNE> (the
converters reorder.cu:138-141 instantiates), and the CuTe scale-factor layout offsets the reference kernel
writes through (reorder.cu:182-185).  Only runs where /root/reference exists; the output is committed.
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref_convert"])
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_convert")
    tmp = os.path.join(ROOT, "oracle", "_ref", "ref_convert_table.bin")
    subprocess.check_call([exe, tmp])
    raw = np.fromfile(tmp, dtype=np.uint8)
    t = raw[:4 * 65536].reshape(4, 65536)
    rest = raw[4 * 65536:].view(np.int64)
    sfa = rest[:300 * 32].reshape(300, 32)
    sfb = rest[300 * 32:300 * 32 + 384 * 20].reshape(384, 20)
    assert rest.size == 300 * 32 + 384 * 20
    out = os.path.join(ROOT, "tests", "golden", "cutlass_convert_table.npz")
    np.savez_compressed(out, e2m1=t[0], e3m2=t[1], e4m3=t[2], ue8m0=t[3], sfa_off_M300_K1024=sfa,
                        sfb_off_N384_K640=sfb)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    sys.exit(main())
