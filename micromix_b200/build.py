"""In-tree build of libmicromix_b200.so (the C-ABI CUDA library) with nvcc for sm_100a.

The .so lands next to this file so that it travels with the repo snapshot to the GPU box and shows up as an
in-tree native library in the loaded-module list.  No torch headers, no pybind: plain `nvcc -shared`.
"""
from __future__ import annotations

import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libmicromix_b200.so")
SOURCES = ["lib.cu", "quantize.cu", "rowquant.cu", "gemm.cu", "tp_reduce.cu", "moe.cu", "rope.cu"]
NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-shared", "-cudart", "shared",
]


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(_ROOT, "include", "micromix_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    tmp = LIB_PATH + f".tmp{os.getpid()}"
    cmd = [nvcc, *NVCC_FLAGS, "-I", os.path.join(_ROOT, "include"), "-I", CSRC,
           *[os.path.join(CSRC, s) for s in SOURCES], "-o", tmp]
    if verbose:
        print(" ".join(cmd))
    try:
        subprocess.check_call(cmd)
        os.replace(tmp, LIB_PATH)  # atomic: a snapshot of the tree never sees a half-written library
    finally:
        if os.path.exists(tmp):
            os.remove(tmp)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
