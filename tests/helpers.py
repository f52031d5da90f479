"""Shared test helpers: synthetic workloads and torch <-> numpy bf16 plumbing."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402  (tests are allowed to use the oracle)


def bits(t: torch.Tensor) -> np.ndarray:
    """bf16 tensor (any device) -> uint16 numpy bit patterns."""
    return t.detach().cpu().contiguous().view(torch.int16).numpy().view(np.uint16)


def from_bits(b: np.ndarray, device="cpu") -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(b).view(np.int16)).view(torch.bfloat16).to(device)


def u8(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().contiguous().numpy()


def make_index(K: int, seed: int = 0, identity: bool = False) -> torch.Tensor:
    if identity:
        return torch.arange(K, dtype=torch.int16)
    g = torch.Generator().manual_seed(seed)
    return torch.randperm(K, generator=g).to(torch.int16)


def make_activations(M: int, K: int, idx: torch.Tensor, seed: int = 721) -> torch.Tensor:
    """bf16 N(0,1) with a post-permutation channel gain 1 + 31*(c/K)^8 so the FP8 segment carries outliers."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(M, K, generator=g, dtype=torch.float32)
    gain = 1.0 + 31.0 * (torch.arange(K, dtype=torch.float32) / K) ** 8
    xg = torch.empty_like(x)
    xg[:, idx.long()] = x * gain  # permuted channel j reads original channel idx[j]
    return xg.to(torch.bfloat16)


def make_weights(N: int, K: int, seed: int = 1234) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(N, K, generator=g, dtype=torch.float32) * 0.02).to(torch.bfloat16)


def make_testpy_activations(M: int, K: int, KN: int, KS: int, KO: int, seed: int = 721) -> torch.Tensor:
    """The input recipe of the reference's smoke script, /root/reference/mgemm/test.py:11-21 (CPU generator)."""
    g = torch.Generator().manual_seed(seed)
    signs = torch.randint(0, 2, (M, K), generator=g).to(torch.float32) * 2 - 1
    X = torch.rand(M, K, generator=g) * 3
    if KN:
        X[:, -KN:] = torch.rand(M, KN, generator=g) * 8 + 8
    if KS:
        X[:, -KS:] = torch.rand(M, KS, generator=g) * 16 + 16
    if KO:
        X[:, -KO:] = torch.rand(M, KO, generator=g) * 32 + 32
    return (X * signs).to(torch.bfloat16)


SPLITS = {4096: (2560, 1024, 512), 14336: (8960, 3584, 1792), 5120: (3200, 1280, 640), 27648: (17280, 6912, 3456)}


def rel_err(got_bits: np.ndarray, ref_bits: np.ndarray):
    """(max, mean) of |got-ref| / max(|ref|, rms(ref)) over bf16 outputs -- the north_star tolerance metric
    (<= 1e-2 max, <= 1e-3 mean).  The rms floor keeps near-zero outputs from dominating."""
    g = O.bf16_bits_to_f32(got_bits).astype(np.float64)
    r = O.bf16_bits_to_f32(ref_bits).astype(np.float64)
    rms = np.sqrt(np.mean(r * r)) + 1e-30
    d = np.abs(g - r) / np.maximum(np.abs(r), rms)
    return float(d.max()), float(d.mean())


# ---- cases whose reference-kernel outputs are committed in tests/golden/ref_reorder_golden.npz
# (generated on a B200 by tools/make_golden_ref.py from the reference's own reorder.cu)
GOLDEN_CASES = {
    # /root/reference/mgemm/test.py:7-25 recipe: M=128, K=11008, identity index, boosted tail channels
    "testpy": (128, 11008, (11008 - 1024, 1024 - 128, 128)),
    "mixed": (300, 4096, (2560, 1024, 512)),
    "thirds": (64, 3072, (1024, 1024, 1024)),
}


def golden_inputs(tag):
    M, K, (KN, KS, KO) = GOLDEN_CASES[tag]
    idx = make_index(K, seed=1, identity=(tag == "testpy"))
    x = make_testpy_activations(M, K, KN, KS, KO) if tag == "testpy" else make_activations(M, K, idx)
    return x, idx


def load_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_reorder_golden.npz"))


# ---- ops without a permutation (activate.cu): inputs and the cases whose reference-kernel outputs are committed in
# tests/golden/ref_rowquant_golden.npz (generated on a B200 by tools/make_golden_rowquant.py)
def make_gate_up(M: int, K: int, seed: int):
    """gate ~ N(0, 1.5), up ~ N(0, 1) with a channel gain so that the FP8 segment carries outliers (bf16)."""
    g = torch.Generator().manual_seed(seed)
    gate = (torch.randn(M, K, generator=g) * 1.5).to(torch.bfloat16)
    gain = 1.0 + 15.0 * (torch.arange(K, dtype=torch.float32) / K) ** 8
    up = (torch.randn(M, K, generator=g) * gain).to(torch.bfloat16)
    return gate, up


# tag -> (mode, rows, split); mode 0 activate_quantize_x, 1 downproj_quantize_w, 2 downproj_quantize_w4.
# KN and KN+KS are multiples of 512: the reference kernel's __syncthreads (activate.cu:187) is divergent otherwise.
ROWQ_GOLDEN = {
    "act_4096": (0, 300, (2560, 1024, 512)), "act_14336": (0, 130, (8704, 3584, 2048)),
    "w_4096": (1, 257, (2560, 1024, 512)), "w4_4096": (2, 257, (2560, 1024, 512)),
}


def rowq_golden_inputs(tag):
    mode, M, split = ROWQ_GOLDEN[tag]
    K = sum(split)
    if mode == 0:
        return make_gate_up(M, K, seed=1000 + K)
    return (make_weights(M, K, seed=2000 + K),)
