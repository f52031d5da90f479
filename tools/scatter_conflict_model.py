#!/usr/bin/env python
"""CPU model of the shared-memory bank conflicts in quantize.cu's scatter-on-write (no GPU needed).

The kernel stores one 4-byte slot (two rows of one channel) per lane and instruction at the slot's PERMUTED position;
with a random reorder_index the 32 destinations of an instruction fall on random banks.  This script counts the
wavefronts per store instruction for (a) the shipped assignment (lane = 16-byte chunk index mod 32) and (b) a greedy
per-layer schedule that picks, for every instruction group, the 16 32-byte sectors whose eight element positions collide
least -- the next step named in DESIGN.md section 10.  ncu measures 50 % of the kernel's shared-memory wavefronts as
replays (profiles/r01_ncu_full_summary_s8.txt); the model's 3.5 wavefronts per instruction for (a) agrees with that.

  python tools/scatter_conflict_model.py [K]
"""
import sys

import numpy as np


def bank(j):
    """Bank of permuted slot j (4-byte slots; 16-byte chunk p = j / 4 stored at p ^ ((p >> 3) & 3), quantize.cu)."""
    return 4 * (((j // 4) % 8) ^ ((j // 32) & 3)) + j % 4


def model(K, seed=0):
    rng = np.random.default_rng(seed)
    idx = rng.permutation(K)
    inv = np.empty(K, int)
    inv[idx] = np.arange(K)
    nchunks = K // 8
    banks = np.stack([bank(inv[8 * np.arange(nchunks) + e]) for e in range(8)], 1)  # [chunk, element]
    shipped = np.mean([np.bincount(banks[32 * g:32 * g + 32, e], minlength=32).max()
                       for g in range(nchunks // 32) for e in range(8)])
    remaining = set(range(nchunks // 2))
    tot, n = 0, 0
    while remaining:
        cnt = np.zeros((8, 32), int)
        for _ in range(16):
            cand = list(remaining)
            if len(cand) > 200:
                cand = list(rng.choice(cand, 200, replace=False))
            cost = [sum(cnt[e, banks[ch, e]] for ch in (2 * s, 2 * s + 1) for e in range(8)) for s in cand]
            best = cand[int(np.argmin(cost))]
            remaining.discard(best)
            for ch in (2 * best, 2 * best + 1):
                for e in range(8):
                    cnt[e, banks[ch, e]] += 1
            if not remaining:
                break
        tot += cnt.max(axis=1).sum()
        n += 8
    return float(shipped), tot / n


if __name__ == "__main__":
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    a, b = model(K)
    print(f"K={K}: wavefronts per scatter instruction  shipped {a:.2f}   greedy sector schedule {b:.2f}   ideal 1.00")
