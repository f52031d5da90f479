"""Calibration: activation statistics -> (reorder_index, p8_num, p6_num) per linear -- replaces
/root/reference/reorder_indices.py:35-151 and the loader in /root/reference/model/main.py:112-124.

Same results as the reference on the same activations, without its memory profile: the reference keeps EVERY |x| row
of every linear on the host (`total_scales[name].append(tensor)`, reorder_indices.py:49-53) and concatenates them at the
end; here the two threshold counts it derives from that concatenation are accumulated batch by batch (they are integer
counts over rows that never interact, so the streaming totals are exactly the reference's).

Per linear input (key `'<module name>.input'`, reorder_indices.py:63):
  * act_scale[c]  = max over forward calls of mean_rows |x[:, c]|                         (:43-51)
  * reorder_index = argsort(act_scale, ascending): outlier channels last                  (:66-70, :99)
  * per row r of all calls: m = max_c |x[r, c]|;
        thr4 = m * 448 / 6 / 2^10 * lamda,  thr6 = m * 448 / 28 / 2^6 * lamda              (:103-104)
    p4_ratio = #(|x| < thr4) / #x,  p6_ratio = #(|x| < thr6) / #x - p4_ratio,  p8_ratio = 1 - p4_ratio - p6_ratio
    p6_num = ceil(K * p6_ratio / 128) * 128,  p8_num likewise,  p4_num = K - p6_num - p8_num  (:106-111)
    (float32 tensor arithmetic, as torch evaluates the reference's expressions).
    The reference does not guard p4_num < 0; here p8_num and then p6_num are clamped so that the three stay in [0, K].

Files: `saved/{model}_reorder_index_wikitext2.pt`, `..._p8_num_...`, `..._p6_num_...` -- dicts keyed as above, loadable by
the reference's main.py and by `load_calibration` here.  `shard_calibration` turns a global entry into the rank-local
permutation / split of a row-parallel (K-sharded) linear.
"""
from __future__ import annotations

import functools
import math
import os
from typing import Dict, Iterable, Optional, Tuple

import torch
import torch.nn as nn


class ActStats:
    """Streaming statistics of ONE linear's input (all arithmetic on the host in fp32, like the reference)."""

    def __init__(self, lamda: float = 1.0):
        self.lamda = float(lamda)
        self.act_scale: Optional[torch.Tensor] = None
        self.n4 = 0      # elements below their row's FP4 threshold
        self.n6 = 0      # ... below the FP6 threshold
        self.numel = 0
        self.in_features = 0

    @torch.no_grad()
    def update(self, x: torch.Tensor) -> None:
        t = x.reshape(-1, x.shape[-1]).float().detach().cpu().abs()
        scales = torch.mean(t, dim=0).float()
        self.act_scale = scales if self.act_scale is None else torch.max(self.act_scale, scales)
        rowmax = t.max(dim=-1, keepdim=True)[0]
        thr4 = rowmax * 448 / 6 / math.pow(2, 10) * self.lamda
        thr6 = rowmax * 448 / 28 / math.pow(2, 6) * self.lamda
        self.n4 += int((t < thr4).sum())
        self.n6 += int((t < thr6).sum())
        self.numel += t.numel()
        self.in_features = t.shape[-1]

    def result(self) -> Tuple[torch.Tensor, int, int, float]:
        """(reorder_index int64 [K], p8_num, p6_num, average bits)."""
        if self.act_scale is None:
            raise ValueError("no activations were recorded")
        K = self.in_features
        _, order = torch.sort(self.act_scale, descending=False)
        # the reference's expressions, evaluated as torch evaluates them (0-dim tensors: int64 sum / int -> float32)
        p4_ratio = torch.tensor(self.n4) / self.numel
        p6_ratio = torch.tensor(self.n6) / self.numel - p4_ratio
        p8_ratio = 1 - p4_ratio - p6_ratio
        p6 = math.ceil(K * p6_ratio / 128) * 128
        p8 = math.ceil(K * p8_ratio / 128) * 128
        p8 = min(max(p8, 0), K)
        p6 = min(max(p6, 0), K - p8)
        avg_bits = float(4 * p4_ratio + 6 * p6_ratio + 8 * p8_ratio)
        return order, int(p8), int(p6), avg_bits


class Calibrator:
    """Forward hooks on every nn.Linear below `root` (the reference hooks `model.model`, reorder_indices.py:73-79).

        cal = Calibrator(model.model, lamda=1.0)
        for batch in batches: model(batch)
        reorder_index, p8_nums, p6_nums = cal.finish()
    """

    def __init__(self, root: nn.Module, lamda: float = 1.0):
        self.lamda = lamda
        self.stats: Dict[str, ActStats] = {}
        self.hooks = []
        for name, m in root.named_modules():
            if isinstance(m, nn.Linear):
                self.hooks.append(m.register_forward_hook(functools.partial(self._hook, name=name)))

    def _hook(self, m, x, y, name):
        if isinstance(x, tuple):
            x = x[0]
        self.observe(name + ".input", x)

    def observe(self, key: str, x: torch.Tensor) -> None:
        st = self.stats.get(key)
        if st is None:
            st = self.stats[key] = ActStats(self.lamda)
        st.update(x)

    def finish(self):
        for h in self.hooks:
            h.remove()
        self.hooks = []
        order, p8, p6, self.average_bits = {}, {}, {}, {}
        for key, st in self.stats.items():
            order[key], p8[key], p6[key], self.average_bits[key] = st.result()
        return order, p8, p6


def calibration_paths(model_name: str, folder: str = "./saved", dataset: str = "wikitext2"):
    """The reference's file names (reorder_indices.py:149-151, main.py:112-114)."""
    base = os.path.join(folder, model_name)
    return (f"{base}_reorder_index_{dataset}.pt", f"{base}_p8_num_{dataset}.pt", f"{base}_p6_num_{dataset}.pt")


def save_calibration(model_name: str, reorder_index, p8_nums, p6_nums, folder: str = "./saved", dataset: str = "wikitext2"):
    os.makedirs(folder, exist_ok=True)
    fi, f8, f6 = calibration_paths(model_name, folder, dataset)
    torch.save(reorder_index, fi)
    torch.save(p8_nums, f8)
    torch.save(p6_nums, f6)
    return fi, f8, f6


def load_calibration(model_name: str, folder: str = "./saved", dataset: str = "wikitext2"):
    """(reorder_index, p8_nums, p6_nums) dicts as main.py:120-123 loads them; raises if the index file is missing."""
    fi, f8, f6 = calibration_paths(model_name, folder, dataset)
    if not os.path.isfile(fi):
        raise FileNotFoundError(f"reorder index file not found: {fi}")
    return (torch.load(fi, weights_only=False), torch.load(f8, weights_only=False), torch.load(f6, weights_only=False))


def shard_calibration(reorder_index, p8_nums, p6_nums, row_parallel_keys: Iterable[str], tp: int, rank: int):
    """Rank-local calibration for tensor parallelism: entries of row-parallel (K-sharded) linears are replaced by
    (local int16 permutation of the rank's channel slice, local p8, local p6); every other entry is shared unchanged."""
    from .parallel_utils import row_shard_plan
    idx, p8, p6 = dict(reorder_index), dict(p8_nums), dict(p6_nums)
    for key in row_parallel_keys:
        _, _, local, _, l6, l8 = row_shard_plan(reorder_index[key], int(p6_nums[key]), int(p8_nums[key]), tp, rank)
        idx[key], p8[key], p6[key] = local, l8, l6
    return idx, p8, p6
