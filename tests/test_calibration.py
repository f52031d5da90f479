"""Calibration (micromix_b200/calibration.py) against the literal restatement of the reference's reorder_indices.py
(oracle.calibrate_reference): identical reorder_index / p8 / p6 on the same activations, the reference's file format,
hooks keyed like the reference's, and rank-local sharding for tensor parallelism.  CPU only."""
import math

import pytest
import torch
import torch.nn as nn

import helpers as H
from micromix_b200 import calibration as C

O = H.O


def _acts(K, calls, seed, outliers=True):
    g = torch.Generator().manual_seed(seed)
    out = []
    for i in range(calls):
        rows = int(torch.randint(3, 40, (1,), generator=g))
        x = torch.randn(1, rows, K, generator=g)
        if outliers:
            gain = 1.0 + 63.0 * (torch.arange(K) / K) ** 6
            x = x * gain[torch.randperm(K, generator=torch.Generator().manual_seed(seed))]
        out.append(x.to(torch.bfloat16))
    return out


@pytest.mark.parametrize("K,calls,lamda,seed", [(256, 1, 1.0, 0), (1024, 5, 1.0, 1), (4096, 3, 0.5, 2), (1024, 4, 2.0, 3),
                                                (512, 2, 8.0, 4)])
def test_matches_reference_algorithm(K, calls, lamda, seed):
    xs = _acts(K, calls, seed)
    st = C.ActStats(lamda)
    for x in xs:
        st.update(x)
    order, p8, p6, bits = st.result()
    r_order, r8, r6, r4 = O.calibrate_reference(xs, lamda)
    assert torch.equal(order, r_order)
    if r4 >= 0:
        assert (p8, p6) == (r8, r6)
    else:  # the reference would hand QLinearLayer a negative FP4 width; we clamp
        assert p8 + p6 <= K and p8 == min(r8, K)
    assert p8 % 128 == 0 and p6 % 128 == 0 and 0 <= p8 + p6 <= K
    assert 4.0 <= bits <= 8.0


def test_hooks_keys_files_and_sharding(tmp_path):
    torch.manual_seed(0)

    class Block(nn.Module):
        def __init__(self):
            super().__init__()
            self.q_proj, self.o_proj, self.act = nn.Linear(256, 512, bias=False), nn.Linear(512, 256, bias=False), nn.GELU()

        def forward(self, x):
            return self.o_proj(self.act(self.q_proj(x)))

    root = nn.ModuleDict({"layers": nn.ModuleList([Block(), Block()])})
    cal = C.Calibrator(root, lamda=1.0)
    seen = {}
    for i in range(3):
        x = torch.randn(2, 7, 256)
        for li, blk in enumerate(root["layers"]):
            seen.setdefault(f"layers.{li}.q_proj.input", []).append(x)
            h = blk.act(blk.q_proj(x))
            seen.setdefault(f"layers.{li}.o_proj.input", []).append(h)
            x = blk(x)
    idx, p8, p6 = cal.finish()
    assert set(idx) == set(seen) == set(p8) == set(p6)
    for key, calls in seen.items():
        r_order, r8, r6, r4 = O.calibrate_reference(calls, 1.0)
        assert torch.equal(idx[key], r_order)
        if r4 >= 0:
            assert (p8[key], p6[key]) == (r8, r6)
    # the reference's file names and dict format
    files = C.save_calibration("tiny", idx, p8, p6, folder=str(tmp_path))
    assert [f.split("/")[-1] for f in files] == ["tiny_reorder_index_wikitext2.pt", "tiny_p8_num_wikitext2.pt",
                                                  "tiny_p6_num_wikitext2.pt"]
    idx2, p82, p62 = C.load_calibration("tiny", folder=str(tmp_path))
    assert p82 == p8 and p62 == p6 and all(torch.equal(idx2[k], idx[k]) for k in idx)
    with pytest.raises(FileNotFoundError):
        C.load_calibration("missing", folder=str(tmp_path))
    # rank-local entries of the K-sharded linear: a permutation of the rank's slice, splits in multiples of 128
    key = "layers.0.o_proj.input"
    for rank in range(2):
        li, l8, l6 = C.shard_calibration(idx, p8, p6, [key], tp=2, rank=rank)
        assert sorted(li[key].tolist()) == list(range(256)) and li[key].dtype == torch.int16
        assert l8[key] % 128 == 0 and l6[key] % 128 == 0 and l8[key] + l6[key] <= 256
        assert torch.equal(li["layers.0.q_proj.input"], idx["layers.0.q_proj.input"])
