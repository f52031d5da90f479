"""QLlamaDecoderLayer -- drop-in for /root/reference/model/qLlamaLayer.py:69-158 (same constructor and forward).

All seven linears run through the B200 hot path (reorder+quantize -> mixed GEMM); q/k/v and gate/up are fused into one
quantize + one GEMM each when they share the calibration entry (they always do: same input tensor); RoPE, SDPA and
RMSNorm stay stock PyTorch, as in the reference.  `tp_group` (extension) shards heads / intermediate channels across
ranks: column-parallel qkv / gate_up, row-parallel o / down with an NCCL all-reduce.
"""
from __future__ import annotations

from ._qdecoder import QAttention as QLlamaAttention  # noqa: F401  (qLlamaLayer.py:196)
from ._qdecoder import QDecoderLayer
from ._qdecoder import QGatedMLP as QLlamaMLP  # noqa: F401  (qLlamaLayer.py:324)


class QLlamaDecoderLayer(QDecoderLayer):
    def __init__(self, originalLayer, kv_cache, p8_nums, p6_nums, reorder_index, layer_idx, tp_group=None, fused=False,
                 workspace=None, sequence_parallel=False, token_parallel_rows=False):
        super().__init__(originalLayer, kv_cache, p8_nums, p6_nums, reorder_index, layer_idx, tp_group, fused, workspace,
                         sequence_parallel, token_parallel_rows)
