"""QQwen2DecoderLayer -- drop-in for /root/reference/model/qQwenLayer.py:87-393 (same constructor and forward).

Identical to the Llama wrapper except that q/k/v carry a bias, which QLinearLayer fuses into the GEMM epilogue with the
rounding of the reference's separate add (model/qLinearLayer.py:70-71).
"""
from __future__ import annotations

from ._qdecoder import QAttention as QQwen2Attention  # noqa: F401
from ._qdecoder import QDecoderLayer
from ._qdecoder import QGatedMLP as QQwen2MLP  # noqa: F401


class QQwen2DecoderLayer(QDecoderLayer):
    def __init__(self, originalLayer, kv_cache, p8_nums, p6_nums, reorder_index, layer_idx, tp_group=None, fused=False,
                 workspace=None, sequence_parallel=False, token_parallel_rows=False):
        super().__init__(originalLayer, kv_cache, p8_nums, p6_nums, reorder_index, layer_idx, tp_group, fused, workspace,
                         sequence_parallel, token_parallel_rows)


QQwenDecoderLayer = QQwen2DecoderLayer
