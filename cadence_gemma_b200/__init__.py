"""B200-native (sm_100a) RG-LRU + Conv1D recurrent hot path of cadence-gemma.

Public surface mirrors ``recurrentgemma.torch`` for this path
(``recurrentgemma/torch/__init__.py:28-31``)."""
from cadence_gemma_b200 import _abi
from cadence_gemma_b200.layers import (BlockDiagonalLinear, Conv1D, RGLRU,
                                       fused_enabled, get_arith_mode, rnn_scan,
                                       set_arith_mode, set_fold_gate, set_fused)

from cadence_gemma_b200.pipeline import GraphedHotPath, recurrent_hot_path, set_fused_conv, set_fused_decode

__all__ = ["recurrent_hot_path", "GraphedHotPath", "BlockDiagonalLinear", "Conv1D", "RGLRU", "rnn_scan",
           "set_arith_mode", "get_arith_mode", "set_fused", "set_fold_gate", "set_fused_conv",
           "set_fused_decode", "fused_enabled", "_abi"]
