"""recommendations_b200: B200 (sm_100a) embedding hot path behind the reference's
nn.Module API (commons/layers.py, commons/transformers/layers.py).

The compute lives in recommendations_b200/lib/librecemb_b200.so (hand-written CUDA,
C ABI in include/recemb_b200.h).  Importing the package does not need a GPU; calling
any op does, and a missing library raises instead of falling back.
"""
from . import _native
from .collection import EmbeddingCollection
from .interaction import DotInteraction, PooledInteraction, dot_interaction
from .layers import (CosineVectorEmbedding, FlatEmbedding, KShiftEmbedding, PatternFromTimelocal,
                     PooledEmbeddingBag, QREmbedding)
from .logq import CascadedStreamingLogQCorrectionModule, StreamingLogQCorrectionModule
from .sequence import SequenceWindow, fused_lookup_sum, sequence_trim
from .table import EmbeddingTable, FusedEmbeddingOptimizer, FusedOptimizerConfig

__all__ = [
    "CosineVectorEmbedding", "DotInteraction", "EmbeddingCollection", "EmbeddingTable", "dot_interaction", "FlatEmbedding", "FusedEmbeddingOptimizer",
    "FusedOptimizerConfig", "KShiftEmbedding", "PatternFromTimelocal", "PooledEmbeddingBag", "QREmbedding",
    "CascadedStreamingLogQCorrectionModule", "StreamingLogQCorrectionModule", "SequenceWindow", "sequence_trim",
    "PooledInteraction", "fused_lookup_sum",
]
