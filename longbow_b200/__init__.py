"""longbow_b200 -- B200-native (sm_100a CUDA) vector-distance / k-NN hot path for Longbow.

Python mirrors of the reference's call surface for this path, over the C ABI in
include/longbow_b200.h:

    longbow_b200.gpu    internal/gpu    (Index / GPUConfig / NewIndexWithConfig)
    longbow_b200.simd   internal/simd   (batch distance functions, ADCDistanceBatch, metric enums)
    longbow_b200.pq     internal/pq     (PQEncoder: BuildADCTable, ADCDistanceBatch, Encode, (de)serialise)
    longbow_b200.store  internal/store  (BruteForceIndex.SearchVectors, RerankBatch, shard merge)
    longbow_b200.shard  row-sharded multi-GPU search + NCCL all-gather merge (torch.distributed)
"""
from . import _lib  # noqa: F401
from ._lib import LongbowError  # noqa: F401

__all__ = ["LongbowError"]
