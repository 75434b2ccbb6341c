"""Mirror of the reference's ``internal/simd`` batch call surface, executed on the GPU.

Function names, argument meaning and error behaviour follow internal/simd/batch_operations.go and
internal/simd/distance_functions.go.  These are the *stateless* entry points (host buffers in and
out; the flat buffer is uploaded per call) -- resident data goes through ``gpu.DenseIndex``.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from ._lib import F16, F32, I8, U8, check

# internal/simd/registry.go:8-14
MetricEuclidean, MetricCosine, MetricDotProduct = 0, 1, 2
# internal/simd/registry.go:31-36
DataTypeFloat32, DataTypeFloat16, DataTypeInt8, DataTypeUint8 = 0, 1, 2, 3

MaxFloat32 = np.finfo(np.float32).max
_DT = {np.dtype(np.float32): F32, np.dtype(np.float16): F16, np.dtype(np.int8): I8, np.dtype(np.uint8): U8}


class SimdError(ValueError):
    pass


def _flat(metric, query, flat, n, dims, results, device):
    query = np.ascontiguousarray(query)
    flat = np.ascontiguousarray(flat, dtype=query.dtype)
    check(_lib.load().lb_simd_distance_batch_flat(device, metric, _DT[query.dtype], query.ctypes.data,
                                                  flat.ctypes.data, n, dims, results.ctypes.data))


def EuclideanDistanceBatchFlat(query, flatVectors, numVectors, dims, results, device=0):
    """batch_operations.go:64-87 / simd.go:203-229."""
    if numVectors == 0:
        return
    if len(results) != numVectors:
        raise SimdError("simd: results length mismatch")
    if np.size(flatVectors) < numVectors * dims:
        raise SimdError("simd: flatVectors too small")
    if np.size(query) != dims:
        raise SimdError("simd: query dimension mismatch")
    _flat(MetricEuclidean, query, np.asarray(flatVectors).reshape(-1)[:numVectors * dims], numVectors, dims,
          results, device)


def _batch(metric, query, vectors, results, device, strict_len):
    n = len(vectors)
    if strict_len and n != len(results):
        raise SimdError("simd: vectors and results length mismatch")
    if n == 0:
        return
    if not strict_len and len(results) < n:
        raise SimdError("simd: results slice too small")
    query = np.ascontiguousarray(query)
    dims = query.size
    # nil / wrong-length rows get MaxFloat32 (batch_operations.go:38-41,49-52)
    ok = np.array([v is not None and len(v) == dims for v in vectors], bool)
    if ok.any():
        flat = np.stack([np.asarray(v, dtype=query.dtype) for v, o in zip(vectors, ok) if o])
        tmp = np.empty(int(ok.sum()), np.float32)
        _flat(metric, query, flat, flat.shape[0], dims, tmp, device)
        results[:n][ok] = tmp
    results[:n][~ok] = MaxFloat32


def EuclideanDistanceBatch(query, vectors, results, device=0):
    """batch_operations.go:29-60."""
    _batch(MetricEuclidean, query, vectors, results, device, True)


def CosineDistanceBatch(query, vectors, results, device=0):
    """batch_operations.go:131-142."""
    _batch(MetricCosine, query, vectors, results, device, False)


def DotProductBatch(query, vectors, results, device=0):
    """batch_operations.go:146-157 -- RAW dot products (not negated)."""
    _batch(MetricDotProduct, query, vectors, results, device, False)


def EuclideanDistanceF16Batch(query, vectors, results, device=0):
    """batch_operations.go:17-25."""
    _batch(MetricEuclidean, np.asarray(query, np.float16), vectors, results, device, True)


def EuclideanDistanceSQ8Batch(query, vectors, results, device=0):
    """batch_operations.go:107-115: squared L2 over uint8 as float32(int32)."""
    _batch(MetricEuclidean, np.asarray(query, np.uint8), vectors, results, device, True)


def ADCDistanceBatch(table, flatCodes, m, results, device=0):
    """batch_operations.go:119-127."""
    if np.size(table) == 0 or np.size(flatCodes) == 0:
        raise SimdError("simd: empty table or codes")
    if m <= 0:
        raise SimdError("simd: invalid m parameter")
    table = np.ascontiguousarray(table, np.float32)
    codes = np.ascontiguousarray(flatCodes, np.uint8)
    check(_lib.load().lb_simd_adc_distance_batch(device, table.ctypes.data, codes.ctypes.data, m, len(results),
                                                 results.ctypes.data))
