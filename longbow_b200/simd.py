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


def EuclideanDistanceVerticalBatch(query, vectors, results, device=0):
    """batch_operations.go:91-104: same values as EuclideanDistanceBatch (the vertical form is a CPU layout trick)."""
    _batch(MetricEuclidean, query, vectors, results, device, True)


def _pair(metric, a, b, dtype, device):
    a, b = np.ascontiguousarray(a, dtype), np.ascontiguousarray(b, dtype)
    if a.size != b.size:
        raise SimdError("simd: vector length mismatch")  # distance_functions.go:18-20
    if a.size == 0:
        return np.float32(1.0) if metric == MetricCosine else np.float32(0.0)  # :21-23, :51-53
    out = np.empty(1, np.float32)
    _flat(metric, a.reshape(-1), b.reshape(1, -1), 1, a.size, out, device)
    return out[0]


def EuclideanDistance(a, b, device=0):
    """distance_functions.go:17-31."""
    return _pair(MetricEuclidean, a, b, np.float32, device)


def CosineDistance(a, b, device=0):
    """distance_functions.go:47-55."""
    return _pair(MetricCosine, a, b, np.float32, device)


def DotProduct(a, b, device=0):
    """distance_functions.go:59-73 (raw similarity)."""
    return _pair(MetricDotProduct, a, b, np.float32, device)


def EuclideanDistanceF16(a, b, device=0):
    """distance_functions.go:76-86."""
    return _pair(MetricEuclidean, a, b, np.float16, device)


def CosineDistanceF16(a, b, device=0):
    """distance_functions.go:88-98."""
    return _pair(MetricCosine, a, b, np.float16, device)


def DotProductF16(a, b, device=0):
    """distance_functions.go:100-109."""
    return _pair(MetricDotProduct, a, b, np.float16, device)


def CosineDistanceF16Batch(query, vectors, results, device=0):
    _batch(MetricCosine, np.asarray(query, np.float16), vectors, results, device, False)


def DotProductF16Batch(query, vectors, results, device=0):
    _batch(MetricDotProduct, np.asarray(query, np.float16), vectors, results, device, False)


def QuantizeSQ8(src, dst, minVal, maxVal, device=0):
    """sq8.go:70-86: dst must be pre-allocated with the source's length."""
    src = np.ascontiguousarray(src, np.float32).reshape(-1)
    if len(dst) < src.size:
        raise SimdError("simd: dst too small")
    out = np.empty(src.size, np.uint8)
    check(_lib.load().lb_simd_quantize_sq8(device, src.ctypes.data, src.size, float(minVal), float(maxVal),
                                           out.ctypes.data))
    dst[:src.size] = out


def DequantizeSQ8(src, minVal, maxVal, device=0):
    """The inline de-quantisation of the HNSW distance computer (internal/store/arrow_hnsw.go:1176-1186)."""
    src = np.ascontiguousarray(src, np.uint8).reshape(-1)
    out = np.empty(src.size, np.float32)
    check(_lib.load().lb_simd_dequantize_sq8(device, src.ctypes.data, src.size, float(minVal), float(maxVal),
                                             out.ctypes.data))
    return out


def ComputeBounds(vec, device=0):
    """sq8.go:88-104."""
    import ctypes as C
    v = np.ascontiguousarray(vec, np.float32).reshape(-1)
    mn, mx = C.c_float(), C.c_float()
    check(_lib.load().lb_simd_compute_bounds(device, v.ctypes.data, v.size, C.byref(mn), C.byref(mx)))
    return mn.value, mx.value


def SQ8DequantDistanceBatch(query, rows, minVal, maxVal, results, device=0):
    """One fp32 query against SQ8 rows as an SQ8-enabled ArrowHNSW ranks them (arrow_hnsw.go:1176-1186)."""
    q = np.ascontiguousarray(query, np.float32).reshape(-1)
    r = np.ascontiguousarray(rows, np.uint8).reshape(-1, q.size)
    check(_lib.load().lb_simd_sq8_dequant_distance_batch(device, q.ctypes.data, r.ctypes.data, r.shape[0], q.size,
                                                         float(minVal), float(maxVal), results.ctypes.data))


def FindNearestCentroid(query, centroids, subDim, k, device=0):
    """simd.go:278-326 -> (index, distance)."""
    import ctypes as C
    cent = np.ascontiguousarray(centroids, np.float32).reshape(-1)
    if cent.size < k * subDim:
        return 0, MaxFloat32  # simd.go:279-281
    q = np.ascontiguousarray(query, np.float32).reshape(-1)
    idx, d = C.c_int(), C.c_float()
    check(_lib.load().lb_simd_find_nearest_centroid(device, q.ctypes.data, cent.ctypes.data, int(subDim), int(k),
                                                    C.byref(idx), C.byref(d)))
    return idx.value, d.value
