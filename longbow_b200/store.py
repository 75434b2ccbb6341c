"""Mirrors of the ``internal/store`` callers of the hot path, routed through the GPU library.

Only the arithmetic + selection of these callers is rebuilt: the Arrow
dataset, locks, metrics and the HNSW graph walk stay on the Go side of the boundary.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import METRIC_L2, check
from .gpu import DenseIndex, _ptr


@dataclass
class SearchResult:  # internal/store/types/types.go:102-108
    ID: int
    Score: float
    Distance: float = 0.0


@dataclass
class RankedResult:  # internal/store/hnsw_batch.go:14-17
    ID: int
    Distance: float


class BruteForceIndex:
    """internal/store/adaptive_index.go:59-225: exact linear-scan k-NN over fp32 rows, Euclidean.

    ``SearchVectors`` returns ascending ``SearchResult{ID: row, Score: distance}`` (Distance left 0
    as in adaptive_index.go:215-222); filters are ignored (as in the reference, :161-166).
    """

    def __init__(self, dims: int, device: int = 0):
        self._idx = DenseIndex(dims, np.float32, METRIC_L2, device)

    def AddBatch(self, vectors):
        self._idx.add(vectors)

    def Len(self) -> int:
        return len(self._idx)

    def SearchVectors(self, q, k: int):
        q = np.asarray(q)
        if q.dtype != np.float32:  # adaptive_index.go:162-166
            raise TypeError("BruteForceIndex only supports []float32 queries")
        if len(self._idx) == 0:
            return None
        d, l = self._idx.search(q.reshape(1, -1), k)
        return [SearchResult(ID=int(i), Score=float(s)) for s, i in zip(d[0], l[0]) if i >= 0]

    def SearchBatch(self, queries, k: int):
        d, l = self._idx.search(queries, k)
        return [[SearchResult(ID=int(i), Score=float(s)) for s, i in zip(dr, lr) if i >= 0] for dr, lr in zip(d, l)]

    def Close(self):
        self._idx.close()


def RerankBatch(index: DenseIndex, query, candidateIDs, k: int):
    """internal/store/hnsw_batch.go:206-245 (one query)."""
    query = np.asarray(query)
    ids = np.asarray(candidateIDs, np.uint32)
    if query.size == 0 or ids.size == 0 or k <= 0:
        return None
    d, l = index.rerank(query.reshape(1, -1), ids.reshape(1, -1), k)
    return [RankedResult(ID=int(i), Distance=float(x)) for x, i in zip(d[0], l[0]) if i >= 0]


def SearchHybrid(gpuIndex, query, k: int, n_vectors: int, locations=None):
    """ArrowHNSW.SearchHybrid (internal/store/hnsw_gpu.go:73-125): the GPU index generates k*10 candidates
    (capped at the index length), labels are VectorIDs, ids whose location lookup fails or whose location is
    tombstoned (BatchIdx == -1) are skipped, the first k survivors are returned with Score = GPU distance.

    ``gpuIndex``: a ``gpu.Index`` (``Search(vector, k) -> (ids, distances)``).  ``locations``: optional
    sequence / array of BatchIdx per VectorID (missing id or -1 = dropped), as ChunkedLocationStore answers."""
    candidateCount = min(k * 10, n_vectors)
    if candidateCount <= 0:
        return []
    ids, dists = gpuIndex.Search(np.asarray(query, np.float32), candidateCount)
    out = []
    for i, d in zip(ids, dists):
        if len(out) >= k:
            break
        i = int(i)
        if i < 0:
            continue  # padding label -1 -> VectorID(0xFFFFFFFF) -> location miss (hnsw_gpu.go:108-111)
        if locations is not None and (i >= len(locations) or int(locations[i]) == -1):
            continue
        out.append(SearchResult(ID=i, Score=float(d)))
    return out


def GenerateFilterBitsetDevice(d_column, op: int, value, d_bitmap=None, device: int = 0, stream=None):
    """Device-resident form of GenerateFilterBitset: ``d_column`` is a CUDA tensor (int64 or float32), the bitmap
    a CUDA int64 tensor of ceil(n/64) words (created zeroed when not given; AND-combined when given).  The result
    can be passed straight to ``DenseIndex.search_device(..., allow=...)``."""
    import torch
    from .gpu import _stream_ptr
    n = d_column.numel()
    and_into = d_bitmap is not None
    if d_bitmap is None:
        d_bitmap = torch.zeros((n + 63) // 64, dtype=torch.int64, device=d_column.device)
    lib = _lib.load()
    if d_column.dtype == torch.int64:
        check(lib.lb_filter_i64_device(device, d_column.data_ptr(), n, op, int(value), int(and_into),
                                       d_bitmap.data_ptr(), _stream_ptr(stream)))
    elif d_column.dtype == torch.float32:
        check(lib.lb_filter_f32_device(device, d_column.data_ptr(), n, op, float(value), int(and_into),
                                       d_bitmap.data_ptr(), _stream_ptr(stream)))
    else:
        raise TypeError("filter columns: int64 or float32")
    return d_bitmap


def MergeShardResults(distances, labels, k: int, device: int = 0):
    """Tail of ShardedHNSW.SearchVectors (internal/store/sharded_hnsw.go:432-503) /
    MergeSortedStreams (internal/store/result_merger.go:34-100), keyed on (distance, id).

    distances/labels: [parts, nq, k_in] per-shard top lists with GLOBAL ids (label -1 = padding).
    """
    d = np.ascontiguousarray(distances, np.float32)
    l = np.ascontiguousarray(labels, np.int64)
    parts, nq, k_in = d.shape
    od = np.empty((nq, k), np.float32)
    ol = np.empty((nq, k), np.int64)
    check(_lib.load().lb_merge_topk(device, _ptr(d), _ptr(l), parts, nq, k_in, k, _ptr(od), _ptr(ol)))
    return od, ol


def SelectTopKNeighbors(distances, k: int, device: int = 0):
    """Arrow compute ``select_k_neighbors`` (internal/store/arrow_kernels.go:230-345,
    arrow_neighbors.go:23-120): indices of the k smallest distances, ascending."""
    d = np.ascontiguousarray(distances, np.float32)
    idx = np.empty(k, np.int64)
    od = np.empty(k, np.float32)
    check(_lib.load().lb_select_k(device, _ptr(d), d.size, k, _ptr(idx), _ptr(od)))
    keep = idx >= 0
    return idx[keep], od[keep]


def L2DistanceOp(query, rows, device: int = 0):
    """Arrow compute ``l2_distance`` (internal/store/arrow_kernels.go:114-211): scalar (+) array broadcast."""
    from .simd import EuclideanDistanceBatchFlat
    rows = np.ascontiguousarray(rows, np.float32)
    out = np.empty(rows.shape[0], np.float32)
    EuclideanDistanceBatchFlat(np.asarray(query, np.float32), rows, rows.shape[0], rows.shape[1], out, device)
    return out


def GenerateFilterBitset(column, op: int, value, bitmap=None, device: int = 0):
    """Predicate -> dense allow-bitmap (internal/query/filter_evaluator.go:700-758 + simd.MatchInt64 /
    MatchFloat32).  op follows simd.CompareOp (internal/simd/simd.go:38-45).  AND-combines into
    ``bitmap`` when given."""
    col = np.ascontiguousarray(column)
    n = col.size
    words = (n + 63) // 64
    and_into = bitmap is not None
    bm = np.ascontiguousarray(bitmap, np.uint64) if and_into else np.zeros(words, np.uint64)
    lib = _lib.load()
    if col.dtype == np.int64:
        check(lib.lb_filter_i64(device, _ptr(col), n, op, int(value), int(and_into), _ptr(bm)))
    elif col.dtype == np.float32:
        check(lib.lb_filter_f32(device, _ptr(col), n, op, float(value), int(and_into), _ptr(bm)))
    else:
        raise TypeError("filter columns: int64 or float32")
    return bm
