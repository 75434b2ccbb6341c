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


class HNSWGraph:
    """GPU mirror of the search half of ``ArrowHNSW`` (internal/store/arrow_hnsw.go): the adjacency arrays of one
    layer as they lie in ``GraphData`` (internal/store/types/graph_data.go:605-670) over the rows of a resident
    ``DenseIndex``.  ``SearchLayer`` is ``searchLayer`` (:1108-1385) for a batch of queries; ``Search`` adds the
    re-rank with tombstones + predicate bitmap applied in-kernel (parallel_search.go:147-365) -- BASELINE config 5.
    Graph construction and the upper-layer descent to the entry points stay on the host."""

    def __init__(self, index: DenseIndex, max_degree: int):
        import ctypes as C
        self._lib = _lib.load()
        self.index = index
        self.max_degree = int(max_degree)
        h = C.c_void_p()
        check(self._lib.lb_graph_create(index._h, self.max_degree, C.byref(h)))
        self._h = h

    def SetNeighbors(self, neighbors, counts=None):
        nb = np.ascontiguousarray(neighbors, np.uint32).reshape(-1, self.max_degree)
        ct = None if counts is None else np.ascontiguousarray(counts, np.int32)
        check(self._lib.lb_graph_set_layer(self._h, _ptr(nb), _ptr(ct), nb.shape[0]))

    def SearchLayer(self, queries, entry_points, ef: int):
        """-> (ids [nq, ef] uint32 ascending by (distance, id), 0xffffffff padded; distances; visited counts)."""
        q = np.ascontiguousarray(queries, self.index.np_dtype).reshape(-1, self.index.dim)
        ep = np.ascontiguousarray(entry_points, np.uint32).reshape(-1)
        nq = q.shape[0]
        ids = np.empty((nq, ef), np.uint32)
        d = np.empty((nq, ef), np.float32)
        vis = np.empty(nq, np.uint32)
        check(self._lib.lb_graph_search_layer(self._h, _ptr(q), nq, _ptr(ep), int(ef), _ptr(ids), _ptr(d), _ptr(vis)))
        return ids, d, vis

    def Search(self, queries, entry_points, ef: int, k: int, allow=None):
        from .gpu import _bitmap
        q = np.ascontiguousarray(queries, self.index.np_dtype).reshape(-1, self.index.dim)
        ep = np.ascontiguousarray(entry_points, np.uint32).reshape(-1)
        nq = q.shape[0]
        d = np.empty((nq, k), np.float32)
        l = np.empty((nq, k), np.int64)
        bm = _bitmap(allow, len(self.index))
        check(self._lib.lb_graph_search(self._h, _ptr(q), nq, _ptr(ep), int(ef), int(k), _ptr(bm), _ptr(d), _ptr(l)))
        return d, l

    def search_device(self, q, entry_points, ef: int, k: int, out_d, out_l, fail_count, allow=None, stream=None):
        from .gpu import _stream_ptr
        check(self._lib.lb_graph_search_device(self._h, q.data_ptr(), q.shape[0], entry_points.data_ptr(), int(ef), int(k),
                                               None if allow is None else allow.data_ptr(), out_d.data_ptr(),
                                               out_l.data_ptr(), fail_count.data_ptr(), _stream_ptr(stream)))

    def Close(self):
        if getattr(self, "_h", None):
            self._lib.lb_graph_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.Close()
        except Exception:
            pass


# ---- Arrow compute operator registry (internal/store/arrow_kernels.go:20-111): the two functions the reference
# registers with compute.GetFunctionRegistry(), backed by the GPU library.
_ARROW_FUNCTIONS = {}
_kernelsRegistered = False


def RegisterHNSWKernels():
    """Idempotent, like the reference's init() hook (arrow_kernels.go:20-31)."""
    global _kernelsRegistered
    if _kernelsRegistered:
        return
    _ARROW_FUNCTIONS["l2_distance"] = lambda left, right, device=0: L2DistanceOp(left, right, device)
    _ARROW_FUNCTIONS["select_k_neighbors"] = lambda distances, ids, k, device=0: _select_k_op(distances, ids, k, device)
    _kernelsRegistered = True


def _select_k_op(distances, ids, k, device=0):
    """selectKExec (arrow_kernels.go:230-345): INDICES (uint32) of the k smallest distances; `ids` only has to
    match in length (the caller applies Take)."""
    d = np.ascontiguousarray(distances, np.float32)
    if ids is not None and len(ids) != d.size:
        raise ValueError("distances and ids must have the same length")
    idx, _ = SelectTopKNeighbors(d, min(int(k), d.size), device)
    return idx.astype(np.uint32)


def CallFunction(name: str, *args, **kwargs):
    """compute.CallFunction(ctx, name, ...) for the registered vector-search functions."""
    RegisterHNSWKernels()
    if name not in _ARROW_FUNCTIONS:
        raise KeyError(f"function not found: {name}")
    return _ARROW_FUNCTIONS[name](*args, **kwargs)


def GenerateFilterBitsetBatches(batches, op: int, value, n_vector_ids: int, device: int = 0):
    """Dataset.GenerateFilterBitset (internal/store/dataset.go:247-300) over several record batches, on the device:
    ``batches`` = [(column (int64 / float32 array), vector_ids (uint32 array with 0xffffffff for rows the index
    does not know) or an int VectorID base for contiguously indexed batches), ...].  Per batch the predicate
    kernel builds the batch-local match bitmap and the scatter kernel sets the matching rows' VectorID bits in one
    global allow-bitmap (packed uint64, n_vector_ids bits), which can be passed to the searches as ``allow``."""
    import torch
    lib = _lib.load()
    dev = torch.device("cuda", device)
    st = torch.cuda.current_stream(dev).cuda_stream
    glob = torch.zeros((n_vector_ids + 63) // 64, dtype=torch.int64, device=dev)
    for column, ids in batches:
        col = torch.from_numpy(np.ascontiguousarray(column)).to(dev)
        n = col.numel()
        bm = GenerateFilterBitsetDevice(col, op, value, device=device)
        if isinstance(ids, (int, np.integer)):
            check(lib.lb_filter_scatter_device(device, bm.data_ptr(), n, None, int(ids), n_vector_ids, glob.data_ptr(), st))
        else:
            idt = torch.from_numpy(np.ascontiguousarray(ids, np.uint32).view(np.int32)).to(dev)
            check(lib.lb_filter_scatter_device(device, bm.data_ptr(), n, idt.data_ptr(), 0, n_vector_ids, glob.data_ptr(), st))
    return glob.cpu().numpy().view(np.uint64)
