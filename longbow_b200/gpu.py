"""Mirror of the reference's ``internal/gpu`` package over the C ABI.

``Index`` / ``GPUConfig`` / ``NewIndex`` / ``NewIndexWithConfig`` / ``FaissGPUIndex`` keep the
reference's names, argument meaning and error behaviour (internal/gpu/interface.go:4-19,
internal/gpu/faiss_gpu.go:45-165, internal/gpu/gpu_enabled.go:9-21).  ``FaissGPUIndex`` binds
exactly the six C symbols the Go file links (faiss_gpu.go:16-21).

``DenseIndex`` is the extended handle (metric / dtype / batches / bitmaps / re-rank) that thin cgo
wrappers in internal/store would call.
"""
from __future__ import annotations

import ctypes as C
import threading
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import F16, F32, I8, METRIC_COSINE, METRIC_DOT, METRIC_L2, LongbowError, check

_NP_DTYPES = {F32: np.float32, F16: np.float16, I8: np.int8}
_DT_OF = {np.dtype(np.float32): F32, np.dtype(np.float16): F16, np.dtype(np.int8): I8}


@dataclass
class GPUConfig:  # internal/gpu/interface.go:4-7
    DeviceID: int = 0
    Dimension: int = 128


class Index:  # internal/gpu/interface.go:10-19
    def Add(self, ids, vectors):
        raise NotImplementedError

    def Search(self, vector, k):
        raise NotImplementedError

    def Close(self):
        raise NotImplementedError


class FaissGPUIndex(Index):
    """internal/gpu/faiss_gpu.go:35-165, call for call."""

    def __init__(self, cfg: GPUConfig):
        if cfg.Dimension <= 0:  # faiss_gpu.go:46-48
            raise ValueError(f"dimension must be positive, got {cfg.Dimension}")
        lib = _lib.load()
        self._lib = lib
        self.dim = cfg.Dimension
        self.deviceID = cfg.DeviceID
        self._mu = threading.RLock()
        self.closed = False
        self.resources = lib.faiss_gpu_resources_new(cfg.DeviceID)
        if not self.resources:  # :57-59
            raise RuntimeError(f"failed to initialize GPU resources for device {cfg.DeviceID}: "
                               f"{lib.lb_last_error().decode()}")
        self.index = lib.faiss_gpu_index_flat_l2_new(self.resources, cfg.Dimension)
        if not self.index:  # :63-66
            lib.faiss_gpu_resources_free(self.resources)
            self.resources = None
            raise RuntimeError("failed to create GPU index")

    def Add(self, ids, vectors):  # faiss_gpu.go:75-104
        with self._mu:
            if self.closed:
                raise RuntimeError("index is closed")
            vectors = np.ascontiguousarray(vectors, dtype=np.float32).reshape(-1)
            if vectors.size % self.dim != 0:
                raise ValueError(f"vector data length {vectors.size} not divisible by dimension {self.dim}")
            n = vectors.size // self.dim
            if len(ids) != n:
                raise ValueError(f"id count {len(ids)} does not match vector count {n}")
            if n == 0:
                return
            ret = self._lib.faiss_gpu_index_add(self.index, n, vectors.ctypes.data)
            if ret != 0:
                raise RuntimeError(f"GPU index add failed with code {ret}")

    def Search(self, vector, k):  # faiss_gpu.go:107-144
        with self._mu:
            if self.closed:
                raise RuntimeError("index is closed")
            vector = np.ascontiguousarray(vector, dtype=np.float32).reshape(-1)
            if vector.size != self.dim:
                raise ValueError(f"query vector dimension {vector.size} does not match index dimension {self.dim}")
            distances = np.empty(k, np.float32)
            labels = np.empty(k, np.int64)
            ret = self._lib.faiss_gpu_index_search(self.index, 1, vector.ctypes.data, k, distances.ctypes.data,
                                                   labels.ctypes.data)
            if ret != 0:
                raise RuntimeError(f"GPU search failed with code {ret}")
            return labels, distances

    def Close(self):  # faiss_gpu.go:147-165 (idempotent)
        with self._mu:
            if self.closed:
                return
            if self.index:
                self._lib.faiss_gpu_index_flat_l2_free(self.index)
                self.index = None
            if self.resources:
                self._lib.faiss_gpu_resources_free(self.resources)
                self.resources = None
            self.closed = True

    def __del__(self):  # runtime.SetFinalizer, faiss_gpu.go:69
        try:
            self.Close()
        except Exception:
            pass


def NewIndex() -> Index:  # gpu_enabled.go:9-14
    return NewIndexWithConfig(GPUConfig(DeviceID=0, Dimension=128))


def NewIndexWithConfig(cfg: GPUConfig) -> Index:  # gpu_enabled.go:17-21
    return FaissGPUIndex(cfg)


def _ptr(a):
    return None if a is None else a.ctypes.data


def _bitmap(bits, n):
    """Accept a bool/0-1 array over rows or an already packed uint64 array; return packed uint64."""
    if bits is None:
        return None
    bits = np.asarray(bits)
    if bits.dtype == np.uint64:
        need = (n + 63) // 64
        if bits.size < need:
            raise ValueError("bitmap shorter than the index")
        return np.ascontiguousarray(bits)
    return pack_bitmap(bits.astype(bool), n)


def pack_bitmap(mask: np.ndarray, n: int | None = None) -> np.ndarray:
    """bool[n] -> little-endian uint64 words, bit i <-> row i (the C ABI's bitmap layout)."""
    mask = np.asarray(mask, dtype=bool)
    n = mask.size if n is None else n
    words = (n + 63) // 64
    padded = np.zeros(words * 64, dtype=bool)
    padded[:mask.size] = mask[:n]
    return np.packbits(padded, bitorder="little").view(np.uint64).copy()


def dense_words(ids, nbits: int) -> np.ndarray:
    """A predicate set given as VectorIDs -- what the reference's roaring Bitset hands out through ToUint32Array
    (internal/query/bitmap.go:95-100) -- as the dense little-endian words the C ABI consumes; ids >= nbits are
    dropped.  Same contract as gpu.DenseWords in go/longbow_b200.go."""
    ids = np.asarray(ids, dtype=np.int64).reshape(-1)
    ids = ids[(ids >= 0) & (ids < nbits)]
    mask = np.zeros(int(nbits), dtype=bool)
    mask[ids] = True
    return pack_bitmap(mask, int(nbits))


class DenseIndex:
    """Extended dense handle: one Arrow FixedSizeList<T, dim> column mirrored in HBM."""

    def __init__(self, dim: int, dtype=np.float32, metric: int = METRIC_L2, device: int = 0):
        self._lib = _lib.load()
        self.dim = int(dim)
        self.np_dtype = np.dtype(dtype)
        if self.np_dtype not in _DT_OF:
            raise LongbowError(_lib.LB_ERR_UNSUPPORTED, f"no kernel for dtype {self.np_dtype}")
        self.dtype = _DT_OF[self.np_dtype]
        self.metric = int(metric)
        self.device = int(device)
        h = C.c_void_p()
        check(self._lib.lb_index_create(self.device, self.dim, self.dtype, self.metric, C.byref(h)))
        self._h = h

    # -- lifecycle
    def close(self):
        if getattr(self, "_h", None):
            self._lib.lb_index_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return int(self._lib.lb_index_size(self._h))

    def reserve(self, n: int):
        check(self._lib.lb_index_reserve(self._h, int(n)))

    def set_id_base(self, base: int):
        check(self._lib.lb_index_set_id_base(self._h, int(base)))

    # -- data
    def add(self, rows: np.ndarray):
        rows = np.ascontiguousarray(rows, dtype=self.np_dtype)
        if rows.size % self.dim != 0:
            raise ValueError(f"vector data length {rows.size} not divisible by dimension {self.dim}")
        check(self._lib.lb_index_add(self._h, _ptr(rows), rows.size // self.dim))

    def add_arrow(self, values: np.ndarray, list_offset: int, n_rows: int, pin: bool = True):
        """Append rows straight from an Arrow FixedSizeList child values buffer (a flat NumPy view of it):
        internal/store/arrow_utils.go:112-171, including the truncated-buffer rule."""
        values = np.ascontiguousarray(values, dtype=self.np_dtype).reshape(-1)
        check(self._lib.lb_index_add_arrow(self._h, _ptr(values), values.nbytes, int(list_offset), int(n_rows), int(pin)))

    def add_device(self, tensor, stream=None):
        """Append rows already resident on this device (a contiguous torch tensor)."""
        assert tensor.is_cuda and tensor.is_contiguous()
        n = tensor.numel() // self.dim
        check(self._lib.lb_index_add_device(self._h, tensor.data_ptr(), n, _stream_ptr(stream)))

    def set_tombstones(self, deleted):
        if deleted is None:
            check(self._lib.lb_index_set_tombstones(self._h, None, 0))
            return
        bm = _bitmap(deleted, len(self))
        check(self._lib.lb_index_set_tombstones(self._h, _ptr(bm), bm.size * 64))

    # -- search (host buffers)
    def search(self, queries: np.ndarray, k: int, allow=None):
        q = np.ascontiguousarray(queries, dtype=self.np_dtype).reshape(-1, self.dim)
        nq = q.shape[0]
        d = np.empty((nq, k), np.float32)
        l = np.empty((nq, k), np.int64)
        bm = _bitmap(allow, len(self))
        check(self._lib.lb_index_search(self._h, _ptr(q), nq, int(k), _ptr(bm), _ptr(d), _ptr(l)))
        return d, l

    def search_into(self, q: np.ndarray, k: int, d: np.ndarray, l: np.ndarray, allow_packed=None):
        """Zero-allocation variant for timing loops: caller-owned (ideally pinned) buffers."""
        check(self._lib.lb_index_search(self._h, q.ctypes.data, q.shape[0], int(k), _ptr(allow_packed),
                                        d.ctypes.data, l.ctypes.data))

    def rerank(self, queries: np.ndarray, cand_ids: np.ndarray, k: int, allow=None):
        q = np.ascontiguousarray(queries, dtype=self.np_dtype).reshape(-1, self.dim)
        ids = np.ascontiguousarray(cand_ids, dtype=np.uint32).reshape(q.shape[0], -1)
        nq, c = ids.shape
        d = np.empty((nq, k), np.float32)
        l = np.empty((nq, k), np.int64)
        bm = _bitmap(allow, len(self))
        check(self._lib.lb_index_rerank(self._h, _ptr(q), nq, _ptr(ids), c, int(k), _ptr(bm), _ptr(d), _ptr(l)))
        return d, l

    def distances(self, query: np.ndarray) -> np.ndarray:
        q = np.ascontiguousarray(query, dtype=self.np_dtype).reshape(self.dim)
        out = np.empty(len(self), np.float32)
        check(self._lib.lb_index_distances(self._h, _ptr(q), _ptr(out)))
        return out

    # -- search (device tensors, asynchronous on `stream`)
    def search_device(self, q, k: int, out_d, out_l, allow=None, stream=None, uncert_flags=None, uncert_count=None):
        """out_d / out_l: device tensors or raw device addresses (ints).  uncert_flags [nq] u32 / uncert_count [1]
        u32 (device tensors, optional): certification outputs of lb_index_search_device_cert -- the count
        ACCUMULATES, so one counter can watch a whole run."""
        pd = out_d if isinstance(out_d, int) else out_d.data_ptr()
        pl = out_l if isinstance(out_l, int) else out_l.data_ptr()
        pa = None if allow is None else allow.data_ptr()
        if uncert_flags is None and uncert_count is None:
            check(self._lib.lb_index_search_device(self._h, q.data_ptr(), q.shape[0], int(k), pa, pd, pl,
                                                   _stream_ptr(stream)))
        else:
            check(self._lib.lb_index_search_device_cert(
                self._h, q.data_ptr(), q.shape[0], int(k), pa, pd, pl,
                None if uncert_flags is None else uncert_flags.data_ptr(),
                None if uncert_count is None else uncert_count.data_ptr(), _stream_ptr(stream)))

    def search_exact_device(self, q, k: int, out_d, out_l, flags_host=None, allow=None, stream=None):
        """Exhaustive exact search of the queries whose host flag is set (None = all): the repair step for
        queries lb_index_search_device_cert flagged."""
        fl = None
        if flags_host is not None:
            fl = np.ascontiguousarray(flags_host, dtype=np.uint32)
        check(self._lib.lb_index_search_exact_device(self._h, q.data_ptr(), q.shape[0], int(k),
                                                     None if allow is None else allow.data_ptr(), _ptr(fl),
                                                     out_d.data_ptr(), out_l.data_ptr(), _stream_ptr(stream)))

    def rerank_device(self, q, cand_ids, k: int, out_d, out_l, allow=None, stream=None):
        check(self._lib.lb_index_rerank_device(self._h, q.data_ptr(), q.shape[0], cand_ids.data_ptr(),
                                               cand_ids.shape[1], int(k),
                                               None if allow is None else allow.data_ptr(), out_d.data_ptr(),
                                               out_l.data_ptr(), _stream_ptr(stream)))

    def coarse_keys(self, queries: np.ndarray, n_rows: int) -> np.ndarray:
        """Diagnostics: the tensor-core scan's coarse ranking keys for rows [0, n_rows) (tests only)."""
        q = np.ascontiguousarray(queries, dtype=self.np_dtype).reshape(-1, self.dim)
        out = np.empty((q.shape[0], n_rows), np.float32)
        check(self._lib.lb_index_coarse_keys(self._h, _ptr(q), q.shape[0], int(n_rows), _ptr(out)))
        return out

    def last_uncertified(self) -> int:
        return int(self._lib.lb_index_last_uncertified(self._h))


def _stream_ptr(stream):
    """torch.cuda.Stream | int | None -> cudaStream_t as an integer (None = torch's current stream)."""
    if stream is None:
        import torch
        return torch.cuda.current_stream().cuda_stream
    if isinstance(stream, int):
        return stream
    return stream.cuda_stream


class ShardedDenseIndex:
    """Row-sharded index over several GPUs driven from one process (``lb_shard_*``): what a Go host binds in
    place of ShardedHNSW's fan-out + merge (internal/store/sharded_hnsw.go:378-503).  Labels are global rows."""

    def __init__(self, devices, dim: int, dtype, metric: int, total_rows: int):
        self._lib = _lib.load()
        self.dim, self.np_dtype = int(dim), np.dtype(dtype)
        if self.np_dtype not in _DT_OF:
            raise LongbowError(_lib.LB_ERR_UNSUPPORTED, f"no kernel for dtype {self.np_dtype}")
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        check(self._lib.lb_shard_create(devs, len(devices), self.dim, _DT_OF[self.np_dtype], int(metric),
                                        int(total_rows), C.byref(h)))
        self._h = h

    def __len__(self):
        return int(self._lib.lb_shard_size(self._h))

    def rows_per_shard(self) -> int:
        return int(self._lib.lb_shard_rows_per_shard(self._h))

    def add(self, rows):
        rows = np.ascontiguousarray(rows, dtype=self.np_dtype)
        if rows.size % self.dim != 0:
            raise ValueError(f"vector data length {rows.size} not divisible by dimension {self.dim}")
        check(self._lib.lb_shard_add(self._h, _ptr(rows), rows.size // self.dim))

    def set_tombstones(self, deleted):
        if deleted is None:
            check(self._lib.lb_shard_set_tombstones(self._h, None, 0))
            return
        bm = _bitmap(deleted, len(self))
        check(self._lib.lb_shard_set_tombstones(self._h, _ptr(bm), bm.size * 64))

    def search(self, queries, k: int, allow=None):
        q = np.ascontiguousarray(queries, dtype=self.np_dtype).reshape(-1, self.dim)
        nq = q.shape[0]
        d = np.empty((nq, k), np.float32)
        l = np.empty((nq, k), np.int64)
        bm = None
        if allow is not None:
            # padded to whole shards so that every shard's word slice exists
            total = self.rows_per_shard() * int(self._lib.lb_shard_count(self._h))
            bm = _bitmap(allow, len(self))
            need = (total + 63) // 64
            if bm.size < need:
                bm = np.concatenate([bm, np.zeros(need - bm.size, np.uint64)])
        check(self._lib.lb_shard_search(self._h, _ptr(q), nq, int(k), _ptr(bm), _ptr(d), _ptr(l)))
        return d, l

    def last_uncertified(self) -> int:
        return int(self._lib.lb_shard_last_uncertified(self._h))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.lb_shard_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
