"""Mirror of the reference's ``internal/pq`` package (query-time part) on the GPU.

``PQEncoder`` keeps the reference's field and method names (internal/pq/encoder.go:12-18,
adc_table.go:15-92, persistence.go:15-80).  Training (k-means) stays on the host side of the
boundary in this round.
"""
from __future__ import annotations

import ctypes as C
import struct

import numpy as np

from . import _lib
from ._lib import check
from .gpu import DenseIndex, _bitmap, _ptr, _stream_ptr


class PQEncoder:
    def __init__(self, dims: int, m: int, k: int, codebooks: np.ndarray, device: int = 0):
        if dims % m != 0:  # encoder.go:22-24
            raise ValueError("dimension must be divisible by M")
        self.Dims, self.M, self.K, self.SubDim = dims, m, k, dims // m
        self.Codebooks = np.ascontiguousarray(codebooks, np.float32).reshape(m, k, self.SubDim)
        self.device = device
        self._lib = _lib.load()
        blob = self.Serialize()
        h = C.c_void_p()
        check(self._lib.lb_pq_create(device, blob, len(blob), C.byref(h)))
        self._h = h
        self._raw = None

    def close(self):
        if getattr(self, "_h", None):
            self._lib.lb_pq_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # persistence.go:15-36
    def Serialize(self) -> bytes:
        return struct.pack("<III", self.Dims, self.M, self.K) + self.Codebooks.astype("<f4").tobytes()

    # persistence.go:38-80
    @staticmethod
    def Deserialize(data: bytes, device: int = 0) -> "PQEncoder":
        if len(data) < 12:
            raise ValueError("invalid PQ data: too short")
        dims, m, k = struct.unpack("<III", data[:12])
        if m == 0 or dims % m != 0:
            raise ValueError("invalid PQ parameters in serialized data")
        sub = dims // m
        if len(data) != 12 + m * k * sub * 4:
            raise ValueError("invalid PQ data: size mismatch")
        cb = np.frombuffer(data, "<f4", offset=12).reshape(m, k, sub)
        return PQEncoder(dims, m, k, cb, device)

    # encoder.go:39-73 + kmeans.go:64-151
    @staticmethod
    def Train(dims: int, m: int, k: int, vectors, init_idx=None, seed: int = 0, max_iter: int = 20,
              device: int = 0) -> "PQEncoder":
        """NewPQEncoder(dims, m, k) followed by Train(vectors): k-means per subspace on the GPU.

        ``init_idx`` [m, k]: the data rows each subspace's centroids start from (the reference draws them with
        ``rand.Perm``); by default a seeded NumPy permutation per subspace."""
        cb, _ = train_codebooks(dims, m, k, vectors, init_idx, seed, max_iter, device)
        return PQEncoder(dims, m, k, cb, device)

    def CodeSize(self) -> int:  # encoder.go:163-165
        return self.M

    def BuildADCTable(self, query) -> np.ndarray:  # adc_table.go:15-51
        q = np.ascontiguousarray(query, np.float32).reshape(-1)
        if q.size != self.M * self.SubDim:
            raise ValueError("query dimension mismatch")
        table = np.empty(self.M * self.K, np.float32)
        check(self._lib.lb_pq_build_adc_table(self._h, _ptr(q), _ptr(table)))
        return table

    ComputeDistanceTableFlat = BuildADCTable  # encoder.go:166-169

    def ADCDistanceBatch(self, table, flatCodes, results):  # adc_table.go:57-73
        if len(results) == 0:
            return
        flatCodes = np.ascontiguousarray(flatCodes, np.uint8).reshape(-1)
        if flatCodes.size < len(results) * self.M:
            raise ValueError("flatCodes buffer too small")
        if np.size(table) != self.M * self.K:
            raise ValueError("invalid table size")
        table = np.ascontiguousarray(table, np.float32)
        check(self._lib.lb_simd_adc_distance_batch(self.device, _ptr(table), _ptr(flatCodes), self.M, len(results),
                                                   _ptr(results)))

    def Encode(self, vector) -> np.ndarray:  # encoder.go:76-93
        v = np.ascontiguousarray(vector, np.float32).reshape(-1)
        if v.size != self.M * self.SubDim:
            raise ValueError("vector dimension mismatch")
        return self.EncodeBatch(v[None, :])[0]

    def EncodeBatch(self, vectors) -> np.ndarray:
        v = np.ascontiguousarray(vectors, np.float32).reshape(-1, self.Dims)
        codes = np.empty((v.shape[0], self.M), np.uint8)
        check(self._lib.lb_pq_encode(self._h, _ptr(v), v.shape[0], _ptr(codes)))
        return codes

    def Decode(self, codes) -> np.ndarray:  # encoder.go:139-160 (host-side gather; not on the hot path)
        codes = np.asarray(codes, np.uint8).reshape(-1)
        if codes.size != self.M:
            raise ValueError("code length mismatch")
        return np.concatenate([self.Codebooks[m, codes[m]] for m in range(self.M)])

    # ---- resident codes + scan
    def add_codes(self, codes):
        c = np.ascontiguousarray(codes, np.uint8).reshape(-1, self.M)
        check(self._lib.lb_pq_add_codes(self._h, _ptr(c), c.shape[0]))

    def add_codes_device(self, tensor, stream=None):
        assert tensor.is_cuda and tensor.is_contiguous()
        check(self._lib.lb_pq_add_codes_device(self._h, tensor.data_ptr(), tensor.numel() // self.M,
                                               _stream_ptr(stream)))

    def __len__(self):
        return int(self._lib.lb_pq_size(self._h))

    def attach_raw(self, raw: DenseIndex | None):
        check(self._lib.lb_pq_attach_raw(self._h, None if raw is None else raw._h))
        self._raw = raw

    def set_tombstones(self, deleted):
        if deleted is None:
            check(self._lib.lb_pq_set_tombstones(self._h, None, 0))
            return
        bm = _bitmap(deleted, len(self))
        check(self._lib.lb_pq_set_tombstones(self._h, _ptr(bm), bm.size * 64))

    def search(self, queries, k: int, kprime: int = 0, allow=None):
        q = np.ascontiguousarray(queries, np.float32).reshape(-1, self.Dims)
        d = np.empty((q.shape[0], k), np.float32)
        l = np.empty((q.shape[0], k), np.int64)
        bm = _bitmap(allow, len(self))
        check(self._lib.lb_pq_search(self._h, _ptr(q), q.shape[0], int(k), int(kprime), _ptr(bm), _ptr(d), _ptr(l)))
        return d, l

    def last_uncertified(self) -> int:
        """Queries of the last host search whose coarse pass could not be certified (they were re-done with the
        exhaustive fp32 kernel before the call returned)."""
        return int(self._lib.lb_pq_last_uncertified(self._h))

    def search_into(self, q, k, kprime, d, l):
        check(self._lib.lb_pq_search(self._h, q.ctypes.data, q.shape[0], int(k), int(kprime), None, d.ctypes.data,
                                     l.ctypes.data))

    def search_device(self, q, k: int, kprime: int, out_d, out_l, allow=None, stream=None):
        check(self._lib.lb_pq_search_device(self._h, q.data_ptr(), q.shape[0], int(k), int(kprime),
                                            None if allow is None else allow.data_ptr(), out_d.data_ptr(),
                                            out_l.data_ptr(), _stream_ptr(stream)))


def train_codebooks(dims: int, m: int, k: int, vectors, init_idx=None, seed: int = 0, max_iter: int = 20,
                    device: int = 0):
    """TrainKMeans for every subspace (internal/pq/kmeans.go:64-151); returns (codebooks [m,k,sub], iters [m])."""
    if dims % m != 0:
        raise ValueError("dimension must be divisible by M")
    v = np.ascontiguousarray(vectors, np.float32).reshape(-1, dims)
    n = v.shape[0]
    if n == 0:
        raise ValueError("empty training data")
    if n < k:
        raise ValueError("insufficient data for k-means: n < k")
    if init_idx is None:
        rng = np.random.RandomState(seed)
        init_idx = np.stack([rng.permutation(n)[:k] for _ in range(m)])
    init_idx = np.ascontiguousarray(init_idx, np.int32).reshape(m, k)
    cb = np.empty((m, k, dims // m), np.float32)
    iters = np.zeros(m, np.int32)
    check(_lib.load().lb_pq_train(device, _ptr(v), n, dims, m, k, max_iter, _ptr(init_idx), _ptr(cb), _ptr(iters)))
    return cb, iters
