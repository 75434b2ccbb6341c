"""ctypes front-end for the CPU oracle (oracle/lb_oracle.c, oracle/lb_oracle_fast.c).

TEST INFRASTRUCTURE ONLY -- importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / ``--impl reference`` legs.  The product package
(longbow_b200/) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

L2, COSINE, DOT = 0, 1, 2
F32, F16, I8, U8 = 0, 1, 2, 3

_NP2DT = {np.dtype(np.float32): F32, np.dtype(np.float16): F16, np.dtype(np.int8): I8,
          np.dtype(np.uint8): U8}


def build(force: bool = False) -> None:
    """Compile both oracle libraries with oracle/Makefile (gcc only)."""
    need = force or not all(os.path.exists(os.path.join(_HERE, n))
                            for n in ("liblb_oracle.so", "liblb_oracle_fast.so"))
    if not need:
        for n, s in (("liblb_oracle.so", "lb_oracle.c"), ("liblb_oracle_fast.so", "lb_oracle_fast.c")):
            if os.path.getmtime(os.path.join(_HERE, s)) > os.path.getmtime(os.path.join(_HERE, n)):
                need = True
    if need:
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])


def _load(name: str) -> C.CDLL:
    path = os.path.join(_HERE, name)
    if not os.path.exists(path):
        build()
    return C.CDLL(path)


_exact = None
_fast = None


def exact() -> C.CDLL:
    global _exact
    if _exact is None:
        lib = _load("liblb_oracle.so")
        f = C.c_float
        for n in ("lbo_euclid_f32", "lbo_cosine_f32", "lbo_dot_f32", "lbo_l2sq_f32", "lbo_cosine_f32_seq",
                  "lbo_dot_f32_seq", "lbo_euclid_f16", "lbo_cosine_f16", "lbo_dot_f16", "lbo_euclid_i8",
                  "lbo_dot_i8", "lbo_distance", "lbo_adc_single", "lbo_h2f"):
            getattr(lib, n).restype = f
        lib.lbo_l2sq_u8.restype = C.c_int32
        _exact = lib
    return _exact


def fast() -> C.CDLL:
    global _fast
    if _fast is None:
        lib = _load("liblb_oracle_fast.so")
        lib.lbf_distance.restype = C.c_float
        _fast = lib
    return _fast


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, dt=None):
    a = np.ascontiguousarray(a) if dt is None else np.ascontiguousarray(a, dtype=dt)
    return a


def dtype_code(a: np.ndarray) -> int:
    return _NP2DT[a.dtype]


def distance(metric: int, a: np.ndarray, b: np.ndarray) -> float:
    a, b = _c(a), _c(b)
    assert a.dtype == b.dtype and a.shape == b.shape
    return float(exact().lbo_distance(metric, dtype_code(a), _p(a), _p(b), C.c_int(a.size)))


def raw(name: str, a: np.ndarray, b: np.ndarray):
    """Call a single-pair kernel by symbol name (e.g. 'lbo_euclid_f32')."""
    a, b = _c(a), _c(b)
    return getattr(exact(), name)(_p(a), _p(b), C.c_int(a.size))


def batch_flat(metric: int, q: np.ndarray, flat: np.ndarray, impl: str = "exact") -> np.ndarray:
    q, flat = _c(q), _c(flat)
    n, dim = flat.shape
    out = np.empty(n, np.float32)
    if impl == "exact":
        exact().lbo_batch_flat(metric, dtype_code(flat), _p(q), _p(flat), C.c_int64(n), dim, _p(out))
    else:
        fast().lbf_batch_flat(metric, dtype_code(flat), _p(q), _p(flat), C.c_int64(n), dim, _p(out))
    return out


def search(metric: int, db: np.ndarray, queries: np.ndarray, k: int, tomb=None, allow=None,
           id_base: int = 0, impl: str = "exact"):
    db, queries = _c(db), _c(queries)
    n, dim = db.shape
    nq = queries.shape[0]
    od = np.empty((nq, k), np.float32)
    oi = np.empty((nq, k), np.int64)
    fn = exact().lbo_search if impl == "exact" else fast().lbf_search
    rc = fn(metric, dtype_code(db), _p(db), C.c_int64(n), dim, _p(queries), C.c_int64(nq), k,
            _p(tomb), _p(allow), C.c_int64(id_base), _p(od), _p(oi))
    assert rc == 0
    return od, oi


def rerank(metric: int, db: np.ndarray, queries: np.ndarray, cand: np.ndarray, k: int, tomb=None,
           allow=None, impl: str = "exact"):
    db, queries = _c(db), _c(queries)
    cand = _c(cand, np.int64)
    n, dim = db.shape
    nq, c = cand.shape
    od = np.empty((nq, k), np.float32)
    oi = np.empty((nq, k), np.int64)
    fn = exact().lbo_rerank if impl == "exact" else fast().lbf_rerank
    rc = fn(metric, dtype_code(db), _p(db), C.c_int64(n), dim, _p(queries), C.c_int64(nq), _p(cand), c, k,
            _p(tomb), _p(allow), _p(od), _p(oi))
    assert rc == 0
    return od, oi


def merge(in_d: np.ndarray, in_id: np.ndarray, k: int):
    in_d, in_id = _c(in_d, np.float32), _c(in_id, np.int64)
    parts, nq, k_in = in_d.shape
    od = np.empty((nq, k), np.float32)
    oi = np.empty((nq, k), np.int64)
    rc = exact().lbo_merge(_p(in_d), _p(in_id), parts, C.c_int64(nq), k_in, k, _p(od), _p(oi))
    assert rc == 0
    return od, oi


def select_k(d: np.ndarray, k: int):
    d = _c(d, np.float32)
    oi = np.empty(k, np.int64)
    od = np.empty(k, np.float32)
    exact().lbo_select_k(_p(d), C.c_int64(d.size), k, _p(oi), _p(od))
    return od, oi


def adc_table(q: np.ndarray, codebooks: np.ndarray) -> np.ndarray:
    """codebooks: [M, K, sub] fp32; returns table [M*K]."""
    q, codebooks = _c(q, np.float32), _c(codebooks, np.float32)
    M, K, sub = codebooks.shape
    table = np.empty(M * K, np.float32)
    exact().lbo_adc_table(_p(q), _p(codebooks), M, K, sub, _p(table))
    return table


def adc_batch(table: np.ndarray, codes: np.ndarray, impl: str = "exact") -> np.ndarray:
    table, codes = _c(table, np.float32), _c(codes, np.uint8)
    n, M = codes.shape
    out = np.empty(n, np.float32)
    if impl == "exact":
        exact().lbo_adc_batch(_p(table), _p(codes), M, C.c_int64(n), _p(out))
    else:
        fast().lbf_adc_batch(_p(table), _p(codes), M, C.c_int64(n), _p(out))
    return out


def adc_single(table: np.ndarray, code: np.ndarray, K: int) -> float:
    table, code = _c(table, np.float32), _c(code, np.uint8)
    return float(exact().lbo_adc_single(_p(table), _p(code), code.size, K))


def pq_encode(vecs: np.ndarray, codebooks: np.ndarray) -> np.ndarray:
    vecs, codebooks = _c(vecs, np.float32), _c(codebooks, np.float32)
    M, K, sub = codebooks.shape
    n = vecs.shape[0]
    codes = np.empty((n, M), np.uint8)
    exact().lbo_pq_encode_batch(_p(vecs), C.c_int64(n), _p(codebooks), M, K, sub, _p(codes))
    return codes


def pq_train(data: np.ndarray, M: int, K: int, init_idx: np.ndarray, max_iter: int = 20):
    """TrainKMeans per subspace (internal/pq/kmeans.go:64-151) from explicit initial rows; returns
    (codebooks [M, K, sub], iterations run per subspace)."""
    data = _c(data, np.float32)
    n, dims = data.shape
    init_idx = _c(init_idx, np.int32).reshape(M, K)
    out = np.empty((M, K, dims // M), np.float32)
    iters = np.zeros(M, np.int32)
    rc = exact().lbo_pq_train(_p(data), C.c_int64(n), dims, M, K, max_iter, _p(init_idx), _p(out), _p(iters))
    assert rc == 0
    return out, iters


def pq_decode(codes: np.ndarray, codebooks: np.ndarray) -> np.ndarray:
    codes, codebooks = _c(codes, np.uint8), _c(codebooks, np.float32)
    M, K, sub = codebooks.shape
    out = np.empty((codes.shape[0], M * sub), np.float32)
    for i in range(codes.shape[0]):
        exact().lbo_pq_decode(_p(codes[i]), _p(codebooks), M, K, sub, _p(out[i]))
    return out


def pq_search(codebooks: np.ndarray, codes: np.ndarray, raw_vecs, queries: np.ndarray, k: int, kprime: int,
              tomb=None, allow=None, impl: str = "exact"):
    codebooks, codes, queries = _c(codebooks, np.float32), _c(codes, np.uint8), _c(queries, np.float32)
    raw_vecs = None if raw_vecs is None else _c(raw_vecs, np.float32)
    M, K, sub = codebooks.shape
    n = codes.shape[0]
    nq = queries.shape[0]
    od = np.empty((nq, k), np.float32)
    oi = np.empty((nq, k), np.int64)
    if impl == "exact":
        rc = exact().lbo_pq_search(_p(codebooks), M, K, sub, _p(codes), C.c_int64(n), _p(raw_vecs), _p(queries),
                                   C.c_int64(nq), k, kprime, _p(tomb), _p(allow), _p(od), _p(oi))
    else:
        assert tomb is None and allow is None
        rc = fast().lbf_pq_search(_p(codebooks), M, sub, _p(codes), C.c_int64(n), _p(raw_vecs), _p(queries),
                                  C.c_int64(nq), k, kprime, _p(od), _p(oi))
    assert rc == 0
    return od, oi


def hnsw_search_layer(metric: int, db: np.ndarray, neighbors: np.ndarray, counts, queries: np.ndarray,
                      entries: np.ndarray, ef: int):
    """ArrowHNSW.searchLayer (internal/store/arrow_hnsw.go:1108-1385) for a batch: returns (ids [nq, ef] uint32
    ascending by (distance, id), 0xffffffff padded; distances [nq, ef]; visited counts [nq])."""
    db, queries = _c(db), _c(queries)
    neighbors = _c(neighbors, np.uint32)
    n, dim = db.shape
    max_degree = neighbors.shape[1]
    counts = None if counts is None else _c(counts, np.int32)
    entries = _c(entries, np.uint32)
    nq = queries.shape[0]
    ids = np.empty((nq, ef), np.uint32)
    d = np.empty((nq, ef), np.float32)
    nv = np.zeros(nq, np.int64)
    rc = exact().lbo_hnsw_search_layer_batch(metric, dtype_code(db), _p(db), C.c_int64(n), dim, _p(neighbors),
                                            _p(counts), max_degree, _p(queries), C.c_int64(nq), _p(entries), ef,
                                            _p(ids), _p(d), _p(nv))
    assert rc == 0
    return ids, d, nv


def fast_threads() -> int:
    return int(fast().lbf_threads())


def fast_use_all_cores() -> int:
    """Give the timed CPU baseline every core this process may run on, whatever OMP_NUM_THREADS says
    (torchrun exports OMP_NUM_THREADS=1 to its ranks).  Returns the thread count now in use."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    fast().lbf_set_threads(int(n))
    return fast_threads()


def fast_isa() -> str:
    return {2: "avx512", 1: "avx2", 0: "scalar"}[int(fast().lbf_isa())]


def quantize_sq8(src: np.ndarray, minv: float, maxv: float) -> np.ndarray:
    src = _c(src, np.float32).reshape(-1)
    dst = np.empty(src.size, np.uint8)
    exact().lbo_quantize_sq8(_p(src), C.c_int64(src.size), C.c_float(minv), C.c_float(maxv), _p(dst))
    return dst


def dequantize_sq8(src: np.ndarray, minv: float, maxv: float) -> np.ndarray:
    src = _c(src, np.uint8).reshape(-1)
    dst = np.empty(src.size, np.float32)
    exact().lbo_dequantize_sq8(_p(src), C.c_int64(src.size), C.c_float(minv), C.c_float(maxv), _p(dst))
    return dst


def compute_bounds(v: np.ndarray):
    v = _c(v, np.float32).reshape(-1)
    mn, mx = C.c_float(), C.c_float()
    exact().lbo_compute_bounds(_p(v), C.c_int64(v.size), C.byref(mn), C.byref(mx))
    return mn.value, mx.value


def sq8_dequant_distance_batch(q: np.ndarray, rows: np.ndarray, minv: float, maxv: float) -> np.ndarray:
    q, rows = _c(q, np.float32), _c(rows, np.uint8)
    out = np.empty(rows.shape[0], np.float32)
    exact().lbo_sq8_dequant_distance_batch(_p(q), _p(rows), C.c_int64(rows.shape[0]), rows.shape[1], C.c_float(minv),
                                           C.c_float(maxv), _p(out))
    return out


def find_nearest_centroid(query: np.ndarray, centroids: np.ndarray):
    query, centroids = _c(query, np.float32), _c(centroids, np.float32)
    k, sub = centroids.shape
    d = C.c_float()
    i = exact().lbo_find_nearest_centroid(_p(query), _p(centroids), sub, k, C.byref(d))
    return int(i), d.value
