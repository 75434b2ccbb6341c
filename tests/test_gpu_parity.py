"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded
inputs.  Run on the B200 box with ``pytest -m gpu``."""
import threading

import numpy as np
import pytest

from tests.util import assert_topk_equal, make_db, random_bitmap

pytestmark = pytest.mark.gpu

L2, COS, DOT = 0, 1, 2


@pytest.fixture(scope="module")
def lbgpu():
    from longbow_b200 import gpu
    return gpu


COMBOS = [(np.float32, L2), (np.float32, COS), (np.float32, DOT),
          (np.float16, L2), (np.float16, COS), (np.float16, DOT),
          (np.int8, L2), (np.int8, DOT)]


# ------------------------------------------------------------------ reference known answers via the drop-in ABI
def test_faiss_abi_gpu_smoke_kat(lbgpu):
    # internal/gpu/gpu_test.go:12-47
    idx = lbgpu.NewIndexWithConfig(lbgpu.GPUConfig(DeviceID=0, Dimension=128))
    vectors = (np.arange(128 * 10, dtype=np.float32) * np.float32(0.01))
    idx.Add(list(range(10)), vectors)
    ids, dist = idx.Search(vectors[:128], 5)
    assert len(ids) == 5 and len(dist) == 5
    assert ids[0] == 0 and dist[0] < 0.01
    assert list(ids) == [0, 1, 2, 3, 4]
    with pytest.raises(ValueError):
        idx.Search(vectors[:64], 5)
    with pytest.raises(ValueError):
        idx.Add([1], vectors[:100])
    idx.Close()
    idx.Close()  # idempotent, faiss_gpu.go:151-153
    with pytest.raises(RuntimeError, match="closed"):
        idx.Search(vectors[:128], 5)


def test_simd_kats_on_gpu():
    # internal/simd/simd_dispatch_test.go:56-132, simd_test.go:146-194
    from longbow_b200 import simd
    q = np.array([1, 2, 3, 4], np.float32)
    out = np.zeros(2, np.float32)
    simd.EuclideanDistanceBatch(q, [np.array([5, 6, 7, 8], np.float32), q.copy()], out)
    assert out[0] == 8.0 and out[1] == 0.0
    simd.DotProductBatch(q, [np.array([5, 6, 7, 8], np.float32)], out)
    assert out[0] == 70.0
    simd.CosineDistanceBatch(np.zeros(4, np.float32), [q], out)
    assert out[0] == 1.0  # zero vector: exactly 1.0
    simd.CosineDistanceBatch(q, [-q], out)
    assert abs(out[0] - 2.0) <= 1e-5
    o3 = np.zeros(3, np.float32)
    simd.EuclideanDistanceBatch(q, [None, np.zeros(3, np.float32), q], o3)  # nil / wrong length -> MaxFloat32
    assert o3[0] == simd.MaxFloat32 and o3[1] == simd.MaxFloat32 and o3[2] == 0.0
    a = np.array([10, 20], np.int8)
    b = np.array([[10, 30]], np.int8)
    o1 = np.zeros(1, np.float32)
    simd._flat(simd.MetricEuclidean, a, b, 1, 2, o1, 0)
    assert o1[0] == 10.0
    u = np.array([0, 255, 3], np.uint8)
    simd.EuclideanDistanceSQ8Batch(u, [np.array([255, 0, 1], np.uint8)], o1)
    assert o1[0] == float(255 * 255 * 2 + 4)


# ------------------------------------------------------------------ dense brute force
@pytest.mark.parametrize("dtype,metric", COMBOS)
@pytest.mark.parametrize("n,dim,nq,k", [(5000, 128, 33, 10), (3000, 96, 70, 100), (700, 7, 5, 3), (1030, 33, 17, 10)])
def test_dense_search_parity(lbgpu, oracle, dtype, metric, n, dim, nq, k):
    rng = np.random.default_rng(1000 + n + dim + metric)
    db = make_db(rng, n, dim, dtype)
    q = make_db(rng, nq, dim, dtype)
    if metric == COS and dtype != np.int8:
        db[[3, n // 2, n - 1]] = 0  # zero rows: cosine distance exactly 1.0
        q[0] = 0
    idx = lbgpu.DenseIndex(dim, dtype, metric)
    idx.add(db[: n // 3])
    idx.add(db[n // 3:])  # two appends (growth path)
    assert len(idx) == n
    gd, gl = idx.search(q, k)
    wd, wl = oracle.search(metric, db, q, k)
    assert_topk_equal(gd, gl, wd, wl, 0.0, f"{dtype.__name__} metric={metric}")
    idx.close()


@pytest.mark.parametrize("dtype,metric", [(np.float32, L2), (np.float16, COS), (np.int8, DOT)])
def test_dense_search_bitmaps(lbgpu, oracle, dtype, metric):
    rng = np.random.default_rng(77)
    n, dim, nq, k = 6000, 64, 40, 10
    db, q = make_db(rng, n, dim, dtype), make_db(rng, nq, dim, dtype)
    tomb = random_bitmap(rng, n, 0.05)
    allow = random_bitmap(rng, n, 0.30)
    idx = lbgpu.DenseIndex(dim, dtype, metric)
    idx.add(db)
    idx.set_tombstones(tomb)
    gd, gl = idx.search(q, k, allow=allow)
    wd, wl = oracle.search(metric, db, q, k, tomb=lbgpu.pack_bitmap(tomb), allow=lbgpu.pack_bitmap(allow))
    assert_topk_equal(gd, gl, wd, wl, 0.0, "bitmaps")
    assert not tomb[gl[gl >= 0]].any() and allow[gl[gl >= 0]].all()
    # almost everything filtered: fewer than k survivors -> -1 / FLT_MAX padding
    allow2 = np.zeros(n, bool)
    allow2[[5, 17, 4000]] = True
    idx.set_tombstones(None)
    gd, gl = idx.search(q, k, allow=allow2)
    wd, wl = oracle.search(metric, db, q, k, allow=lbgpu.pack_bitmap(allow2))
    assert_topk_equal(gd, gl, wd, wl, 0.0, "sparse allow")
    assert (gl[:, 3:] == -1).all() and (gd[:, 3:] == np.finfo(np.float32).max).all()
    idx.close()


def test_dense_ties_and_edges(lbgpu, oracle):
    # duplicates: ties resolved by id (adaptive_index.go:200-211 keeps the lowest ids)
    dim = 16
    base = np.random.default_rng(5).random((50, dim), dtype=np.float32)
    db = np.concatenate([base, base, base])  # every row three times
    idx = lbgpu.DenseIndex(dim, np.float32, L2)
    idx.add(db)
    gd, gl = idx.search(base[:7], 6)
    wd, wl = oracle.search(L2, db, base[:7], 6)
    assert_topk_equal(gd, gl, wd, wl, 0.0, "ties")
    assert (gl[:, 0] == np.arange(7)).all() and (gl[:, 1] == np.arange(7) + 50).all()
    # k larger than the index
    gd, gl = idx.search(base[:2], 200)
    wd, wl = oracle.search(L2, db, base[:2], 200)
    assert_topk_equal(gd, gl, wd, wl, 0.0, "k > n")
    # empty query batch / empty index
    d0, l0 = idx.search(np.zeros((0, dim), np.float32), 5)
    assert d0.shape == (0, 5)
    e = lbgpu.DenseIndex(dim, np.float32, L2)
    d1, l1 = e.search(base[:3], 4)
    assert (l1 == -1).all() and (d1 == np.finfo(np.float32).max).all()
    e.close()
    idx.close()


def test_id_base_and_distances(lbgpu, oracle):
    rng = np.random.default_rng(9)
    db, q = make_db(rng, 900, 40, np.float32), make_db(rng, 4, 40, np.float32)
    idx = lbgpu.DenseIndex(40, np.float32, L2)
    idx.add(db)
    idx.set_id_base(10_000_000_000)
    gd, gl = idx.search(q, 5)
    wd, wl = oracle.search(L2, db, q, 5, id_base=10_000_000_000)
    assert_topk_equal(gd, gl, wd, wl, 0.0, "id_base")
    assert np.array_equal(idx.distances(q[0]), oracle.batch_flat(L2, q[0], db))
    idx.close()


def test_concurrent_search_threads(lbgpu, oracle):
    # faiss_gpu.go:40,108: Search takes a read lock -> many concurrent callers
    rng = np.random.default_rng(11)
    db = make_db(rng, 4000, 64, np.float32)
    idx = lbgpu.DenseIndex(64, np.float32, L2)
    idx.add(db)
    qs = [make_db(np.random.default_rng(100 + t), 8, 64, np.float32) for t in range(8)]
    want = [oracle.search(L2, db, q, 10) for q in qs]
    got = [None] * 8

    def work(t):
        for _ in range(5):
            got[t] = idx.search(qs[t], 10)

    th = [threading.Thread(target=work, args=(t,)) for t in range(8)]
    [t.start() for t in th]
    [t.join() for t in th]
    for t in range(8):
        assert_topk_equal(got[t][0], got[t][1], want[t][0], want[t][1], 0.0, f"thread {t}")
    idx.close()


# ------------------------------------------------------------------ simd flat batch (all dtypes)
@pytest.mark.parametrize("dtype,metric", COMBOS)
@pytest.mark.parametrize("dim", [1, 3, 7, 8, 15, 16, 31, 32, 64, 128, 256, 384, 512, 768, 1024, 1536])
def test_batch_flat_exact(oracle, dtype, metric, dim):
    from longbow_b200 import simd
    rng = np.random.default_rng(dim * 7 + metric)
    flat, q = make_db(rng, 37, dim, dtype), make_db(rng, 1, dim, dtype)[0]
    out = np.empty(37, np.float32)
    simd._flat(metric, q, flat, 37, dim, out, 0)
    want = oracle.batch_flat(metric, q, flat)
    if metric == DOT:
        want = -want  # oracle reports the distance (negated); simd.DotProductBatch is raw
    assert np.array_equal(out, want)


# ------------------------------------------------------------------ re-rank (HNSW candidate lists)
@pytest.mark.parametrize("dtype,metric", [(np.float32, L2), (np.float16, L2), (np.float32, COS)])
def test_rerank_parity(lbgpu, oracle, dtype, metric):
    rng = np.random.default_rng(5001)
    n, dim, nq, c, k = 20000, 96, 64, 128, 10
    db, q = make_db(rng, n, dim, dtype), make_db(rng, nq, dim, dtype)
    cand = rng.integers(0, n, (nq, c)).astype(np.uint32)
    cand[0, :5] = [n, n + 7, 0xFFFFFFFF, 3, 3]  # location misses + a duplicate
    tomb, allow = random_bitmap(rng, n, 0.05), random_bitmap(rng, n, 0.30)
    idx = lbgpu.DenseIndex(dim, dtype, metric)
    idx.add(db)
    gd, gl = idx.rerank(q, cand, k)
    c64 = cand.astype(np.int64)
    c64[c64 >= n] = -1
    wd, wl = oracle.rerank(metric, db, q, c64, k)
    assert_topk_equal(gd, gl, wd, wl, 0.0, "rerank")
    idx.set_tombstones(tomb)
    gd, gl = idx.rerank(q, cand, k, allow=allow)
    wd, wl = oracle.rerank(metric, db, q, c64, k, tomb=lbgpu.pack_bitmap(tomb), allow=lbgpu.pack_bitmap(allow))
    assert_topk_equal(gd, gl, wd, wl, 0.0, "rerank+bitmaps")
    from longbow_b200 import store
    rr = store.RerankBatch(idx, q[1], cand[1], 5)  # tombstones still set, no predicate
    wd1, wl1 = oracle.rerank(metric, db, q[1:2], c64[1:2], 5, tomb=lbgpu.pack_bitmap(tomb))
    assert [r.ID for r in rr] == [int(x) for x in wl1[0] if x >= 0]
    assert [np.float32(r.Distance) for r in rr] == [x for x, i in zip(wd1[0], wl1[0]) if i >= 0]
    idx.close()


# ------------------------------------------------------------------ PQ
def _pq_setup(rng, n, M, sub):
    cb = rng.standard_normal((M, 256, sub)).astype(np.float32)
    codes = rng.integers(0, 256, (n, M), dtype=np.uint8)
    return cb, codes


@pytest.mark.parametrize("M,sub", [(96, 8), (8, 4), (5, 3), (32, 16)])
def test_adc_table_and_batch_bit_exact(oracle, M, sub):
    from longbow_b200 import pq, simd
    rng = np.random.default_rng(3000 + M)
    cb, codes = _pq_setup(rng, 1000, M, sub)
    enc = pq.PQEncoder(M * sub, M, 256, cb)
    q = rng.standard_normal(M * sub).astype(np.float32)
    table = enc.BuildADCTable(q)
    assert np.array_equal(table, oracle.adc_table(q, cb))
    out = np.empty(1000, np.float32)
    enc.ADCDistanceBatch(table, codes, out)
    assert np.array_equal(out, oracle.adc_batch(table, codes))
    out2 = np.empty(1000, np.float32)
    simd.ADCDistanceBatch(table, codes.reshape(-1), M, out2)
    assert np.array_equal(out2, out)
    vecs = rng.standard_normal((200, M * sub)).astype(np.float32)
    assert np.array_equal(enc.EncodeBatch(vecs), oracle.pq_encode(vecs, cb))
    blob = enc.Serialize()
    enc2 = pq.PQEncoder.Deserialize(blob)
    assert np.array_equal(enc2.BuildADCTable(q), table)
    enc.close(); enc2.close()


@pytest.mark.parametrize("M,sub,n", [(96, 8, 30000), (16, 4, 5000)])
def test_pq_search_bit_exact(lbgpu, oracle, M, sub, n):
    from longbow_b200 import pq
    rng = np.random.default_rng(3100 + M)
    cb, codes = _pq_setup(rng, n, M, sub)
    dim = M * sub
    raw = oracle.pq_decode(codes, cb) + rng.normal(0, 0.05, (n, dim)).astype(np.float32)
    q = rng.standard_normal((9, dim)).astype(np.float32)
    enc = pq.PQEncoder(dim, M, 256, cb)
    enc.add_codes(codes[: n // 2]); enc.add_codes(codes[n // 2:])
    # ADC-only top-k
    gd, gl = enc.search(q, 10)
    wd, wl = oracle.pq_search(cb, codes, None, q, 10, 0)
    assert_topk_equal(gd, gl, wd, wl, 0.0, "adc top-k")
    # with fp32 re-rank of k' = 100 candidates
    rawidx = lbgpu.DenseIndex(dim, np.float32, L2)
    rawidx.add(raw)
    enc.attach_raw(rawidx)
    gd, gl = enc.search(q, 10, 100)
    wd, wl = oracle.pq_search(cb, codes, raw, q, 10, 100)
    assert_topk_equal(gd, gl, wd, wl, 0.0, "adc + rerank")
    # bitmaps
    tomb, allow = random_bitmap(rng, n, 0.05), random_bitmap(rng, n, 0.3)
    enc.set_tombstones(tomb)
    gd, gl = enc.search(q, 10, 100, allow=allow)
    wd, wl = oracle.pq_search(cb, codes, raw, q, 10, 100, tomb=lbgpu.pack_bitmap(tomb), allow=lbgpu.pack_bitmap(allow))
    assert_topk_equal(gd, gl, wd, wl, 0.0, "adc + rerank + bitmaps")
    enc.close(); rawidx.close()


# ------------------------------------------------------------------ merge / select / filters
def test_merge_and_select(oracle):
    from longbow_b200 import store
    rng = np.random.default_rng(4)
    parts, nq, k_in, k = 8, 50, 10, 10
    d = np.sort(rng.random((parts, nq, k_in), dtype=np.float32), axis=2)
    d[rng.random(d.shape) < 0.2] = 0.5  # many exact ties
    d = np.sort(d, axis=2)
    l = rng.permutation(parts * nq * k_in).reshape(parts, nq, k_in).astype(np.int64) + (1 << 33)
    l[3, :, 7:] = -1
    gd, gl = store.MergeShardResults(d, l, k)
    wd, wl = oracle.merge(d, l, k)
    assert_topk_equal(gd, gl, wd, wl, 0.0, "merge")
    x = rng.random(100000, dtype=np.float32)
    x[500] = x[7]
    gi, gd2 = store.SelectTopKNeighbors(x, 25)
    wd2, wi = oracle.select_k(x, 25)
    assert np.array_equal(gi, wi) and np.array_equal(gd2, wd2)


def test_filter_bitmaps(oracle):
    from longbow_b200 import store
    rng = np.random.default_rng(8)
    n = 100_003
    ci = rng.integers(-5, 5, n).astype(np.int64)
    cf = rng.standard_normal(n).astype(np.float32)
    for op, fn in enumerate([np.equal, np.not_equal, np.greater, np.greater_equal, np.less, np.less_equal]):
        bm = store.GenerateFilterBitset(ci, op, 1)
        want = np.packbits(np.pad(fn(ci, 1), (0, (-n) % 64)), bitorder="little").view(np.uint64)
        assert np.array_equal(bm, want), op
        bm2 = store.GenerateFilterBitset(cf, op, 0.25, bitmap=bm)
        want2 = np.packbits(np.pad(fn(ci, 1) & fn(cf, np.float32(0.25)), (0, (-n) % 64)), bitorder="little").view(np.uint64)
        assert np.array_equal(bm2, want2), op
    # internal/store/bitmap_filter_test.go:94-153: 3 rows, category == "B" -> only row 1 (dictionary code 1)
    cat = np.array([0, 1, 2], np.int64)
    assert store.GenerateFilterBitset(cat, 0, 1)[0] == 0b010


def test_brute_force_index_mirror(oracle):
    from longbow_b200 import store
    rng = np.random.default_rng(1001)
    db = rng.random((1500, 128), dtype=np.float32)
    q = rng.random(128, dtype=np.float32)
    bf = store.BruteForceIndex(128)
    assert bf.SearchVectors(q, 10) is None  # empty index -> nil, adaptive_index.go:172-174
    bf.AddBatch(db)
    res = bf.SearchVectors(q, 10)
    wd, wl = oracle.search(L2, db, q[None, :], 10)
    assert [r.ID for r in res] == list(wl[0]) and [np.float32(r.Score) for r in res] == list(wd[0])
    with pytest.raises(TypeError):
        bf.SearchVectors(q.astype(np.float64), 10)
    bf.Close()


# ------------------------------------------------------------------ tensor-core scan vs SIMT scan vs oracle
@pytest.fixture
def scan_mode():
    from longbow_b200 import _lib
    yield lambda m: _lib.set_option("dense_scan", m)
    _lib.set_option("dense_scan", 0)


@pytest.mark.parametrize("dtype,metric", [(np.float16, COS), (np.float16, L2), (np.float16, DOT), (np.int8, DOT), (np.int8, L2),
                                          (np.float32, L2), (np.float32, COS), (np.float32, DOT)])
@pytest.mark.parametrize("n,dim,nq,k", [(70001, 768, 300, 100), (40000, 64, 129, 10), (9000, 208, 40, 32), (300, 128, 5, 10), (20000, 128, 40, 300)])
def test_tensor_core_scan_parity(lbgpu, oracle, scan_mode, dtype, metric, n, dim, nq, k):
    rng = np.random.default_rng(2000 + n + metric)
    db, q = make_db(rng, n, dim, dtype), make_db(rng, nq, dim, dtype)
    if metric == COS:
        db[[3, n // 2, n - 1]] = 0
    idx = lbgpu.DenseIndex(dim, dtype, metric)
    idx.add(db)
    wd, wl = oracle.search(metric, db, q, k)
    for mode in (2, 1):  # forced tensor-core, forced SIMT
        scan_mode(mode)
        gd, gl = idx.search(q, k)
        assert_topk_equal(gd, gl, wd, wl, 0.0, f"mode={mode} {dtype.__name__} metric={metric}")
    # bitmaps through the tensor-core epilogue
    scan_mode(2)
    tomb, allow = random_bitmap(rng, n, 0.05), random_bitmap(rng, n, 0.30)
    idx.set_tombstones(tomb)
    gd, gl = idx.search(q[:50], k, allow=allow)
    wd, wl = oracle.search(metric, db, q[:50], k, tomb=lbgpu.pack_bitmap(tomb), allow=lbgpu.pack_bitmap(allow))
    assert_topk_equal(gd, gl, wd, wl, 0.0, "tc + bitmaps")
    idx.close()


def test_tensor_core_ineligible_is_loud(lbgpu, scan_mode):
    from longbow_b200 import LongbowError
    idx = lbgpu.DenseIndex(7, np.float16, COS)  # 14-byte rows: TMA pitch not a multiple of 16
    idx.add(np.ones((10, 7), np.float16))
    scan_mode(2)
    with pytest.raises(LongbowError):
        idx.search(np.ones((1, 7), np.float16), 3)
    scan_mode(0)
    d, l = idx.search(np.ones((1, 7), np.float16), 3)  # auto falls to the SIMT scan
    assert list(l[0]) == [0, 1, 2]
    idx.close()


def test_tensor_core_bootstrap_ties_and_toggle(lbgpu, oracle, scan_mode):
    # every vector appears 4 times, copies straddle the bootstrap sample / main scan boundary: the
    # (distance, id) order must survive the bootstrap threshold (it admits ties: nextafter)
    from longbow_b200 import _lib
    rng = np.random.default_rng(31)
    base = make_db(rng, 5000, 128, np.float16)
    db = np.concatenate([base, base, base, base])
    q = base[rng.integers(0, 5000, 150)]
    idx = lbgpu.DenseIndex(128, np.float16, COS)
    idx.add(db)
    wd, wl = oracle.search(COS, db, q, 10)
    scan_mode(2)
    try:
        for boot in (1, 0):
            _lib.set_option("tc_boot", boot)
            gd, gl = idx.search(q, 10)
            assert_topk_equal(gd, gl, wd, wl, 0.0, f"boot={boot}")
    finally:
        _lib.set_option("tc_boot", 1)
    idx.close()


# ------------------------------------------------------------------ streaming scan (small query batches)
@pytest.mark.parametrize("dtype,metric", [(np.float16, COS), (np.float16, L2), (np.float32, L2), (np.float32, DOT),
                                          (np.float32, COS), (np.int8, DOT), (np.int8, L2)])
@pytest.mark.parametrize("n,dim,k", [(70001, 768, 100), (40000, 128, 10), (300, 128, 5), (9000, 256, 32), (131072, 64, 10)])
@pytest.mark.parametrize("nq", [1, 3, 8])
def test_streaming_scan_parity(lbgpu, oracle, scan_mode, dtype, metric, n, dim, k, nq):
    if (dim * np.dtype(dtype).itemsize) % 128 != 0:
        pytest.skip("row bytes not a multiple of 128: not eligible for the streaming scan")
    rng = np.random.default_rng(3000 + n + metric + nq)
    db, q = make_db(rng, n, dim, dtype), make_db(rng, nq, dim, dtype)
    if metric == COS:
        db[[3, n // 2, n - 1]] = 0
    if n > 1000:
        db[n // 3] = db[5]  # exact duplicate rows across the bootstrap sample boundary: ties by id
        q[0] = db[5]
    idx = lbgpu.DenseIndex(dim, dtype, metric)
    idx.add(db)
    wd, wl = oracle.search(metric, db, q, k)
    scan_mode(3)
    gd, gl = idx.search(q, k)
    assert_topk_equal(gd, gl, wd, wl, 0.0, f"stream {dtype.__name__} metric={metric} nq={nq}")
    scan_mode(0)  # auto must pick an exact path too
    gd, gl = idx.search(q, k)
    assert_topk_equal(gd, gl, wd, wl, 0.0, "auto")
    tomb, allow = random_bitmap(rng, n, 0.05), random_bitmap(rng, n, 0.30)
    idx.set_tombstones(tomb)
    scan_mode(3)
    gd, gl = idx.search(q, k, allow=allow)
    wd, wl = oracle.search(metric, db, q, k, tomb=lbgpu.pack_bitmap(tomb), allow=lbgpu.pack_bitmap(allow))
    assert_topk_equal(gd, gl, wd, wl, 0.0, "stream + bitmaps")
    idx.close()


def test_streaming_scan_adversarial_order(lbgpu, oracle, scan_mode):
    """Rows sorted from worst to best: every row beats the running threshold, so the lists overflow and the
    compaction path runs on every trip."""
    rng = np.random.default_rng(77)
    n, dim = 60000, 128
    q = rng.random((2, dim), dtype=np.float32)
    db = rng.random((n, dim), dtype=np.float32)
    order = np.argsort(-np.linalg.norm(db - q[0], axis=1))
    db = np.ascontiguousarray(db[order])
    idx = lbgpu.DenseIndex(dim, np.float32, L2)
    idx.add(db)
    scan_mode(3)
    from longbow_b200 import _lib
    for boot in (1, 0):
        _lib.set_option("tc_boot", boot)
        gd, gl = idx.search(q, 50)
        wd, wl = oracle.search(L2, db, q, 50)
        assert_topk_equal(gd, gl, wd, wl, 0.0, f"adversarial boot={boot}")
    _lib.set_option("tc_boot", 1)
    idx.close()


# ------------------------------------------------------------------ coarse-key accuracy of the tensor-core scan
@pytest.mark.parametrize("dtype,metric,bound", [
    # 3xTF32: operand error ~2^-21; what remains (~4e-6 on the all-positive test data) is the tensor core's
    # truncating fp32 accumulation over 3 * dim/8 steps.  Plain TF32 would sit at 1e-4 .. 1e-3.
    (np.float32, L2, 1.2e-5), (np.float32, COS, 1.2e-5), (np.float32, DOT, 1.2e-5),
    (np.float16, COS, 4e-6), (np.float16, L2, 4e-6), (np.int8, DOT, 1e-12), (np.int8, L2, 1e-12)])  # int8: exact
def test_coarse_keys_accuracy(lbgpu, dtype, metric, bound):
    """The coarse keys only rank candidates, but the candidate margin (kc - k) assumes they are accurate to a
    few fp32 ulps of |q||x|.  For fp32 rows this pins the 3xTF32 split: a plain TF32 product (or a hardware
    conversion that rounds instead of truncating the hi part) would miss the bound by two orders of magnitude."""
    rng = np.random.default_rng(5)
    n, dim, nq = 2048, 256, 16
    db, q = make_db(rng, n, dim, dtype), make_db(rng, nq, dim, dtype)
    idx = lbgpu.DenseIndex(dim, dtype, metric)
    idx.add(db)
    keys = idx.coarse_keys(q, n).astype(np.float64)
    d64, q64 = db.astype(np.float64), q.astype(np.float64)
    dots = q64 @ d64.T
    xn = np.linalg.norm(d64, axis=1)
    if metric == L2:
        want = (xn ** 2)[None, :] - 2.0 * dots
    elif metric == COS:
        want = -dots / np.maximum(xn, 1e-300)[None, :]
    else:
        want = -dots
    scale = np.linalg.norm(q64, axis=1)[:, None] * np.maximum(xn, 1e-30)[None, :]
    if metric == COS:
        scale = np.linalg.norm(q64, axis=1)[:, None] * np.ones_like(xn)[None, :]
    if metric == L2:
        scale = scale + (xn ** 2)[None, :]
    err = np.abs(keys - want) / scale
    assert err.max() <= bound, f"coarse key error {err.max():.3e} of |q||x| (bound {bound})"
    idx.close()


# ------------------------------------------------------------------ certification of the coarse stage
def test_certification_near_ties(lbgpu, oracle):
    """300 rows that differ from each other by one fp16 ulp in one coordinate: their exact distances to the
    query differ by less than the coarse error, and there are more of them than the candidate margin (kc - k).
    The host search must notice (certification), redo those queries exhaustively, and still match the oracle."""
    from longbow_b200 import _lib
    rng = np.random.default_rng(99)
    n, dim, k = 30000, 256, 100
    db = make_db(rng, n, dim, np.float16)
    base = db[17].copy()
    pos = rng.choice(np.arange(1000, n), 300, replace=False)
    for j, r in enumerate(pos):
        v = base.copy()
        toward = np.float16(np.inf) if j % 2 else np.float16(-np.inf)
        v[j % dim] = np.nextafter(v[j % dim], toward, dtype=np.float16)
        db[r] = v
    q = np.stack([base, db[5], db[12345]])  # query 0 sits in the middle of the near-tie cluster
    for metric in (COS, L2, DOT):
        idx = lbgpu.DenseIndex(dim, np.float16, metric)
        idx.add(db)
        gd, gl = idx.search(q, k)
        wd, wl = oracle.search(metric, db, q, k)
        assert_topk_equal(gd, gl, wd, wl, 0.0, f"near ties, metric {metric}")
        assert idx.last_uncertified() >= 1, "the near-tie query must be flagged"
        # the same through the device entry point is NOT certified; with certification off the host call may differ
        idx.close()
    # ordinary data: nothing is flagged
    db2, q2 = make_db(rng, 20000, 128, np.float32), make_db(rng, 50, 128, np.float32)
    idx = lbgpu.DenseIndex(128, np.float32, L2)
    idx.add(db2)
    gd, gl = idx.search(q2, 10)
    wd, wl = oracle.search(L2, db2, q2, 10)
    assert_topk_equal(gd, gl, wd, wl, 0.0, "plain fp32")
    assert idx.last_uncertified() == 0
    _lib.set_option("certify", 0)
    try:
        idx.search(q2, 10)
        assert idx.last_uncertified() == 0
    finally:
        _lib.set_option("certify", 1)
    idx.close()


# ------------------------------------------------------------------ PQ training (k-means) on the GPU
@pytest.mark.parametrize("n,dims,M,K", [(4000, 32, 4, 16), (6000, 64, 8, 256), (700, 12, 4, 32), (3000, 40, 2, 64)])
def test_pq_train_bit_exact(oracle, n, dims, M, K):
    """lb_pq_train == TrainKMeans (internal/pq/kmeans.go:64-151) restated in the oracle, bit for bit, from the
    same initial rows -- including the number of iterations each subspace runs before the early stop."""
    from longbow_b200 import pq
    rng = np.random.default_rng(n + K)
    centers = rng.standard_normal((K // 2, dims)).astype(np.float32) * 3
    data = (centers[rng.integers(0, K // 2, n)] + rng.standard_normal((n, dims)).astype(np.float32)).astype(np.float32)
    data[10] = data[11]  # duplicate rows: equal distances, the first centroid index must win
    init = np.stack([rng.permutation(n)[:K] for _ in range(M)]).astype(np.int32)
    cb, iters = pq.train_codebooks(dims, M, K, data, init_idx=init)
    wcb, witers = oracle.pq_train(data, M, K, init)
    assert np.array_equal(iters, witers), (iters, witers)
    assert np.array_equal(cb, wcb)
    if K == 256:  # resident handles need K = 256 (simd.ADCDistanceBatch hard-codes the table stride)
        enc = pq.PQEncoder.Train(dims, M, K, data, init_idx=init)
        codes = enc.EncodeBatch(data[:50])
        assert np.array_equal(codes, oracle.pq_encode(data[:50], wcb))
        enc.close()


def test_pq_train_validation():
    from longbow_b200 import pq
    with pytest.raises(ValueError):
        pq.train_codebooks(10, 3, 4, np.zeros((8, 10), np.float32))       # dims % M
    with pytest.raises(ValueError):
        pq.train_codebooks(8, 2, 16, np.zeros((4, 8), np.float32))        # n < k (kmeans.go:65-67)


# ------------------------------------------------------------------ SearchHybrid mirror, device-resident predicates
def test_search_hybrid_and_device_filters(lbgpu, oracle):
    import torch
    from longbow_b200 import store
    rng = np.random.default_rng(8)
    n, dim, k = 3000, 128, 10
    db = rng.random((n, dim), dtype=np.float32)
    q = rng.random(dim, dtype=np.float32)
    g = lbgpu.NewIndexWithConfig(lbgpu.GPUConfig(DeviceID=0, Dimension=dim))
    g.Add(list(range(n)), db)
    wd, wl = oracle.search(L2, db, q[None, :], k * 10)
    loc = np.zeros(n, np.int64)
    loc[wl[0][:3]] = -1  # the three nearest are tombstoned at the location store
    res = store.SearchHybrid(g, q, k, n, locations=loc)
    want = [int(i) for i in wl[0] if loc[i] != -1][:k]
    assert [r.ID for r in res] == want
    assert np.array_equal(np.array([r.Score for r in res], np.float32), wd[0][3:3 + k])
    g.Close()
    # predicates evaluated on the device feed search_device without visiting the host
    dev = torch.device("cuda", 0)
    price = torch.from_numpy(rng.random(n).astype(np.float32)).to(dev)
    cat = torch.from_numpy(rng.integers(0, 5, n)).to(dev)
    bm = store.GenerateFilterBitsetDevice(price, 4, 0.5)            # price < 0.5
    bm = store.GenerateFilterBitsetDevice(cat, 0, 2, d_bitmap=bm)   # AND category == 2
    mask = (price.cpu().numpy() < 0.5) & (cat.cpu().numpy() == 2)
    assert np.array_equal(bm.cpu().numpy().view(np.uint64), lbgpu.pack_bitmap(mask))
    idx = lbgpu.DenseIndex(dim, np.float32, L2)
    idx.add(db)
    od = torch.empty((1, k), dtype=torch.float32, device=dev)
    ol = torch.empty((1, k), dtype=torch.int64, device=dev)
    idx.search_device(torch.from_numpy(q[None, :]).to(dev), k, od, ol, allow=bm)
    torch.cuda.synchronize()
    wd, wl = oracle.search(L2, db, q[None, :], k, allow=lbgpu.pack_bitmap(mask))
    assert np.array_equal(ol.cpu().numpy(), wl) and np.array_equal(od.cpu().numpy(), wd)
    idx.close()


# ------------------------------------------------------------------ k beyond the fused selector
@pytest.mark.parametrize("dtype,metric", [(np.float32, L2), (np.float16, COS), (np.int8, DOT)])
def test_large_k_exact_path(lbgpu, oracle, dtype, metric):
    """SearchHybrid asks the GPU index for k * 10 candidates (internal/store/hnsw_gpu.go:85): k = 100 means 1000
    neighbours, beyond the fused selector (k <= 704).  Those searches take the exhaustive exact kernel."""
    rng = np.random.default_rng(123)
    n, dim, nq, k = 6000, 64, 3, 1000
    db, q = make_db(rng, n, dim, dtype), make_db(rng, nq, dim, dtype)
    idx = lbgpu.DenseIndex(dim, dtype, metric)
    idx.add(db)
    tomb = random_bitmap(rng, n, 0.1)
    idx.set_tombstones(tomb)
    gd, gl = idx.search(q, k)
    wd, wl = oracle.search(metric, db, q, k, tomb=lbgpu.pack_bitmap(tomb))
    assert_topk_equal(gd, gl, wd, wl, 0.0, "k = 1000")
    gd, gl = idx.search(q[:1], 2048)
    wd, wl = oracle.search(metric, db, q[:1], 2048, tomb=lbgpu.pack_bitmap(tomb))
    assert_topk_equal(gd, gl, wd, wl, 0.0, "k = 2048")
    idx.close()
    if dtype == np.float32:
        from longbow_b200 import store
        g = lbgpu.NewIndexWithConfig(lbgpu.GPUConfig(DeviceID=0, Dimension=dim))
        g.Add(list(range(n)), db)
        res = store.SearchHybrid(g, q[0], 100, n)          # 1000 candidates through faiss_gpu_index_search
        wd, wl = oracle.search(L2, db, q[:1], 100)
        assert [r.ID for r in res] == [int(i) for i in wl[0]]
        g.Close()


def test_single_query_large_k_plans_agree(lbgpu, oracle):
    """One query with k >= 256 takes the exhaustive exact chain by default (measured cheaper than the padded
    tensor-core block); lb_set_option("exhaustive_k", ...) moves the switch.  Both plans give the oracle's answer."""
    from longbow_b200 import _lib
    rng = np.random.default_rng(77)
    n, dim, k = 60013, 128, 500   # 15 chunks of 4096 rows: the two-level select with one empty chunk list
    db, q = make_db(rng, n, dim, np.float16), make_db(rng, 2, dim, np.float16)
    idx = lbgpu.DenseIndex(dim, np.float16, COS)
    idx.add(db)
    allow = random_bitmap(rng, n, 0.5)
    wd, wl = oracle.search(COS, db, q, k, allow=lbgpu.pack_bitmap(allow))
    counts = []
    try:
        for xk in (256, 100000, 1):
            _lib.set_option("exhaustive_k", xk)
            l0 = _lib.launch_count()
            gd, gl = idx.search(q[:1], k, allow=allow)
            counts.append(_lib.launch_count() - l0)
            assert_topk_equal(gd, gl, wd[:1], wl[:1], 0.0, f"one query, exhaustive_k={xk}")
            gd, gl = idx.search(q, k, allow=allow)      # two queries: always the ordinary plan
            assert_topk_equal(gd, gl, wd, wl, 0.0, f"two queries, exhaustive_k={xk}")
    finally:
        _lib.set_option("exhaustive_k", 256)
    assert counts[0] == counts[2], counts   # the default plan is the forced exhaustive chain's kernel sequence
    idx.close()


# ------------------------------------------------------------------ short rows on the tensor-core path
@pytest.mark.parametrize("dtype,dims", [(np.float32, (4, 8, 12, 16, 24)), (np.float16, (8, 16, 24, 40)), (np.int8, (16, 32, 48))])
def test_tensor_core_short_rows(lbgpu, oracle, scan_mode, dtype, dims):
    """Rows shorter than one 128-byte k-block (TMA boxes reach past the row: zero fill) on the forced and the
    automatic path, several query blocks."""
    rng = np.random.default_rng(321)
    for dim in dims:
        n, nq, k = 5000, 260, 10
        db, q = make_db(rng, n, dim, dtype), make_db(rng, nq, dim, dtype)
        for metric in ((L2, DOT) if dtype == np.int8 else (L2, COS, DOT)):
            idx = lbgpu.DenseIndex(dim, dtype, metric)
            idx.add(db)
            wd, wl = oracle.search(metric, db, q, k)
            for mode in (2, 0):
                scan_mode(mode)
                gd, gl = idx.search(q, k)
                assert_topk_equal(gd, gl, wd, wl, 0.0, f"dim {dim} {dtype.__name__} metric {metric} mode {mode}")
            idx.close()


# ------------------------------------------------------------------ growth, query chunking
def test_grow_after_search_and_query_chunks(lbgpu, oracle):
    """fp32 index: the 3xTF32 low parts are built by the first search and must follow later adds / reserves;
    more than 4096 queries are processed in chunks (certification flags included)."""
    rng = np.random.default_rng(55)
    dim, k = 64, 5
    db = rng.random((9000, dim), dtype=np.float32)
    q = rng.random((5000, dim), dtype=np.float32)
    idx = lbgpu.DenseIndex(dim, np.float32, L2)
    idx.add(db[:3000])
    gd, gl = idx.search(q[:100], k)
    wd, wl = oracle.search(L2, db[:3000], q[:100], k)
    assert_topk_equal(gd, gl, wd, wl, 0.0, "first third")
    idx.add(db[3000:6000])          # grows the mirror (and the low parts on the next search)
    idx.reserve(9000)
    idx.add(db[6000:])
    assert len(idx) == 9000
    gd, gl = idx.search(q, k)       # 5000 queries: two chunks
    wd, wl = oracle.search(L2, db, q, k)
    assert_topk_equal(gd, gl, wd, wl, 0.0, "all rows, 5000 queries")
    assert idx.last_uncertified() == 0
    idx.close()


@pytest.mark.parametrize("dims,M,K,n", [(32, 4, 16, 100), (268, 4, 16, 34), (91, 91, 1, 2), (64, 8, 256, 600)])
def test_pq_train_reference_fuzz_shapes(oracle, dims, M, K, n):
    """Shapes and the deterministic data generator of the reference's fuzz harness and its seed corpus
    (internal/pq/fuzz_test.go:9-30, internal/pq/testdata/fuzz/FuzzPQEncoder_TrainAndEncode/*): odd sub-vector
    length 67, K = 1 with two samples, many empty clusters (the data is a ramp, so rows repeat)."""
    from longbow_b200 import pq
    data = np.array([[np.float32(i + j) / np.float32(n) for j in range(dims)] for i in range(n)], np.float32)
    rng = np.random.default_rng(dims + n)
    init = np.stack([rng.permutation(n)[:K] for _ in range(M)]).astype(np.int32)
    cb, iters = pq.train_codebooks(dims, M, K, data, init_idx=init)
    wcb, witers = oracle.pq_train(data, M, K, init)
    assert np.array_equal(iters, witers)
    assert np.array_equal(cb, wcb)
    assert cb.shape == (M, K, dims // M)


# ------------------------------------------------------------------ Arrow ingest, growth, bitmap updates
def test_add_arrow_offsets_pinned_and_growth(lbgpu, oracle):
    """lb_index_add_arrow: list offset, the reference's truncated-buffer rule (arrow_utils.go:146-160), pinned chunked
    upload; the mirror grows across many appends without copying (base pointer stable), results stay bit-exact."""
    rng = np.random.default_rng(2024)
    n, dim, k = 30000, 96, 10
    db = make_db(rng, n, dim, np.float16)
    idx = lbgpu.DenseIndex(dim, np.float16, COS)
    pad = make_db(rng, 7, dim, np.float16)
    buf = np.concatenate([pad, db[:10000]]).reshape(-1)          # list offset 7 inside a larger values buffer
    idx.add_arrow(buf, 7, 10000, pin=True)
    idx.add_arrow(db[10000:20000].reshape(-1), 123, 10000, pin=True)   # IPC-flattened: offset kept, buffer relative
    idx.add_arrow(db[20000:].reshape(-1), 0, 10000, pin=False)
    with pytest.raises(Exception):
        idx.add_arrow(db[:10].reshape(-1), 0, 11)                  # too small even for relative access
    assert len(idx) == n
    q = make_db(rng, 20, dim, np.float16)
    gd, gl = idx.search(q, k)
    wd, wl = oracle.search(COS, db, q, k)
    assert_topk_equal(gd, gl, wd, wl, 0.0, "arrow ingest")
    # many small appends: growth maps memory behind the mirror, earlier rows stay where they were
    more = make_db(rng, 5000, dim, np.float16)
    for lo in range(0, 5000, 500):
        idx.add(more[lo:lo + 500])
    full = np.concatenate([db, more])
    gd, gl = idx.search(q, k)
    wd, wl = oracle.search(COS, full, q, k)
    assert_topk_equal(gd, gl, wd, wl, 0.0, "after growth")
    # tombstone updates in a loop (no device stall per update, old buffers retired)
    for i in range(12):
        tomb = np.zeros(len(idx), bool)
        tomb[rng.integers(0, len(idx), 200)] = True
        idx.set_tombstones(tomb)
    gd, gl = idx.search(q, k)
    wd, wl = oracle.search(COS, full, q, k, tomb=lbgpu.pack_bitmap(tomb))
    assert_topk_equal(gd, gl, wd, wl, 0.0, "after tombstone updates")
    idx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,metric,n,dim,k", [(np.int8, DOT, 300000, 128, 10), (np.float16, COS, 120000, 256, 100),
                                                  (np.float32, L2, 150000, 64, 10)])
def test_self_match_queries_and_duplicates(lbgpu, oracle, scan_mode, dtype, metric, n, dim, k):
    """Queries that ARE rows of the index (some of them inside the bootstrap sample, some stored several times): the
    sample then holds an outlier far below the rest of its keys.  The threshold ladder reads its extrapolation slope
    off mid ranks, sorts its 16 edges and seeds every bucket with the sample rows that fall into it, so the answers
    stay the oracle's -- tensor-core scan (batch) and streaming scan (one query)."""
    rng = np.random.default_rng(77 + dim)
    db = make_db(rng, n, dim, dtype)
    rows = np.concatenate([np.arange(0, 40), rng.integers(0, n, 88)])   # first 40: inside the sample
    q = db[rows].copy()
    db[n // 2: n // 2 + 5] = db[7]       # a query stored six times
    db[100:103] = db[n - 1]              # duplicates of a far row inside the sample
    idx = lbgpu.DenseIndex(dim, dtype, metric)
    idx.add(db)
    wd, wl = oracle.search(metric, db, q, k)
    scan_mode(2)
    gd, gl = idx.search(q, k)
    assert_topk_equal(gd, gl, wd, wl, 0.0, "self-match batch")
    scan_mode(0)
    for i in (0, 7, 50):
        gd1, gl1 = idx.search(q[i:i + 1], k)
        assert_topk_equal(gd1, gl1, wd[i:i + 1], wl[i:i + 1], 0.0, f"self-match single query {i}")
    idx.close()
