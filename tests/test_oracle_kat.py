"""Pin the CPU oracle against the reference's own known-answer tests.

Each case cites the reference test it restates.  The reference has no golden
files for this path; these KATs plus the scalar float64 relations below are the
pins.
"""
import math

import numpy as np
import pytest

f32 = np.float32


def test_euclid_dot_kat(oracle):
    # internal/simd/simd_dispatch_test.go:56-78
    a = np.array([1, 2, 3, 4], f32)
    b = np.array([5, 6, 7, 8], f32)
    assert oracle.raw("lbo_euclid_f32", a, b) == 8.0
    assert oracle.raw("lbo_dot_f32", a, b) == 70.0
    c = oracle.raw("lbo_cosine_f32", a, b)
    assert 0 <= c <= 2  # :80-87


def test_batch_kat(oracle):
    # internal/simd/simd_dispatch_test.go:89-105
    q = np.array([1, 2, 3, 4], f32)
    vs = np.array([[5, 6, 7, 8], [1, 2, 3, 4]], f32)
    r = oracle.batch_flat(oracle.L2, q, vs)
    assert r[0] == 8.0 and r[1] == 0.0


def test_128d_unit_diff(oracle):
    # internal/simd/simd_dispatch_test.go:115-123
    a = np.zeros(128, f32)
    b = np.zeros(128, f32)
    a[0], b[0] = 1.0, 2.0
    assert oracle.raw("lbo_euclid_f32", a, b) == 1.0


def test_int8_kat(oracle):
    # internal/simd/simd_dispatch_test.go:125-132
    a = np.array([10, 20], np.int8)
    b = np.array([10, 30], np.int8)
    assert oracle.raw("lbo_euclid_i8", a, b) == 10.0


def test_cosine_kats(oracle):
    # internal/simd/simd_test.go:146-194
    a = np.array([1, 2, 3, 4, 5, 6, 7, 8], f32)
    assert abs(oracle.raw("lbo_cosine_f32", a, a)) <= 1e-5
    assert abs(oracle.raw("lbo_cosine_f32", np.array([1, 0, 0, 0], f32), np.array([0, 1, 0, 0], f32)) - 1.0) <= 1e-5
    assert abs(oracle.raw("lbo_cosine_f32", np.array([1, 2, 3, 4], f32), np.array([-1, -2, -3, -4], f32)) - 2.0) <= 1e-5
    z = oracle.raw("lbo_cosine_f32", np.zeros(4, f32), np.array([1, 2, 3, 4], f32))
    assert z == 1.0  # exactly


def test_l2_distance_operator_kat(oracle):
    # internal/store/arrow_kernels_test.go:56-69: rows {1,2,3,4} vs query {1,2,3,4} -> 0, next -> sqrt(30)...
    q = np.array([1, 2, 3, 4], f32)
    rows = np.array([[1, 2, 3, 4], [2, 4, 6, 8]], f32)
    r = oracle.batch_flat(oracle.L2, q, rows)
    assert r[0] == 0.0
    assert abs(r[1] - math.sqrt(30.0)) <= 1e-6


def test_doc_examples(oracle):
    # docs/distance_metrics.md:21 (sqrt 27) and :55 (-11 as a distance)
    a = np.array([1, 2, 3], f32)
    b = np.array([4, 5, 6], f32)
    assert abs(oracle.raw("lbo_euclid_f32", a, b) - math.sqrt(27.0)) <= 1e-6
    x = np.array([1, 2], f32)
    y = np.array([3, 4], f32)
    assert oracle.distance(oracle.DOT, x, y) == -11.0


def test_gpu_smoke_kat(oracle):
    # internal/gpu/gpu_test.go:25-46: 10x128 ramp v[i]=0.01*i, query = row 0, k=5 -> ids[0]==0, d<0.01
    v = (np.arange(1280, dtype=np.float32) * f32(0.01)).reshape(10, 128)
    d, i = oracle.search(oracle.L2, v, v[:1], 5)
    assert i[0, 0] == 0 and d[0, 0] < 0.01
    assert list(i[0]) == [0, 1, 2, 3, 4]


def test_empty_vectors(oracle):
    # distance_functions.go:21-23, 51-53, 63-65
    e = np.zeros(0, f32)
    assert oracle.raw("lbo_euclid_f32", e, e) == 0.0
    assert oracle.raw("lbo_cosine_f32", e, e) == 1.0
    assert oracle.raw("lbo_dot_f32", e, e) == 0.0


DIMS = [1, 3, 7, 8, 15, 16, 31, 32, 64, 128, 256, 384, 512, 768, 1024, 1536]  # simd_test.go dims


@pytest.mark.parametrize("dim", DIMS)
def test_vs_float64_reference(oracle, dim):
    # internal/simd/simd_test.go:105-122,196-213,238-260: optimised vs scalar, 1e-5 (small) / 1e-3 (large)
    rng = np.random.default_rng(42 + dim)
    a = rng.random(dim, dtype=np.float32)
    b = rng.random(dim, dtype=np.float32)
    tol = 1e-5 if dim <= 128 else 1e-3
    a64, b64 = a.astype(np.float64), b.astype(np.float64)
    ref_l2 = math.sqrt(((a64 - b64) ** 2).sum())
    ref_dot = float((a64 * b64).sum())
    ref_cos = 1.0 - ref_dot / math.sqrt((a64 ** 2).sum() * (b64 ** 2).sum())
    assert abs(oracle.raw("lbo_euclid_f32", a, b) - ref_l2) <= tol * max(1.0, ref_l2)
    assert abs(oracle.raw("lbo_dot_f32", a, b) - ref_dot) <= tol * max(1.0, abs(ref_dot))
    assert abs(oracle.raw("lbo_cosine_f32", a, b) - ref_cos) <= tol
    # the reference's own single-accumulator variants agree within the same tolerance
    assert abs(oracle.raw("lbo_dot_f32_seq", a, b) - ref_dot) <= tol * max(1.0, abs(ref_dot))
    assert abs(oracle.raw("lbo_cosine_f32_seq", a, b) - ref_cos) <= tol
    # fast (AVX) baseline agrees with exact within the reference's SIMD tolerance
    for m in (oracle.L2, oracle.COSINE, oracle.DOT):
        e = oracle.distance(m, a, b)
        f = float(oracle.fast().lbf_distance(m, 0, a.ctypes.data, b.ctypes.data, dim))
        assert abs(e - f) <= tol * max(1.0, abs(e))


@pytest.mark.parametrize("dim", DIMS)
def test_f16_vs_float64(oracle, dim):
    # internal/simd/simd_f16_test.go:13-60 (float64 reference)
    rng = np.random.default_rng(123 + dim)
    a = rng.standard_normal(dim).astype(np.float16)
    b = rng.standard_normal(dim).astype(np.float16)
    a64, b64 = a.astype(np.float64), b.astype(np.float64)
    ref_l2 = math.sqrt(((a64 - b64) ** 2).sum())
    ref_dot = float((a64 * b64).sum())
    na, nb = (a64 ** 2).sum(), (b64 ** 2).sum()
    ref_cos = 1.0 - ref_dot / math.sqrt(na * nb) if na > 0 and nb > 0 else 1.0
    assert abs(oracle.raw("lbo_euclid_f16", a, b) - ref_l2) <= 1e-3 * max(1.0, ref_l2)
    assert abs(oracle.raw("lbo_dot_f16", a, b) - ref_dot) <= 1e-3 * max(1.0, abs(ref_dot))
    assert abs(oracle.raw("lbo_cosine_f16", a, b) - ref_cos) <= 1e-3
    for m in (oracle.L2, oracle.COSINE, oracle.DOT):
        e = oracle.distance(m, a, b)
        f = float(oracle.fast().lbf_distance(m, 1, a.ctypes.data, b.ctypes.data, dim))
        assert abs(e - f) <= 1e-3 * max(1.0, abs(e))


def test_h2f_exact(oracle):
    # every binary16 bit pattern widens exactly as numpy's IEEE conversion does
    bits = np.arange(65536, dtype=np.uint16)
    want = bits.view(np.float16).astype(np.float32)
    got = np.array([oracle.exact().lbo_h2f(int(b)) for b in bits], np.float32)
    nan = np.isnan(want)
    assert np.array_equal(got[~nan], want[~nan]) and np.isnan(got[nan]).all()


def test_int8_exact_integers(oracle):
    # for D<=128 every fp32 partial sum is an exact integer
    rng = np.random.default_rng(4001)
    a = rng.integers(-128, 128, 128, dtype=np.int8)
    b = rng.integers(-128, 128, 128, dtype=np.int8)
    dot = int((a.astype(np.int64) * b.astype(np.int64)).sum())
    l2 = int(((a.astype(np.int64) - b.astype(np.int64)) ** 2).sum())
    assert oracle.raw("lbo_dot_i8", a, b) == float(dot)
    assert oracle.raw("lbo_euclid_i8", a, b) == np.float32(math.sqrt(l2))
    for m, want in ((oracle.L2, np.float32(math.sqrt(l2))), (oracle.DOT, -float(dot))):
        f = float(oracle.fast().lbf_distance(m, 2, a.ctypes.data, b.ctypes.data, 128))
        assert f == want


def test_sq8(oracle):
    # internal/simd/sq8.go:45-66; batch == single exactly (simd_batch_test.go)
    rng = np.random.default_rng(7)
    a = rng.integers(0, 256, 100, dtype=np.uint8)
    b = rng.integers(0, 256, 100, dtype=np.uint8)
    want = int(((a.astype(np.int64) - b.astype(np.int64)) ** 2).sum())
    assert oracle.exact().lbo_l2sq_u8(a.ctypes.data, b.ctypes.data, 100) == want
    assert oracle.batch_flat(oracle.L2, a, b[None, :])[0] == float(want)


def test_topk_ties_by_id(oracle):
    # internal/store/adaptive_index.go:200-211: strict '<' keeps the lowest ids among equal distances
    db = np.zeros((8, 4), f32)
    db[5] = 1.0
    d, i = oracle.search(oracle.L2, db, np.zeros((1, 4), f32), 3)
    assert list(i[0]) == [0, 1, 2] and (d[0] == 0).all()
    d, i = oracle.search(oracle.L2, db[:2], np.zeros((1, 4), f32), 4)
    assert list(i[0]) == [0, 1, -1, -1] and d[0, 2] == np.finfo(np.float32).max


def test_adc_relations(oracle):
    # internal/pq/adc_test.go:53-65: ADC == L2^2(q, decode(code)) within 1e-4;
    # internal/store/pq_simd_test.go:11-53: batch vs scalar loop within 1e-5
    rng = np.random.default_rng(12345)
    M, K, sub = 4, 256, 8
    cb = rng.random((M, K, sub), dtype=np.float32)
    vecs = rng.random((64, M * sub), dtype=np.float32)
    q = rng.random(M * sub, dtype=np.float32)
    codes = oracle.pq_encode(vecs, cb)
    table = oracle.adc_table(q, cb)
    dec = oracle.pq_decode(codes, cb)
    batch = oracle.adc_batch(table, codes)
    fastb = oracle.adc_batch(table, codes, impl="fast")
    for r in range(64):
        manual = np.float32(0)
        for i in range(M * sub):
            d = q[i] - dec[r, i]
            manual += d * d
        single = oracle.adc_single(table, codes[r], K)
        assert abs(single - manual) <= 1e-4
        assert abs(batch[r] - math.sqrt(single)) <= 1e-5
        assert abs(fastb[r] - batch[r]) <= 1e-5
    # encode picks the nearest centroid: decoded vector is no farther than any other centroid choice
    for r in range(8):
        for m in range(M):
            dists = ((cb[m] - vecs[r, m * sub:(m + 1) * sub]) ** 2).sum(1)
            assert dists[codes[r, m]] <= dists.min() * (1 + 1e-5) + 1e-7


def test_merge_and_select(oracle):
    # internal/store/sharded_hnsw.go:494-503, result_merger.go:34-100, arrow_kernels.go:230-345
    d = np.array([[[0.5, 0.7, 0.9]], [[0.1, 0.7, 1.5]]], f32)
    i = np.array([[[3, 9, 4]], [[11, 2, -1]]], np.int64)
    od, oi = oracle.merge(d, i, 4)
    assert list(oi[0]) == [11, 3, 2, 9]
    assert list(od[0]) == [f32(0.1), f32(0.5), f32(0.7), f32(0.7)]
    sd, si = oracle.select_k(np.array([3, 1, 2, 1], f32), 3)
    assert list(si) == [1, 3, 2]


def test_fast_search_matches_exact_ids(oracle):
    rng = np.random.default_rng(1001)
    db = rng.random((2000, 128), dtype=np.float32)
    q = rng.random((16, 128), dtype=np.float32)
    de, ie = oracle.search(oracle.L2, db, q, 10)
    df, if_ = oracle.search(oracle.L2, db, q, 10, impl="fast")
    assert np.array_equal(ie, if_)
    assert np.allclose(de, df, rtol=1e-5)
