"""The deterministic inputs the reference's own tests build WITHOUT math/rand, rebuilt here bit for bit, with the
answers that follow from their construction (no oracle needed to know them) -- a wider pin for the oracle than the
<= 8-element known-answer tests, and the same fixtures through the GPU library.

* createTestDataset / makeTestVector (internal/store/hnsw_batch_test.go:18-36, 239-245):
  v[i][j] = float32((i*dims + j) % 100) / 100.  With dims = 16 the generator has period 25 in i (16 * 25 = 400 = 0 mod
  100), so rows s, s+25, s+50, s+75 are IDENTICAL and equal to makeTestVector(16, s): the exact top-4 of query s is
  that tie group at distance exactly 0, ordered by id -- the (distance, id) tie rule on the reference's own data.
* the RerankBatch test (hnsw_batch_test.go:96-125): candidates 0..9, k = 3, ascending distances.
* the PQ fuzz harness generator (internal/pq/fuzz_test.go:21-29): vec[j] = float32(i + j) / float32(numSamples).
* the GPU smoke ramp (internal/gpu/gpu_test.go:25-46) is covered in test_golden.py / test_gpu_parity.py.
"""
import numpy as np
import pytest

L2, COS, DOT = 0, 1, 2


def create_test_dataset(dims, n):
    i = np.arange(n, dtype=np.int64)[:, None]
    j = np.arange(dims, dtype=np.int64)[None, :]
    return (((i * dims + j) % 100).astype(np.float32) / np.float32(100.0)).astype(np.float32)


def make_test_vector(dims, seed):
    return (((seed * dims + np.arange(dims)) % 100).astype(np.float32) / np.float32(100.0)).astype(np.float32)


def _expected_tie_groups(seed, n):
    return [s for s in range(seed % 25, n, 25)]


def test_fixture_construction():
    db = create_test_dataset(16, 100)
    for s in range(3):
        assert np.array_equal(db[s], make_test_vector(16, s))
        for t in _expected_tie_groups(s, 100):
            assert np.array_equal(db[t], db[s])


def test_oracle_on_reference_dataset(oracle):
    db = create_test_dataset(16, 100)
    qs = np.stack([make_test_vector(16, s) for s in range(3)])
    d, l = oracle.search(L2, db, qs, 5)                       # TestHNSWIndex_SearchBatch: k = 5
    for s in range(3):
        assert l[s, :4].tolist() == _expected_tie_groups(s, 100), l[s]
        assert (d[s, :4] == 0.0).all() and d[s, 4] > 0.0
        assert (np.diff(d[s]) >= 0).all()
    # TestHNSWIndex_RerankBatch: 50 rows, query seed 0, candidates 0..9, k = 3
    db50 = create_test_dataset(16, 50)
    rd, rl = oracle.rerank(L2, db50, qs[:1], np.arange(10, dtype=np.int64)[None, :], 3)
    assert rl[0, 0] == 0 and rd[0, 0] == 0.0 and (np.diff(rd[0]) >= 0).all() and (rl[0] >= 0).all()
    # cosine on the same rows: identical rows -> 1 - x/x exactly 0 in the reference's arithmetic
    dc, lc = oracle.search(COS, db, qs, 4)
    for s in range(3):
        assert sorted(lc[s].tolist()) == _expected_tie_groups(s, 100)
        assert np.abs(dc[s]).max() <= 1e-6


def test_oracle_on_fuzz_generator_data(oracle):
    """fuzz_test.go seed (32, 4, 16, 100): ADC == L2^2(q, decode(code)) on its data (internal/pq/adc_test.go:53-65
    relation, 1e-4), codes are bytes < K, decode has `dims` entries."""
    dims, M, K, n = 32, 4, 16, 100
    i = np.arange(n, dtype=np.float32)[:, None]
    j = np.arange(dims, dtype=np.float32)[None, :]
    data = ((i + j) / np.float32(n)).astype(np.float32)
    init = np.stack([np.arange(K, dtype=np.int32) * (n // K) for _ in range(M)])
    cb, iters = oracle.pq_train(data, M, K, init)
    assert cb.shape == (M, K, dims // M) and (iters >= 1).all()
    codes = oracle.pq_encode(data, cb)
    assert codes.shape == (n, M) and codes.max() < K
    dec = oracle.pq_decode(codes, cb)
    assert dec.shape == (n, dims)
    q = data[7]
    # K = 16 < 256: the table stride of simd.ADCDistanceBatch is hard-coded 256 (simd.go:350), so the coherent
    # relation is the single-code form with stride K (adc_table.go:77-92): un-sqrt'd sum == L2^2(q, decode(code))
    table = np.concatenate([((cb[m] - q[m * 8:(m + 1) * 8]) ** 2).sum(1) for m in range(M)]).astype(np.float32)
    for r in (0, 7, 99):
        got = oracle.adc_single(table, codes[r], K)
        want = float(((dec[r] - q) ** 2).sum())
        assert abs(got - want) <= 1e-4 * max(1.0, want)


@pytest.mark.gpu
def test_gpu_on_reference_dataset(oracle):
    from longbow_b200 import gpu, store
    db = create_test_dataset(16, 100)
    qs = np.stack([make_test_vector(16, s) for s in range(3)])
    for metric in (L2, COS):
        idx = gpu.DenseIndex(16, np.float32, metric)
        idx.add(db)
        d, l = idx.search(qs, 5)
        wd, wl = oracle.search(metric, db, qs, 5)
        assert np.array_equal(l, wl) and np.array_equal(d, wd)
        if metric == L2:
            for s in range(3):
                assert l[s, :4].tolist() == _expected_tie_groups(s, 100) and (d[s, :4] == 0.0).all()
        idx.close()
    idx = gpu.DenseIndex(16, np.float32, L2)
    idx.add(create_test_dataset(16, 50))
    res = store.RerankBatch(idx, qs[0], list(range(10)), 3)     # hnsw_batch_test.go:96-125
    assert len(res) == 3 and res[0].ID == 0 and res[0].Distance == 0.0
    assert all(a.Distance <= b.Distance for a, b in zip(res, res[1:]))
    assert store.RerankBatch(idx, qs[0], [], 3) is None          # TestHNSWIndex_RerankBatch_Empty
    idx.close()
    bf = store.BruteForceIndex(16)
    bf.AddBatch(db)
    out = bf.SearchBatch(qs, 5)
    assert len(out) == 3 and all(0 < len(r) <= 5 for r in out)   # TestHNSWIndex_SearchBatch's own assertions
    bf.Close()
