"""An independent numpy restatement of the reference's PQ training / encoding / ADC arithmetic, written from the Go
sources and not from oracle/lb_oracle.c, that the C oracle must agree with bit for bit (the same double-entry check
tests/test_hnsw.py makes for searchLayer).  CPU only.

* TrainKMeans            internal/pq/kmeans.go:64-151 (E-step: first strictly smaller L2Squared wins; sums in data
                         order; M-step: sum / float32(count); early stop iter > 0 && changed < n/1000 + 1)
* Encode                 internal/pq/encoder.go:76-136 + internal/simd/simd.go:278-326 (K <= 16: squared distances;
                         larger K: the sqrt'd batch distances, first strictly smaller wins)
* BuildADCTable / ADC    internal/pq/adc_table.go:15-51,77-92, internal/simd/simd.go:345-355
* L2SquaredFloat32       internal/simd/distance_functions.go:195-227 (four lanes, remainder into lane 0)

The initial centroids are passed in explicitly (rand.Perm is Go's generator: unpinned); the empty-
cluster re-seed never fires on these inputs (asserted).
"""
import math

import numpy as np
import pytest

f32 = np.float32


def l2sq_lanes(a, b):
    """a [..., d], b [..., d] float32 -> float32, the reference's lane order, vectorised over the leading axes."""
    d = (a - b).astype(f32)
    sq = (d * d).astype(f32)
    n = sq.shape[-1]
    lanes = [np.zeros(sq.shape[:-1], f32) for _ in range(4)]
    full = n - n % 4
    for i in range(0, full, 4):
        for l in range(4):
            lanes[l] = (lanes[l] + sq[..., i + l]).astype(f32)
    for i in range(full, n):
        lanes[0] = (lanes[0] + sq[..., i]).astype(f32)
    return (((lanes[0] + lanes[1]).astype(f32) + lanes[2]).astype(f32) + lanes[3]).astype(f32)


def kmeans_py(data, k, init_rows, max_iter):
    n, dim = data.shape
    cent = data[init_rows].copy()
    assign = np.full(n, -1, np.int64)
    iters = 0
    for it in range(max_iter):
        iters = it + 1
        dist = l2sq_lanes(data[:, None, :], cent[None, :, :])          # [n, k]
        best = np.argmin(dist, axis=1)                                   # first minimum == strict '<' scan
        changed = int((assign != best).sum())
        assign = best
        sums = np.zeros((k, dim), f32)
        counts = np.zeros(k, np.int64)
        for i in range(n):                                               # data order: float32 sums are order-sensitive
            sums[best[i]] = (sums[best[i]] + data[i]).astype(f32)
            counts[best[i]] += 1
        assert (counts > 0).all(), "empty cluster: the re-seed rule is not part of this restatement"
        cent = (sums / counts.astype(f32)[:, None]).astype(f32)
        if it > 0 and changed < n // 1000 + 1:
            break
    return cent, iters


def encode_py(vec, codebooks):
    M, K, sub = codebooks.shape
    codes = np.zeros(M, np.uint8)
    for m in range(M):
        d = l2sq_lanes(vec[m * sub:(m + 1) * sub][None, :], codebooks[m])
        if K > 16:                                                       # simd.go:300-325: sqrt'd batch distances
            d = np.array([f32(math.sqrt(float(x))) for x in d], f32)
        codes[m] = int(np.argmin(d))
    return codes


def _clustered(rng, n, dim, k):
    """n rows around k well separated centres, every centre populated; returns (rows, one row index per centre)."""
    centres = rng.standard_normal((k, dim)).astype(f32) * f32(6)
    label = rng.permutation(np.arange(n) % k)
    rows = (centres[label] + rng.standard_normal((n, dim)).astype(f32) * f32(0.2)).astype(f32)
    first = np.array([int(np.flatnonzero(label == c)[0]) for c in range(k)], np.int32)
    return rows, first


@pytest.mark.parametrize("n,M,sub,K", [(600, 2, 4, 16), (900, 3, 5, 8), (1500, 1, 8, 32)])
def test_oracle_kmeans_matches_python_restatement(oracle, n, M, sub, K):
    rng = np.random.default_rng(n + K)
    parts = [_clustered(rng, n, sub, K) for _ in range(M)]
    data = np.concatenate([p[0] for p in parts], axis=1)
    # one initial row per true cluster keeps every cluster populated (the re-seed rule never fires)
    init = np.stack([p[1] for p in parts]).astype(np.int32)
    want = [kmeans_py(np.ascontiguousarray(data[:, m * sub:(m + 1) * sub]), K, init[m], 20) for m in range(M)]
    got_cb, got_it = oracle.pq_train(data, M, K, init, 20)
    for m in range(M):
        assert got_it[m] == want[m][1], (m, got_it[m], want[m][1])
        assert np.array_equal(got_cb[m], want[m][0]), m


@pytest.mark.parametrize("M,sub,K", [(4, 4, 16), (3, 6, 8), (2, 8, 256), (5, 3, 17)])
def test_oracle_encode_and_adc_match_python_restatement(oracle, M, sub, K):
    rng = np.random.default_rng(M * 100 + K)
    cb = rng.standard_normal((M, K, sub)).astype(f32)
    vecs = rng.standard_normal((40, M * sub)).astype(f32)
    vecs[3] = cb[:, K // 2, :].reshape(-1)            # exactly a centroid in every subspace
    cb[:, K - 1, :] = cb[:, K // 2, :]                 # and a duplicate of it later: the first one must win
    got = oracle.pq_encode(vecs, cb)
    for i, v in enumerate(vecs):
        assert np.array_equal(got[i], encode_py(v, cb)), i
    assert (got[3] == K // 2).all()
    # ADC table and the un-sqrt'd single distance (adc_table.go:15-51,77-92)
    q = rng.standard_normal(M * sub).astype(f32)
    table = np.stack([l2sq_lanes(q[m * sub:(m + 1) * sub][None, :], cb[m]) for m in range(M)]).reshape(-1)
    if K == 256:
        assert np.array_equal(oracle.adc_table(q, cb), table)
        flat = got.astype(np.uint8)
        want = np.empty(len(flat), f32)
        for i, code in enumerate(flat):                # simd.go:345-355: sequential fp32 sum, sqrt through float64
            s = f32(0)
            for j in range(M):
                s = f32(s + table[j * 256 + int(code[j])])
            want[i] = f32(math.sqrt(float(s)))
        assert np.array_equal(oracle.adc_batch(oracle.adc_table(q, cb), flat), want)
