"""An independent numpy restatement of the reference's dense distance kernels and brute-force top-k, written from the
Go sources (not from oracle/lb_oracle.c), that the C oracle must agree with bit for bit.  CPU only.

* euclideanUnrolled4x / cosineUnrolled4x / dotUnrolled4x   internal/simd/simd.go:365-479
* the F16 and int8 forms (widen every element to float32, same four lanes)   simd.go:767-848, simd_baseline.go:13-54
* brute-force top-k: per query a scan with a size-k heap under a strict '<' (internal/store/adaptive_index.go:161-225),
  i.e. the k smallest by (distance, row); tombstoned / not-allowed rows skipped
"""
import numpy as np
import pytest

f32 = np.float32
L2, COS, DOT = 0, 1, 2


def lane_sum(terms):
    """terms [rows, d] float32 -> [rows]: lane l accumulates elements i == l (mod 4) of the full groups of four in
    increasing i, the 0-3 remainder elements go to lane 0, then ((s0 + s1) + s2) + s3 -- all in float32."""
    rows, d = terms.shape
    s = [np.zeros(rows, f32) for _ in range(4)]
    full = d - d % 4
    for i in range(0, full, 4):
        for l in range(4):
            s[l] = (s[l] + terms[:, i + l]).astype(f32)
    for i in range(full, d):
        s[0] = (s[0] + terms[:, i]).astype(f32)
    return (((s[0] + s[1]).astype(f32) + s[2]).astype(f32) + s[3]).astype(f32)


def distances_py(metric, q, db):
    """One query against every row; inputs of any dtype are widened to float32 first (exact for fp16 / int8)."""
    a = np.broadcast_to(q.astype(f32), db.shape)
    b = db.astype(f32)
    if b.shape[1] == 0:                                   # distance_functions.go:21-23,51-53,63-65
        return np.full(b.shape[0], 1.0 if metric == COS else 0.0, f32)
    if metric == L2:
        d = (a - b).astype(f32)
        return np.sqrt(lane_sum((d * d).astype(f32)).astype(np.float64)).astype(f32)
    dot = lane_sum((a * b).astype(f32))
    if metric == DOT:
        return (-dot).astype(f32)                         # as a distance (the index negates the raw similarity)
    na, nb = lane_sum((a * a).astype(f32)), lane_sum((b * b).astype(f32))
    with np.errstate(divide="ignore", invalid="ignore"):
        den = np.sqrt(na.astype(np.float64) * nb.astype(np.float64)).astype(f32)
        out = (f32(1.0) - (dot / den).astype(f32)).astype(f32)
    out[(na == 0) | (nb == 0)] = f32(1.0)                 # exactly 1.0 (simd.go:446-448)
    return out


def search_py(metric, db, queries, k, dead=None, allow=None):
    n = db.shape[0]
    od = np.full((len(queries), k), np.finfo(f32).max, f32)
    ol = np.full((len(queries), k), -1, np.int64)
    live = np.ones(n, bool)
    if dead is not None:
        live &= ~dead
    if allow is not None:
        live &= allow
    ids = np.flatnonzero(live)
    for qi, q in enumerate(queries):
        d = distances_py(metric, q, db)[ids]
        order = np.lexsort((ids, d))[:k]                  # (distance, row) ascending
        od[qi, :len(order)] = d[order]
        ol[qi, :len(order)] = ids[order]
    return od, ol


def _data(rng, n, dim, dtype):
    if dtype == np.int8:
        return rng.integers(-128, 128, (n, dim), dtype=np.int8)
    x = rng.standard_normal((n, dim)).astype(f32)
    return x.astype(dtype)


@pytest.mark.parametrize("dtype", [np.float32, np.float16, np.int8], ids=["f32", "f16", "i8"])
@pytest.mark.parametrize("dim", [1, 2, 3, 4, 5, 7, 8, 9, 33, 128, 384, 769])
def test_oracle_search_matches_python_restatement(oracle, dtype, dim):
    rng = np.random.default_rng(dim * 7 + np.dtype(dtype).itemsize)
    n, nq, k = 257, 4, 10
    db, q = _data(rng, n, dim, dtype), _data(rng, nq, dim, dtype)
    db[17] = db[200]                                      # an exact tie: the lower row first
    if dtype != np.int8:
        db[5] = 0                                         # cosine's zero-row rule
    q[1] = db[200]
    dead, allow = rng.random(n) < 0.1, rng.random(n) < 0.7
    from longbow_b200.gpu import pack_bitmap
    for metric in ((L2, DOT) if dtype == np.int8 else (L2, COS, DOT)):
        wd, wl = search_py(metric, db, q, k)
        gd, gl = oracle.search(metric, db, q, k)
        assert np.array_equal(gl, wl) and np.array_equal(gd, wd), (metric, dim)
        wd, wl = search_py(metric, db, q, k, dead, allow)
        gd, gl = oracle.search(metric, db, q, k, tomb=pack_bitmap(dead), allow=pack_bitmap(allow))
        assert np.array_equal(gl, wl) and np.array_equal(gd, wd), (metric, dim, "bitmaps")


def test_oracle_search_k_larger_than_live_rows(oracle):
    rng = np.random.default_rng(3)
    db, q = _data(rng, 6, 16, np.float32), _data(rng, 2, 16, np.float32)
    wd, wl = search_py(L2, db, q, 10)
    gd, gl = oracle.search(L2, db, q, 10)
    assert np.array_equal(gl, wl) and np.array_equal(gd[gl >= 0], wd[wl >= 0])
    assert (gl[:, 6:] == -1).all()


# ------------------------------------------------------------------ re-rank, shard merge, select_k, SQ8
def rerank_py(metric, db, q, cand, k, dead=None, allow=None):
    """internal/store/hnsw_batch.go:206-245: candidates that resolve to a stored vector (here: in range, not
    tombstoned, allowed) keep their exact distance; ascending by (distance, id); at most k."""
    d_all = distances_py(metric, q, db)
    keep = [int(c) for c in cand if 0 <= c < db.shape[0] and (dead is None or not dead[c]) and (allow is None or allow[c])]
    keep.sort(key=lambda c: (d_all[c], c))
    return [(c, d_all[c]) for c in keep[:k]]


def test_oracle_rerank_merge_select_match_python_restatement(oracle):
    from longbow_b200.gpu import pack_bitmap
    rng = np.random.default_rng(9)
    n, dim, nq, c, k = 500, 24, 6, 40, 10
    db, q = _data(rng, n, dim, np.float32), _data(rng, nq, dim, np.float32)
    db[3] = db[400]
    cand = np.stack([rng.permutation(n + 30)[:c] for _ in range(nq)]).astype(np.int64)   # some ids past the index
    cand[0, :2] = (3, 400)
    dead, allow = rng.random(n) < 0.1, rng.random(n) < 0.8
    for metric in (L2, COS, DOT):
        gd, gl = oracle.rerank(metric, db, q, cand, k, tomb=pack_bitmap(dead), allow=pack_bitmap(allow))
        for qi in range(nq):
            want = rerank_py(metric, db, q[qi], cand[qi], k, dead, allow)
            assert [int(x) for x in gl[qi][:len(want)]] == [w[0] for w in want], (metric, qi)
            assert [f32(x) for x in gd[qi][:len(want)]] == [w[1] for w in want]
            assert (gl[qi][len(want):] == -1).all()
    # shard merge (internal/store/sharded_hnsw.go:494-503): the k smallest of the concatenation by (distance, label)
    parts, k_in = 5, 8
    d = np.sort(rng.random((parts, nq, k_in)).astype(f32), axis=2)
    d[rng.random(d.shape) < 0.3] = f32(0.25)
    d = np.sort(d, axis=2)
    l = rng.permutation(parts * nq * k_in).reshape(parts, nq, k_in).astype(np.int64)
    l[2, :, 5:] = -1
    md, ml = oracle.merge(d, l, k)
    for qi in range(nq):
        pairs = sorted((float(d[p, qi, j]), int(l[p, qi, j])) for p in range(parts) for j in range(k_in) if l[p, qi, j] >= 0)
        assert [(float(a), int(b)) for a, b in zip(md[qi], ml[qi])] == pairs[:k]
    # select_k (internal/store/arrow_kernels.go:230-345): indices of the k smallest values, (value, index) order
    x = rng.random(5000).astype(f32)
    x[77] = x[4000]
    sd, si = oracle.select_k(x, 25)
    order = np.lexsort((np.arange(x.size), x))[:25]
    assert np.array_equal(si, order) and np.array_equal(sd, x[order])


def test_oracle_sq8_distance_matches_python_restatement(oracle):
    """EuclideanSQ8Generic (internal/simd/sq8.go:45-66): exact int32 sum of squared byte differences; the batch form
    reports float32(d) (simd.go:170-182)."""
    rng = np.random.default_rng(10)
    for dim in (0, 1, 7, 8, 9, 33, 127, 1024):
        a = rng.integers(0, 256, dim, dtype=np.uint8)
        rows = rng.integers(0, 256, (5, dim), dtype=np.uint8)
        want = ((rows.astype(np.int64) - a.astype(np.int64)) ** 2).sum(axis=1)
        assert want.max(initial=0) < 2 ** 31
        if dim == 0:
            continue
        got = oracle.batch_flat(L2, a, rows)
        assert np.array_equal(got, want.astype(np.int32).astype(f32)), dim
