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
