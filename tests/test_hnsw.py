"""HNSW layer search: the C restatement of ArrowHNSW.searchLayer (oracle, checked here against an independent
pure-Python restatement and against brute force), and the GPU walk (csrc/hnsw.cu) against the oracle -- identical
frontier ids, distances and visited counts -- plus config 5 end to end (walk -> in-kernel-bitmap re-rank)."""
import heapq

import numpy as np
import pytest

from tests.util import assert_topk_equal, make_db, random_bitmap

L2, COS, DOT = 0, 1, 2


def build_graph(rng, db, degree, n_random, counts_ragged=True):
    """kNN edges (numpy, exact) + a few random long-range edges; ragged counts, 0xffffffff padding, one duplicate."""
    n = db.shape[0]
    x = db.astype(np.float32)
    sq = (x * x).sum(1)
    nbrs = np.full((n, degree), 0xFFFFFFFF, np.uint32)
    blk = 2000
    for lo in range(0, n, blk):
        d = sq[lo:lo + blk, None] + sq[None, :] - 2.0 * (x[lo:lo + blk] @ x.T)
        d[np.arange(d.shape[0]), np.arange(lo, lo + d.shape[0])] = np.inf
        nbrs[lo:lo + blk, :degree - n_random] = np.argsort(d, axis=1)[:, :degree - n_random]
    nbrs[:, degree - n_random:] = rng.integers(0, n, (n, n_random))
    counts = np.full(n, degree, np.int32)
    if counts_ragged:
        short = rng.random(n) < 0.2
        counts[short] = rng.integers(1, degree, short.sum())
        for i in np.nonzero(short)[0]:
            nbrs[i, counts[i]:] = 0xFFFFFFFF
        nbrs[5, 1] = nbrs[5, 0]  # a duplicate neighbour
    return nbrs, counts


def py_search_layer(dist, neighbors, counts, entry, ef):
    """arrow_hnsw.go:1108-1385 with heapq and (distance, id) tuples -- independent of the C restatement."""
    visited = {entry}
    ed = dist(entry)
    cand = [(ed, entry)]
    res = [(-ed, -entry)]  # max-heap on (d, id)
    nv = 1
    while cand:
        d, c = heapq.heappop(cand)
        if len(res) >= ef and d > -res[0][0]:
            break
        for i in range(counts[c]):
            nb = int(neighbors[c, i])
            if nb >= len(counts) or nb in visited:
                continue
            visited.add(nb)
            dn = dist(nb)
            nv += 1
            if len(res) < ef or dn < -res[0][0]:
                heapq.heappush(cand, (dn, nb))
                heapq.heappush(res, (-dn, -nb))
                if len(res) > ef:
                    heapq.heappop(res)
    out = sorted((-a, -b) for a, b in res)
    return [b for _, b in out], [a for a, _ in out], nv


def test_oracle_walk_matches_python_restatement(oracle):
    rng = np.random.default_rng(3)
    n, dim, deg = 1500, 24, 12
    db = make_db(rng, n, dim, np.float32)
    nbrs, counts = build_graph(rng, db, deg, 3)
    qs = make_db(rng, 12, dim, np.float32)
    entries = rng.integers(0, n, 12).astype(np.uint32)
    for ef in (1, 8, 40):
        ids, d, nv = oracle.hnsw_search_layer(L2, db, nbrs, counts, qs, entries, ef)
        for qi in range(12):
            dist = lambda i: np.float32(oracle.distance(L2, qs[qi], db[i]))
            pi, pd, pnv = py_search_layer(dist, nbrs, counts, int(entries[qi]), ef)
            assert ids[qi, :len(pi)].tolist() == pi and (ids[qi, len(pi):] == 0xFFFFFFFF).all()
            assert np.array_equal(d[qi, :len(pd)], np.array(pd, np.float32))
            assert nv[qi] == pnv


def test_oracle_walk_recall_and_order(oracle):
    rng = np.random.default_rng(4)
    n, dim, deg = 4000, 32, 16
    db = make_db(rng, n, dim, np.float32)
    nbrs, counts = build_graph(rng, db, deg, 4, counts_ragged=False)
    qs = make_db(rng, 30, dim, np.float32)
    ids, d, nv = oracle.hnsw_search_layer(L2, db, nbrs, None, qs, np.zeros(30, np.uint32), 64)
    assert (np.diff(d, axis=1) >= 0).all()
    wd, wl = oracle.search(L2, db, qs, 10)
    recall = np.mean([len(set(ids[i, :10].tolist()) & set(wl[i].tolist())) / 10 for i in range(30)])
    assert recall > 0.7, recall
    assert (nv < n).all() and (nv >= 64).all()


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,metric", [(np.float32, L2), (np.float32, COS), (np.float16, L2), (np.float16, DOT),
                                          (np.int8, L2), (np.int8, DOT)])
@pytest.mark.parametrize("ef", [1, 16, 128])
def test_gpu_walk_equals_oracle(oracle, dtype, metric, ef):
    from longbow_b200 import gpu, store
    rng = np.random.default_rng(100 + ef)
    n, dim, deg = 12000, 64, 32
    db = make_db(rng, n, dim, dtype)
    nbrs, counts = build_graph(rng, db, deg, 6)
    qs = make_db(rng, 200, dim, dtype)
    entries = rng.integers(0, n, 200).astype(np.uint32)
    idx = gpu.DenseIndex(dim, dtype, metric)
    idx.add(db)
    g = store.HNSWGraph(idx, deg)
    g.SetNeighbors(nbrs, counts)
    gi, gd, gv = g.SearchLayer(qs, entries, ef)
    wi, wd, wv = oracle.hnsw_search_layer(metric, db, nbrs, counts, qs, entries, ef)
    assert np.array_equal(gi, wi), f"frontier ids differ ({(gi != wi).sum()} slots)"
    assert np.array_equal(gd, wd)
    assert np.array_equal(gv.astype(np.int64), wv)
    g.Close(); idx.close()


@pytest.mark.gpu
def test_gpu_walk_then_rerank_with_bitmaps(oracle):
    """Config 5 in small: ef=128 walk, tombstones 5 % + allow 30 % applied by the re-rank kernel, k=10."""
    from longbow_b200 import gpu, store
    rng = np.random.default_rng(55)
    n, dim, deg, nq, ef, k = 30000, 96, 32, 500, 128, 10
    db = make_db(rng, n, dim, np.float32)
    nbrs, counts = build_graph(rng, db, deg, 6)
    qs = make_db(rng, nq, dim, np.float32)
    entries = np.zeros(nq, np.uint32)
    tomb, allow = random_bitmap(rng, n, 0.05), random_bitmap(rng, n, 0.3)
    idx = gpu.DenseIndex(dim, np.float32, L2)
    idx.add(db)
    idx.set_tombstones(tomb)
    g = store.HNSWGraph(idx, deg)
    g.SetNeighbors(nbrs, counts)
    gd, gl = g.Search(qs, entries, ef, k, allow=allow)
    wi, _, _ = oracle.hnsw_search_layer(L2, db, nbrs, counts, qs, entries, ef)
    cand = wi.astype(np.int64)
    cand[wi == 0xFFFFFFFF] = -1
    wd, wl = oracle.rerank(L2, db, qs, cand, k, tomb=gpu.pack_bitmap(tomb), allow=gpu.pack_bitmap(allow))
    assert_topk_equal(gd, gl, wd, wl, 0.0, "walk + rerank")
    g.Close(); idx.close()


@pytest.mark.gpu
def test_gpu_walk_small_table_retries(oracle):
    """ef = 1 sizes the visited table at its minimum; a dense graph walk that outgrows it must be retried with a
    larger table by the host call, not return garbage."""
    from longbow_b200 import gpu, store
    rng = np.random.default_rng(77)
    n, dim, deg = 20000, 16, 64
    db = make_db(rng, n, dim, np.float32)
    nbrs = rng.integers(0, n, (n, deg)).astype(np.uint32)   # random graph: walks wander far
    qs = make_db(rng, 20, dim, np.float32)
    entries = np.zeros(20, np.uint32)
    idx = gpu.DenseIndex(dim, np.float32, L2)
    idx.add(db)
    g = store.HNSWGraph(idx, deg)
    g.SetNeighbors(nbrs, None)
    gi, gd, gv = g.SearchLayer(qs, entries, 40)
    wi, wd, wv = oracle.hnsw_search_layer(L2, db, nbrs, None, qs, entries, 40)
    assert np.array_equal(gi, wi) and np.array_equal(gd, wd)
    g.Close(); idx.close()


@pytest.mark.gpu
def test_arrow_function_registry(oracle):
    """internal/store/arrow_kernels_test.go:56-69: l2_distance {0, sqrt(30)}; select_k_neighbors returns indices."""
    from longbow_b200 import store
    store.RegisterHNSWKernels(); store.RegisterHNSWKernels()  # idempotent
    rows = np.array([[1, 2, 3, 4], [2, 4, 6, 8]], np.float32)
    out = store.CallFunction("l2_distance", np.array([1, 2, 3, 4], np.float32), rows)
    assert out[0] == 0.0 and abs(out[1] - np.sqrt(30.0)) <= 1e-6
    idx = store.CallFunction("select_k_neighbors", np.array([0.5, 0.1, 0.9, 0.3], np.float32),
                             np.array([10, 11, 12, 13], np.uint32), 2)
    assert idx.tolist() == [1, 3]
    with pytest.raises(KeyError):
        store.CallFunction("no_such_function")
