"""Parity at BASELINE.json's FULL sizes (one GPU), through properties that do not need the whole answer
from the CPU: planted neighbours, (distance, id) sortedness, exact re-computation of every returned distance,
agreement of two independent coarse paths (tensor-core vs SIMT scan), shard-and-merge == whole, idempotence,
plus the O-exact oracle on a handful of queries against the full database."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def _sorted_by_dist_then_id(d, l):
    for q in range(d.shape[0]):
        valid = l[q] >= 0
        dd, ll = d[q][valid], l[q][valid]
        order = np.lexsort((ll, dd))
        if not np.array_equal(order, np.arange(len(dd))):
            return False
    return True


def test_c2_full_size_properties(torch_cuda, oracle):
    """C2: 1M x 768 fp16 cosine, 1024 queries, k=100."""
    torch = torch_cuda
    from longbow_b200 import _lib, gpu
    dev = torch.device("cuda", 0)
    N, D, Q, K = 1_000_000, 768, 1024, 100
    g = torch.Generator(device=dev).manual_seed(2001)
    db = torch.randn((N, D), generator=g, device=dev)
    db = (db / db.norm(dim=1, keepdim=True)).half()
    db[3] = 0
    db[N - 1] = 0  # zero rows: cosine distance exactly 1.0
    qs = torch.randn((Q, D), generator=g, device=dev)
    qs = (qs / qs.norm(dim=1, keepdim=True)).half()
    planted = torch.randint(0, N, (Q,), generator=g, device=dev)
    qs[:64] = db[planted[:64]]          # query == a database row: that row (or an identical one) must rank first
    idx = gpu.DenseIndex(D, np.float16, _lib.METRIC_COSINE)
    idx.reserve(N)
    idx.add_device(db)
    od = torch.empty((Q, K), dtype=torch.float32, device=dev)
    ol = torch.empty((Q, K), dtype=torch.int64, device=dev)
    idx.search_device(qs, K, od, ol)
    torch.cuda.synchronize()
    d, l = od.cpu().numpy(), ol.cpu().numpy()
    assert (l >= 0).all() and _sorted_by_dist_then_id(d, l)
    assert len(set(map(tuple, np.sort(l, axis=1)))) > 1 and all(len(set(r)) == K for r in l)  # no duplicate ids
    p = planted[:64].cpu().numpy()
    assert np.array_equal(l[:64, 0], p) and (np.abs(d[:64, 0]) < 1e-3).all()
    # idempotence
    od2, ol2 = torch.empty_like(od), torch.empty_like(ol)
    idx.search_device(qs, K, od2, ol2)
    torch.cuda.synchronize()
    assert torch.equal(ol, ol2) and torch.equal(od, od2)
    # two independent coarse paths agree (tensor-core scan vs SIMT scan), 32 queries
    _lib.set_option("dense_scan", 1)
    try:
        od3 = torch.empty((32, K), dtype=torch.float32, device=dev)
        ol3 = torch.empty((32, K), dtype=torch.int64, device=dev)
        idx.search_device(qs[100:132].contiguous(), K, od3, ol3)
        torch.cuda.synchronize()
    finally:
        _lib.set_option("dense_scan", 0)
    assert torch.equal(ol3, ol[100:132]) and torch.equal(od3, od[100:132])
    # every returned distance equals the reference arithmetic on that (query, row) pair; and the O-exact
    # oracle's full answer for a few queries
    h_db = db.cpu().numpy()
    h_q = qs.cpu().numpy()
    for qi in (0, 500, 1023):
        rows = h_db[l[qi]]
        want = np.array([oracle.distance(oracle.COSINE, h_q[qi], r) for r in rows], np.float32)
        assert np.array_equal(want, d[qi])
    sel = [5, 700]
    wd, wl = oracle.search(oracle.COSINE, h_db, h_q[sel], K)
    assert np.array_equal(wl, l[sel]) and np.array_equal(wd, d[sel])
    # shard-and-merge == whole (exactness of the multi-GPU merge), two shards on one device
    half = N // 2
    parts_d, parts_l = [], []
    for lo, hi in ((0, half), (half, N)):
        sh = gpu.DenseIndex(D, np.float16, _lib.METRIC_COSINE)
        sh.reserve(hi - lo)
        sh.add_device(db[lo:hi].contiguous())
        sh.set_id_base(lo)
        sd = torch.empty((Q, K), dtype=torch.float32, device=dev)
        sl = torch.empty((Q, K), dtype=torch.int64, device=dev)
        sh.search_device(qs, K, sd, sl)
        torch.cuda.synchronize()
        parts_d.append(sd)
        parts_l.append(sl)
        sh.close()
    gd, gl = torch.stack(parts_d), torch.stack(parts_l)
    md, ml = torch.empty_like(od), torch.empty_like(ol)
    _lib.check(_lib.load().lb_merge_topk_device(0, gd.data_ptr(), gl.data_ptr(), 2, Q, K, K, md.data_ptr(), ml.data_ptr(),
                                                torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert torch.equal(ml, ol) and torch.equal(md, od)
    idx.close()


def test_c4_shard_full_size_properties(torch_cuda, oracle):
    """C4: one GPU's shard of the 100M x 128 int8 database (12.5M rows), dot metric, k=10: bit-exact."""
    torch = torch_cuda
    from longbow_b200 import _lib, gpu
    dev = torch.device("cuda", 0)
    N, D, Q, K = 12_500_000, 128, 1024, 10
    g = torch.Generator(device=dev).manual_seed(4001)
    db = torch.randint(-128, 128, (N, D), generator=g, device=dev, dtype=torch.int8)
    qs = torch.randint(-128, 128, (Q, D), generator=g, device=dev, dtype=torch.int8)
    idx = gpu.DenseIndex(D, np.int8, _lib.METRIC_DOT)
    idx.reserve(N)
    idx.add_device(db)
    idx.set_id_base(25_000_000)  # as rank 2 of 8 would
    od = torch.empty((Q, K), dtype=torch.float32, device=dev)
    ol = torch.empty((Q, K), dtype=torch.int64, device=dev)
    idx.search_device(qs, K, od, ol)
    torch.cuda.synchronize()
    d, l = od.cpu().numpy(), ol.cpu().numpy()
    assert _sorted_by_dist_then_id(d, l) and (l >= 25_000_000).all() and (l < 25_000_000 + N).all()
    # exact integer re-computation of every returned distance (torch int64 on the device)
    rows = db[(ol - 25_000_000).reshape(-1)].to(torch.int64).reshape(Q, K, D)
    want = -(rows * qs.to(torch.int64)[:, None, :]).sum(dim=2)
    assert torch.equal(want.to(torch.float32), od)
    # global optimality on the device: no row beats the k-th result (int64 matmul in row blocks, 8 queries)
    qsel = qs[:8].to(torch.float32)
    best = torch.full((8,), float("inf"), device=dev)
    cnt_better = torch.zeros(8, dtype=torch.int64, device=dev)
    kth = od[:8, K - 1]
    for lo in range(0, N, 2_500_000):
        blk = db[lo:lo + 2_500_000].to(torch.float32)
        dist = -(qsel @ blk.T)  # exact: |dot| < 2^24
        cnt_better += (dist < kth[:, None]).sum(dim=1)
    assert (cnt_better <= K - 1).all()
    # O-exact oracle on two queries over the full shard
    h_db, h_q = db.cpu().numpy(), qs[:2].cpu().numpy()
    wd, wl = oracle.search(oracle.DOT, h_db, h_q, K, id_base=25_000_000)
    assert np.array_equal(wl, l[:2]) and np.array_equal(wd, d[:2])
    idx.close()


def test_c3_full_size_pq_bit_exact(torch_cuda, oracle):
    """C3: 10M x 96 PQ codes, ADC scan, k=10: bit-exact ids and distances vs the oracle for 3 queries, and
    scan == stateless simd.ADCDistanceBatch on the returned rows."""
    torch = torch_cuda
    from longbow_b200 import pq
    dev = torch.device("cuda", 0)
    N, D, M, K = 10_000_000, 768, 96, 10
    g = torch.Generator(device=dev).manual_seed(3001)
    cb = torch.randn((M, 256, D // M), generator=g, device=dev).cpu().numpy()
    codes = torch.randint(0, 256, (N, M), generator=g, device=dev, dtype=torch.uint8)
    qs = torch.randn((3, D), generator=g, device=dev).cpu().numpy()
    enc = pq.PQEncoder(D, M, 256, cb)
    enc.add_codes_device(codes)
    d, l = enc.search(qs, K)
    h_codes = codes.cpu().numpy()
    wd, wl = oracle.pq_search(cb, h_codes, None, qs, K, 0)
    assert np.array_equal(l, wl) and np.array_equal(d, wd)
    table = enc.BuildADCTable(qs[0])
    out = np.empty(K, np.float32)
    enc.ADCDistanceBatch(table, h_codes[l[0]], out)
    assert np.array_equal(out, d[0])
    enc.close()


def test_c5_full_size_rerank(torch_cuda, oracle):
    """C5: 4096 queries x 128 candidate ids over 10M x 384 fp32 with tombstones (5%) and an allow bitmap (30%).
    Checked against the oracle on a compacted copy of the touched rows for 64 queries, and by properties."""
    torch = torch_cuda
    from longbow_b200 import _lib, gpu
    dev = torch.device("cuda", 0)
    N, D, Q, C, K = 10_000_000, 384, 4096, 128, 10
    g = torch.Generator(device=dev).manual_seed(5001)
    idx = gpu.DenseIndex(D, np.float32, _lib.METRIC_L2)
    idx.reserve(N)
    chunks = []
    for lo in range(0, N, 2_000_000):
        c = torch.randn((2_000_000, D), generator=g, device=dev)
        idx.add_device(c)
        chunks.append(c)
    qs = torch.randn((Q, D), generator=g, device=dev)
    cand = torch.randint(0, N, (Q, C), generator=g, device=dev, dtype=torch.int64)
    tomb = torch.rand(N, generator=g, device=dev) < 0.05
    allow = torch.rand(N, generator=g, device=dev) < 0.30
    h_tomb, h_allow = tomb.cpu().numpy(), allow.cpu().numpy()
    idx.set_tombstones(h_tomb)
    allow_d = torch.from_numpy(gpu.pack_bitmap(h_allow).view(np.int64)).to(dev)
    od = torch.empty((Q, K), dtype=torch.float32, device=dev)
    ol = torch.empty((Q, K), dtype=torch.int64, device=dev)
    idx.rerank_device(qs, cand.to(torch.uint32), K, od, ol, allow=allow_d)
    torch.cuda.synchronize()
    d, l = od.cpu().numpy(), ol.cpu().numpy()
    assert _sorted_by_dist_then_id(d, l)
    valid = l >= 0
    assert (~h_tomb[l[valid]]).all() and h_allow[l[valid]].all()          # bitmaps honoured
    h_cand = cand.cpu().numpy()
    for qi in range(0, Q, 97):                                            # returned ids come from the candidates
        assert set(l[qi][l[qi] >= 0]) <= set(h_cand[qi])
    # oracle on a compacted database: the rows the first 64 queries touch
    nsel = 64
    ids = np.unique(h_cand[:nsel])
    remap = {int(v): i for i, v in enumerate(ids)}
    full = torch.cat(chunks)
    small = full[torch.from_numpy(ids).to(dev)].cpu().numpy()
    del full
    c_small = np.vectorize(remap.get)(h_cand[:nsel]).astype(np.int64)
    wd, wl = oracle.rerank(oracle.L2, small, qs[:nsel].cpu().numpy(), c_small, K,
                           tomb=gpu.pack_bitmap(h_tomb[ids]), allow=gpu.pack_bitmap(h_allow[ids]))
    wl_global = np.where(wl >= 0, ids[np.clip(wl, 0, len(ids) - 1)], -1)
    assert np.array_equal(wl_global, l[:nsel]) and np.array_equal(wd, d[:nsel])
    idx.close()
