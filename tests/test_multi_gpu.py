"""Multi-GPU exchange tests: the real chain -- per-GPU search (lb_index_search_device_cert) ->
exchange of the top-k records over peer memory (csrc/exchange.cu) or one packed NCCL all-gather -> merge kernel
-- against the CPU oracle's single-index answer.

* single-GPU cases run anywhere with one B200: packed-record merge, world-1 exchange, device-path certification;
* two-device cases (one process, peer access; and two processes, CUDA IPC + NCCL) are skipped when the box has
  fewer than two GPUs.
"""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest

from tests.util import assert_topk_equal, make_db

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
L2, COS, DOT = 0, 1, 2


def _ngpu():
    import torch
    return torch.cuda.device_count()


def test_packed_merge_matches_oracle(oracle):
    """lb_merge_topk_packed_device over [distances | labels] records == oracle merge (ties across parts, padding)."""
    import torch
    from longbow_b200 import _lib
    from longbow_b200.shard import record_layout
    rng = np.random.default_rng(5)
    parts, nq, k = 4, 37, 10
    d = np.sort(rng.random((parts, nq, k), dtype=np.float32), axis=2)
    d[rng.random(d.shape) < 0.3] = 0.25  # long runs of equal distances across parts
    d = np.sort(d, axis=2)
    l = rng.permutation(parts * nq * k).reshape(parts, nq, k).astype(np.int64) + (1 << 34)
    l[2, :, 6:] = -1
    loff, rec = record_layout(nq, k)
    buf = np.zeros((parts, rec), np.uint8)
    for p in range(parts):
        buf[p, :nq * k * 4] = d[p].view(np.uint8).reshape(-1)
        buf[p, loff:loff + nq * k * 8] = l[p].view(np.uint8).reshape(-1)
    t = torch.from_numpy(buf).cuda()
    od = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    ol = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    _lib.check(_lib.load().lb_merge_topk_packed_device(0, t.data_ptr(), rec, loff, parts, nq, k, k, od.data_ptr(),
                                                      ol.data_ptr(), torch.cuda.current_stream().cuda_stream))
    wd, wl = oracle.merge(d, l, k)
    assert_topk_equal(od.cpu().numpy(), ol.cpu().numpy(), wd, wl, 0.0, "packed merge")


def test_merge_all_equal_distances(oracle):
    """Pathological tie run: every entry has the same distance; the merge must order purely by label."""
    from longbow_b200 import store
    rng = np.random.default_rng(6)
    parts, nq, k_in, k = 8, 3, 64, 50
    d = np.full((parts, nq, k_in), 0.5, np.float32)
    l = np.sort(rng.permutation(parts * nq * k_in).reshape(parts, nq, k_in).astype(np.int64), axis=2)
    gd, gl = store.MergeShardResults(d, l, k)
    wd, wl = oracle.merge(d, l, k)
    assert_topk_equal(gd, gl, wd, wl, 0.0, "all-equal merge")


def test_device_path_certification(oracle):
    """lb_index_search_device_cert flags the near-tie query on the device (no host sync), the accumulating
    counter counts it, and lb_index_search_exact_device repairs exactly the flagged rows."""
    import torch
    from longbow_b200 import gpu
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
    q = np.stack([base, db[5], db[12345]])
    idx = gpu.DenseIndex(dim, np.float16, COS)
    idx.add(db)
    dq = torch.from_numpy(q).cuda()
    od = torch.empty((3, k), dtype=torch.float32, device="cuda")
    ol = torch.empty((3, k), dtype=torch.int64, device="cuda")
    flags = torch.full((3,), 7, dtype=torch.int32, device="cuda")
    count = torch.zeros(1, dtype=torch.int32, device="cuda")
    idx.search_device(dq, k, od, ol, uncert_flags=flags, uncert_count=count)
    idx.search_device(dq, k, od, ol, uncert_flags=flags, uncert_count=count)  # the counter accumulates
    f = flags.cpu().numpy()
    # query 0 sits inside the 300-row near-tie cluster; row 17 (the cluster's centre) is the 26th nearest row of
    # query 1, so the cluster straddles its rank-100 boundary as well; query 2 is far from it
    assert f[0] == 1 and f[1] == 1 and f[2] == 0, f
    assert int(count.item()) == 4
    idx.search_exact_device(dq, k, od, ol, flags_host=f)
    wd, wl = oracle.search(COS, db, q, k)
    assert_topk_equal(od.cpu().numpy(), ol.cpu().numpy(), wd, wl, 0.0, "device path after repair")
    idx.close()


def test_simt_l2_certification_key_space(oracle):
    """the SIMT scan ranks L2 by |q - x|^2, not |x|^2 - 2 q.x; its certification must test in that key
    space.  Odd dim forces the SIMT scan; a cluster of near-identical rows larger than the margin must be flagged
    and repaired."""
    from longbow_b200 import _lib, gpu
    rng = np.random.default_rng(123)
    n, dim, k = 20000, 33, 10
    db = rng.random((n, dim), dtype=np.float32)
    base = db[11].copy() + np.float32(3.0)  # far from the origin: |q|^2 >> d^2, the old test could never fail
    pos = rng.choice(np.arange(100, n), 200, replace=False)
    for j, r in enumerate(pos):
        v = base.copy()
        v[j % dim] = np.nextafter(v[j % dim], np.float32(np.inf if j % 2 else -np.inf))
        db[r] = v
    q = np.stack([base, db[5]])
    idx = gpu.DenseIndex(dim, np.float32, L2)
    idx.add(db)
    _lib.set_option("dense_scan", 1)
    try:
        gd, gl = idx.search(q, k)
    finally:
        _lib.set_option("dense_scan", 0)
    wd, wl = oracle.search(L2, db, q, k)
    assert_topk_equal(gd, gl, wd, wl, 0.0, "SIMT L2 near ties")
    assert idx.last_uncertified() >= 1, "the near-tie query must be flagged on the SIMT path"
    idx.close()


def test_exchange_world_one(oracle):
    """The exchange with a single rank degenerates to the merge of one record (no peers, no waiting)."""
    import torch
    from longbow_b200 import _lib, gpu
    from longbow_b200.shard import record_layout
    rng = np.random.default_rng(8)
    n, dim, nq, k = 4000, 64, 9, 10
    db, q = make_db(rng, n, dim, np.float32), make_db(rng, nq, dim, np.float32)
    idx = gpu.DenseIndex(dim, np.float32, L2)
    idx.add(db)
    lib = _lib.load()
    ex = C.c_void_p()
    _lib.check(lib.lb_exchange_create(0, 0, 1, record_layout(nq, k)[1], C.byref(ex)))
    pd, pl = C.c_void_p(), C.c_void_p()
    od = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    ol = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    dq = torch.from_numpy(q).cuda()
    for _ in range(3):  # both parities
        _lib.check(lib.lb_exchange_slot(ex, nq, k, C.byref(pd), C.byref(pl)))
        idx.search_device(dq, k, pd.value, pl.value)
        _lib.check(lib.lb_exchange_all_gather_merge(ex, nq, k, k, od.data_ptr(), ol.data_ptr(),
                                                    torch.cuda.current_stream().cuda_stream))
    _lib.check(lib.lb_exchange_error(ex))
    wd, wl = oracle.search(L2, db, q, k)
    assert_topk_equal(od.cpu().numpy(), ol.cpu().numpy(), wd, wl, 0.0, "world-1 exchange")
    lib.lb_exchange_free(ex)
    idx.close()


@pytest.mark.skipif("_ngpu() < 2")
def test_exchange_two_devices_one_process(oracle):
    """Two GPUs driven from ONE process (the Go host's situation): peer access, no IPC, no NCCL."""
    import torch
    from longbow_b200 import _lib, gpu
    from longbow_b200.shard import record_layout, shard_range
    rng = np.random.default_rng(9)
    n, dim, nq, k, world = 30011, 128, 50, 10, 2
    db, q = make_db(rng, n, dim, np.float16), make_db(rng, nq, dim, np.float16)
    db[100] = db[20000]  # an exact tie across the shards
    lib = _lib.load()
    exs, idxs = [], []
    for r in range(world):
        ex = C.c_void_p()
        _lib.check(lib.lb_exchange_create(r, r, world, record_layout(nq, k)[1], C.byref(ex)))
        exs.append(ex)
        lo, hi = shard_range(n, r, world)
        ix = gpu.DenseIndex(dim, np.float16, COS, r)
        ix.add(db[lo:hi])
        ix.set_id_base(lo)
        idxs.append(ix)
    arr = (C.c_void_p * world)(*[e.value for e in exs])
    for r in range(world):
        _lib.check(lib.lb_exchange_connect_local(exs[r], arr))
    outs = []
    for it in range(4):
        outs = []
        for r in range(world):
            with torch.cuda.device(r):
                dq = torch.from_numpy(q).cuda(r)
                pd, pl = C.c_void_p(), C.c_void_p()
                _lib.check(lib.lb_exchange_slot(exs[r], nq, k, C.byref(pd), C.byref(pl)))
                idxs[r].search_device(dq, k, pd.value, pl.value)
                od = torch.empty((nq, k), dtype=torch.float32, device=f"cuda:{r}")
                ol = torch.empty((nq, k), dtype=torch.int64, device=f"cuda:{r}")
                _lib.check(lib.lb_exchange_all_gather_merge(exs[r], nq, k, k, od.data_ptr(), ol.data_ptr(),
                                                            torch.cuda.current_stream(r).cuda_stream))
                outs.append((od, ol))
    wd, wl = oracle.search(COS, db, q, k)
    for r in range(world):
        with torch.cuda.device(r):
            torch.cuda.synchronize()
        _lib.check(lib.lb_exchange_error(exs[r]))
        assert_topk_equal(outs[r][0].cpu().numpy(), outs[r][1].cpu().numpy(), wd, wl, 0.0, f"rank {r}")
    for r in range(world):
        lib.lb_exchange_free(exs[r])
        idxs[r].close()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from longbow_b200.shard import ShardedIndex
    from oracle import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    ok = True
    try:
        rng = np.random.default_rng(7)
        n, dim, nq, k = 60013, 128, 70, 10  # odd n: uneven shards
        db = make_db(rng, n, dim, np.float16)
        db[100] = db[40000]  # an exact tie across shards
        qs = [make_db(rng, nq, dim, np.float16) for _ in range(6)]
        want = [oracle.search(oracle.COSINE, db, q, k) for q in qs]
        for mode in ("p2p", "nccl"):
            sidx = ShardedIndex(dim, np.float16, 1, n, rank, world, rank, exchange=mode)
            sidx.add_local(db[sidx.lo:sidx.hi])
            for overlap in (False, True):
                outs = []
                for q in qs:
                    od = torch.empty((nq, k), dtype=torch.float32, device=dev)
                    ol = torch.empty((nq, k), dtype=torch.int64, device=dev)
                    sidx.search_device(torch.from_numpy(q).to(dev), k, od, ol, overlap=overlap)
                    outs.append((od, ol))
                sidx.wait()
                torch.cuda.synchronize()
                sidx.check_exchange()
                for (od, ol), (wd, wl) in zip(outs, want):
                    ok = ok and np.array_equal(ol.cpu().numpy(), wl) and np.array_equal(od.cpu().numpy(), wd)
            ok = ok and sidx.uncertified() == 0
            dist.barrier()
            sidx.close()
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif("_ngpu() < 2")
def test_sharded_index_two_ranks_nccl_and_p2p():
    """ShardedIndex over two processes (NCCL process group): both exchange implementations, ordered and
    overlapped, every rank's merged answer bit-equal to the oracle's single-index answer."""
    import torch.multiprocessing as mp
    port = _free_port()
    ctx = mp.get_context("spawn")
    mgr = ctx.Manager()
    ret = mgr.dict()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert ret.get(0) and ret.get(1)


def _shard_case(oracle, devices, dtype, metric, seed):
    from longbow_b200 import gpu
    from tests.util import random_bitmap
    rng = np.random.default_rng(seed)
    n, dim, nq, k = 50021, 128, 37, 10
    db, q = make_db(rng, n, dim, dtype), make_db(rng, nq, dim, dtype)
    db[100] = db[40000]  # exact tie across shards: (distance, label) order after the merge
    sh = gpu.ShardedDenseIndex(devices, dim, dtype, metric, n)
    for lo in range(0, n, 7001):  # blocks that straddle shard boundaries
        sh.add(db[lo:lo + 7001])
    assert len(sh) == n and sh.rows_per_shard() % 64 == 0
    gd, gl = sh.search(q, k)
    wd, wl = oracle.search(metric, db, q, k)
    assert_topk_equal(gd, gl, wd, wl, 0.0, f"sharded {devices}")
    tomb, allow = random_bitmap(rng, n, 0.05), random_bitmap(rng, n, 0.3)
    sh.set_tombstones(tomb)
    gd, gl = sh.search(q, k, allow=allow)
    wd, wl = oracle.search(metric, db, q, k, tomb=gpu.pack_bitmap(tomb), allow=gpu.pack_bitmap(allow))
    assert_topk_equal(gd, gl, wd, wl, 0.0, f"sharded + bitmaps {devices}")
    sh.close()


@pytest.mark.parametrize("dtype,metric", [(np.float16, COS), (np.float32, L2), (np.int8, DOT)])
def test_shard_handle_two_shards_one_device(oracle, dtype, metric):
    """lb_shard_* with both shards on device 0: the whole C-ABI path (range split, per-shard bitmaps slices, records
    stored into the root's gather buffer, event-ordered merge) on a one-GPU box."""
    _shard_case(oracle, [0, 0], dtype, metric, 11)


def test_shard_handle_three_uneven_shards_one_device(oracle):
    _shard_case(oracle, [0, 0, 0], np.float16, L2, 12)


@pytest.mark.skipif("_ngpu() < 2")
def test_shard_handle_two_devices(oracle):
    """Two real GPUs from one process: the re-score kernel of GPU 1 stores its record into GPU 0's memory over the
    NVLink peer mapping."""
    _shard_case(oracle, [0, 1], np.float16, COS, 13)
    _shard_case(oracle, [0, 1], np.int8, DOT, 14)


def test_shard_handle_near_ties_are_repaired(oracle):
    """A near-tie cluster larger than the coarse margin, spread over both shards: flagged, repaired exhaustively."""
    from longbow_b200 import gpu
    rng = np.random.default_rng(99)
    n, dim, k = 30000, 256, 100
    db = make_db(rng, n, dim, np.float16)
    base = db[17].copy()
    pos = rng.choice(np.arange(1000, n), 300, replace=False)
    for j, r in enumerate(pos):
        v = base.copy()
        v[j % dim] = np.nextafter(v[j % dim], np.float16(np.inf if j % 2 else -np.inf), dtype=np.float16)
        db[r] = v
    q = np.stack([base, db[12345]])
    sh = gpu.ShardedDenseIndex([0, 0], dim, np.float16, COS, n)
    sh.add(db)
    gd, gl = sh.search(q, k)
    wd, wl = oracle.search(COS, db, q, k)
    assert_topk_equal(gd, gl, wd, wl, 0.0, "sharded near ties")
    assert sh.last_uncertified() >= 1
    sh.close()


def test_shard_handle_concurrent_searches(oracle):
    """Four threads search ONE lb_shard handle at once, as goroutines holding the Go shim's read lock would
    (go/longbow_b200.go SearchBatch): the handle serialises them on its own mutex (its streams and the root's gather
    buffer belong to one search at a time), every caller gets its own exact answer."""
    import threading
    from longbow_b200 import gpu
    rng = np.random.default_rng(21)
    n, dim, k = 40003, 128, 10
    db = make_db(rng, n, dim, np.float16)
    sh = gpu.ShardedDenseIndex([0, 0], dim, np.float16, COS, n)
    sh.add(db)
    batches = [make_db(rng, nq, dim, np.float16) for nq in (33, 5, 64, 17)]  # different record sizes: the gather buffer regrows
    want = [oracle.search(COS, db, q, k) for q in batches]
    got = [None] * len(batches)
    errs = []

    def worker(i):
        try:
            for _ in range(6):
                got[i] = sh.search(batches[i], k)
                assert_topk_equal(got[i][0], got[i][1], want[i][0], want[i][1], 0.0, f"concurrent caller {i}")
        except BaseException as e:  # noqa: BLE001 -- reported by the main thread
            errs.append(e)

    ts = [threading.Thread(target=worker, args=(i,)) for i in range(len(batches))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs[0]
    sh.close()
