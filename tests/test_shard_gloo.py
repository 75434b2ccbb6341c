"""Row-sharded search across ranks: the N>1 host logic (shard ranges, id bases, all-gather layout, merge)
on CPU with world_size 2 over gloo.  Each rank computes its shard's top-k with the CPU oracle (this is the
test's stand-in for the per-GPU search), all-gathers the lists exactly as longbow_b200.shard does, merges
them with the oracle's merge, and rank 0 compares with the single-shard answer."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


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
    from longbow_b200.shard import gather_layout, shard_range
    from oracle import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)
        n, dim, nq, k = 5003, 48, 11, 10  # odd n: uneven shards
        db = rng.random((n, dim), dtype=np.float32)
        db[100] = db[4000]  # an exact tie across shards: (distance, id) order must hold after the merge
        q = rng.random((nq, dim), dtype=np.float32)
        lo, hi = shard_range(n, rank, world)
        d, l = oracle.search(oracle.L2, db[lo:hi], q, k, id_base=lo)
        gd, gl = gather_layout(torch.from_numpy(d), torch.from_numpy(l), world,
                               lambda out, inp: dist.all_gather_into_tensor(out, inp))
        md, ml = oracle.merge(gd.numpy(), gl.numpy(), k)
        wd, wl = oracle.search(oracle.L2, db, q, k)
        ok = np.array_equal(ml, wl) and np.array_equal(md, wd)
        covered = torch.tensor([hi - lo])
        dist.all_reduce(covered)
        ret[rank] = bool(ok and int(covered.item()) == n)
    finally:
        dist.destroy_process_group()


def test_two_rank_shard_gather_merge():
    import torch.multiprocessing as mp
    port = _free_port()
    ctx = mp.get_context("spawn")
    mgr = ctx.Manager()
    ret = mgr.dict()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert ret.get(0) and ret.get(1)


def test_shard_range_partition():
    from longbow_b200.shard import shard_range
    for n in (0, 1, 7, 100, 1_000_003):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _packed_worker(rank, world, port, ret):
    """The exchange as ShardedIndex(exchange="nccl") really does it: ONE all-gather of the packed record
    [nq*k f32 | pad to 16 | nq*k i64] per rank (shard.record_layout), parsed back with the record / label strides the
    merge kernel is given (lb_merge_topk_packed_device)."""
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from longbow_b200.shard import record_layout, shard_range
    from oracle import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(11)
        n, dim, nq, k = 4099, 32, 7, 9          # nq*k*4 = 252: the label block needs the 16-byte pad
        db = rng.integers(-128, 128, (n, dim), dtype=np.int8)
        db[5] = db[4000]                          # exact tie across shards
        q = rng.integers(-128, 128, (nq, dim), dtype=np.int8)
        lo, hi = shard_range(n, rank, world)
        d, l = oracle.search(oracle.DOT, db[lo:hi], q, k, id_base=lo)
        loff, rec = record_layout(nq, k)
        assert loff % 16 == 0 and loff >= nq * k * 4 and rec == loff + nq * k * 8
        local = torch.zeros(rec, dtype=torch.uint8)
        local[:nq * k * 4] = torch.from_numpy(d.reshape(-1).view(np.uint8))
        local[loff:] = torch.from_numpy(l.reshape(-1).view(np.uint8))
        gathered = torch.empty(world * rec, dtype=torch.uint8)
        dist.all_gather_into_tensor(gathered, local)
        g = gathered.numpy()
        gd = np.stack([g[p * rec:p * rec + nq * k * 4].view(np.float32).reshape(nq, k) for p in range(world)])
        gl = np.stack([g[p * rec + loff:p * rec + loff + nq * k * 8].view(np.int64).reshape(nq, k) for p in range(world)])
        md, ml = oracle.merge(gd, gl, k)
        wd, wl = oracle.search(oracle.DOT, db, q, k)
        ret[rank] = bool(np.array_equal(ml, wl) and np.array_equal(md, wd))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_packed_record_exchange(world):
    import torch.multiprocessing as mp
    port = _free_port()
    ctx = mp.get_context("spawn")
    mgr = ctx.Manager()
    ret = mgr.dict()
    procs = [ctx.Process(target=_packed_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert all(ret.get(r) for r in range(world))
