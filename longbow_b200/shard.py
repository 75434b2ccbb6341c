"""Row-sharded multi-GPU search: one process per GPU (torch.distributed), contiguous row ranges,
queries replicated, ONE exchange step -- an all-gather of the per-GPU top-k lists over
NCCL/NVLink -- followed by a (distance, id) merge on every rank.

This is the B200 replacement for ShardedHNSW's fan-out + concat + sort
(internal/store/sharded_hnsw.go:378-503) and the mesh-level MergeSortedStreams
(internal/store/result_merger.go:34-100).  Merging exact per-shard top-k lists is exact, so no
k*2 oversampling (sharded_hnsw.go:421) is needed.
"""
from __future__ import annotations

import numpy as np


def shard_range(n_rows: int, rank: int, world: int):
    """Rows [lo, hi) owned by `rank`: contiguous, sizes differ by at most one."""
    base, rem = divmod(n_rows, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_layout(local_d, local_l, world: int, all_gather):
    """All-gather [nq, k] lists into [world, nq, k] with the supplied collective (NCCL or gloo)."""
    import torch
    gd = torch.empty((world,) + tuple(local_d.shape), dtype=local_d.dtype, device=local_d.device)
    gl = torch.empty((world,) + tuple(local_l.shape), dtype=local_l.dtype, device=local_l.device)
    # the concatenated [world * nq, k] view is the one layout every backend (NCCL, gloo) accepts
    all_gather(gd.view((-1,) + tuple(local_d.shape[1:])), local_d)
    all_gather(gl.view((-1,) + tuple(local_l.shape[1:])), local_l)
    return gd, gl


class ShardedIndex:
    """Each rank owns one `gpu.DenseIndex` over its row range; `search` returns the global top-k.

    The exchange (all-gather + merge) runs on its own CUDA stream, so with ``overlap=True`` the exchange of
    batch i overlaps the scan of batch i+1 (results are valid after ``wait()``)."""

    def __init__(self, dim, dtype, metric, n_rows_total: int, rank: int, world: int, device: int, group=None):
        from . import gpu
        self.rank, self.world, self.group, self.device = rank, world, group, device
        self.lo, self.hi = shard_range(n_rows_total, rank, world)
        self.index = gpu.DenseIndex(dim, dtype, metric, device)
        self.index.set_id_base(self.lo)
        self._comm = None      # exchange stream
        self._slots = {}       # (nq, k) -> two sets of staging buffers + their "exchange done" events
        self._turn = 0
        self._last_done = None

    def add_local_device(self, tensor):
        self.index.add_device(tensor)

    def add_local(self, rows: np.ndarray):
        self.index.add(rows)

    def _slot(self, nq, k, dev):
        import torch
        key = (nq, k)
        if key not in self._slots:
            sets = []
            for _ in range(2):
                sets.append({
                    "ld": torch.empty((nq, k), dtype=torch.float32, device=dev),
                    "ll": torch.empty((nq, k), dtype=torch.int64, device=dev),
                    "gd": torch.empty((self.world, nq, k), dtype=torch.float32, device=dev),
                    "gl": torch.empty((self.world, nq, k), dtype=torch.int64, device=dev),
                    "done": None})
            self._slots[key] = sets
        self._turn ^= 1
        return self._slots[key][self._turn]

    def search_device(self, q, k: int, out_d, out_l, allow=None, overlap: bool = False):
        """q, out_d [nq,k] f32, out_l [nq,k] i64: device tensors; asynchronous.

        overlap=False: everything is ordered on the current stream.  overlap=True: the local scan runs on the
        current stream, the exchange on the side stream; call wait() before reading out_d / out_l."""
        import torch
        import torch.distributed as dist
        from . import _lib
        if self.world == 1:
            self.index.search_device(q, k, out_d, out_l, allow=allow)
            return
        nq = q.shape[0]
        if not overlap:
            # plain ordered form: local search, all-gather, merge -- all on the current stream
            ld = torch.empty((nq, k), dtype=torch.float32, device=q.device)
            ll = torch.empty((nq, k), dtype=torch.int64, device=q.device)
            self.index.search_device(q, k, ld, ll, allow=allow)
            gd, gl = gather_layout(ld, ll, self.world,
                                   lambda out, inp: dist.all_gather_into_tensor(out, inp, group=self.group))
            _lib.check(_lib.load().lb_merge_topk_device(self.device, gd.data_ptr(), gl.data_ptr(), self.world, nq, k, k,
                                                        out_d.data_ptr(), out_l.data_ptr(),
                                                        torch.cuda.current_stream().cuda_stream))
            return
        # experimental: exchange on a side stream (verified on 2 GPUs only)
        cur = torch.cuda.current_stream()
        s = self._slot(nq, k, q.device)
        if s["done"] is not None:
            cur.wait_event(s["done"])  # the exchange that last used these staging buffers has finished
        self.index.search_device(q, k, s["ld"], s["ll"], allow=allow)
        if self._comm is None:
            self._comm = torch.cuda.Stream(device=q.device)
        scanned = torch.cuda.Event()
        scanned.record(cur)
        comm = self._comm
        with torch.cuda.stream(comm):
            comm.wait_event(scanned)
            dist.all_gather_into_tensor(s["gd"].view(-1, k), s["ld"], group=self.group)
            dist.all_gather_into_tensor(s["gl"].view(-1, k), s["ll"], group=self.group)
            _lib.check(_lib.load().lb_merge_topk_device(self.device, s["gd"].data_ptr(), s["gl"].data_ptr(), self.world,
                                                        nq, k, k, out_d.data_ptr(), out_l.data_ptr(), comm.cuda_stream))
            done = torch.cuda.Event()
            done.record(comm)
        s["done"] = done
        self._last_done = done

    def wait(self):
        """Make the current stream wait for every exchange issued so far (overlap=True)."""
        import torch
        if self._last_done is not None:
            torch.cuda.current_stream().wait_event(self._last_done)

    def close(self):
        self.index.close()
