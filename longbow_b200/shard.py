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
    """Each rank owns one `gpu.DenseIndex` over its row range; `search` returns the global top-k."""

    def __init__(self, dim, dtype, metric, n_rows_total: int, rank: int, world: int, device: int, group=None):
        from . import gpu
        self.rank, self.world, self.group, self.device = rank, world, group, device
        self.lo, self.hi = shard_range(n_rows_total, rank, world)
        self.index = gpu.DenseIndex(dim, dtype, metric, device)
        self.index.set_id_base(self.lo)

    def add_local_device(self, tensor):
        self.index.add_device(tensor)

    def add_local(self, rows: np.ndarray):
        self.index.add(rows)

    def search_device(self, q, k: int, out_d, out_l, allow=None):
        """q, out_d [nq,k] f32, out_l [nq,k] i64: device tensors; asynchronous on the current stream."""
        import torch
        import torch.distributed as dist
        from . import _lib
        if self.world == 1:
            self.index.search_device(q, k, out_d, out_l, allow=allow)
            return
        nq = q.shape[0]
        ld = torch.empty((nq, k), dtype=torch.float32, device=q.device)
        ll = torch.empty((nq, k), dtype=torch.int64, device=q.device)
        self.index.search_device(q, k, ld, ll, allow=allow)
        gd, gl = gather_layout(ld, ll, self.world,
                               lambda out, inp: dist.all_gather_into_tensor(out, inp, group=self.group))
        _lib.check(_lib.load().lb_merge_topk_device(self.device, gd.data_ptr(), gl.data_ptr(), self.world, nq, k, k,
                                                    out_d.data_ptr(), out_l.data_ptr(),
                                                    torch.cuda.current_stream().cuda_stream))

    def close(self):
        self.index.close()
