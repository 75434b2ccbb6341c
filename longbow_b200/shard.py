"""Row-sharded multi-GPU search: one process per GPU (torch.distributed), contiguous row ranges,
queries replicated, ONE exchange step per batch -- an all-gather of the per-GPU top-k records over
NVLink followed by a (distance, id) merge on every rank.

This is the B200 replacement for ShardedHNSW's fan-out + concat + sort
(internal/store/sharded_hnsw.go:378-503) and the mesh-level MergeSortedStreams
(internal/store/result_merger.go:34-100).  Merging exact per-shard top-k lists is exact, so no
k*2 oversampling (sharded_hnsw.go:421) is needed.

Two exchange implementations behind the same call:

* ``"p2p"`` (default): liblongbow_b200's own kernels over peer memory (csrc/exchange.cu).  The local search
  writes its record straight into this rank's receive slot, a push kernel stores it into every peer's slot
  over NVLink, a release-store signals the batch number, and the merge kernel waits for all records before
  merging.  torch.distributed only ships the 64-byte CUDA IPC handles once, at set-up.
* ``"nccl"``: one ``all_gather_into_tensor`` of the packed record ([distances | labels] bytes) per batch, then
  the same merge kernel (``lb_merge_topk_packed_device``).

With ``overlap=True`` the exchange of batch i runs on ONE side stream, in batch order, while the scan of batch
i+1 runs on the caller's stream.
"""
from __future__ import annotations

import ctypes as C

import numpy as np


def shard_range(n_rows: int, rank: int, world: int):
    """Rows [lo, hi) owned by `rank`: contiguous, sizes differ by at most one."""
    base, rem = divmod(n_rows, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def record_layout(nq: int, k: int):
    """(label_offset, record_bytes) of one rank's exchange record: [nq*k f32 | pad to 16 | nq*k i64]."""
    loff = (nq * k * 4 + 15) & ~15
    return loff, loff + nq * k * 8


def gather_layout(local_d, local_l, world: int, all_gather):
    """All-gather [nq, k] lists into [world, nq, k] with the supplied collective (gloo tests)."""
    import torch
    gd = torch.empty((world,) + tuple(local_d.shape), dtype=local_d.dtype, device=local_d.device)
    gl = torch.empty((world,) + tuple(local_l.shape), dtype=local_l.dtype, device=local_l.device)
    all_gather(gd.view((-1,) + tuple(local_d.shape[1:])), local_d)
    all_gather(gl.view((-1,) + tuple(local_l.shape[1:])), local_l)
    return gd, gl


class ShardedIndex:
    """Each rank owns one `gpu.DenseIndex` over its row range; `search_device` returns the global top-k."""

    def __init__(self, dim, dtype, metric, n_rows_total: int, rank: int, world: int, device: int, group=None,
                 exchange: str = "p2p"):
        from . import gpu
        self.rank, self.world, self.group, self.device = rank, world, group, device
        self.lo, self.hi = shard_range(n_rows_total, rank, world)
        self.index = gpu.DenseIndex(dim, dtype, metric, device)
        self.index.set_id_base(self.lo)
        self.exchange = exchange
        self._ex = None          # lb_exchange handle (p2p)
        self._ex_bytes = 0
        self._comm = None        # side stream of the overlapped exchange
        self._nccl = {}          # (nq, k) -> two sets of packed staging buffers (nccl mode)
        self._turn = 0
        self._done = [None, None]  # per parity: event after the exchange that last used that staging slot
        self._last_done = None
        self._uncert = None      # device u32: queries whose coarse stage could not be certified (accumulates)

    # ------------------------------------------------------------------ data
    def add_local_device(self, tensor):
        self.index.add_device(tensor)

    def add_local(self, rows: np.ndarray):
        self.index.add(rows)

    # ------------------------------------------------------------------ exchange set-up
    def _lib(self):
        from . import _lib
        return _lib

    def _ensure_p2p(self, nq, k, dev):
        """Create the peer-memory exchange (once; re-created if a larger record is needed)."""
        import torch
        import torch.distributed as dist
        _lib = self._lib()
        _, need = record_layout(nq, k)
        if self._ex is not None and need <= self._ex_bytes:
            return
        if self._ex is not None:
            torch.cuda.synchronize()
            dist.barrier(group=self.group)
            _lib.load().lb_exchange_free(self._ex)
            self._ex = None
        lib = _lib.load()
        h = C.c_void_p()
        _lib.check(lib.lb_exchange_create(self.device, self.rank, self.world, need, C.byref(h)))
        mine = (C.c_ubyte * 64)()
        _lib.check(lib.lb_exchange_handle(h, mine))
        # ship the 64-byte IPC handles through the process group (set-up only)
        t = torch.tensor(list(bytes(mine)), dtype=torch.uint8, device=dev)
        allh = torch.empty((self.world, 64), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allh.view(-1), t, group=self.group)
        blob = bytes(allh.cpu().numpy().tobytes())
        buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        _lib.check(lib.lb_exchange_connect_ipc(h, buf))
        dist.barrier(group=self.group)  # every rank has mapped every peer before the first push
        self._ex, self._ex_bytes = h, need

    def _nccl_slot(self, nq, k, dev, parity):
        import torch
        key = (nq, k)
        if key not in self._nccl:
            loff, rec = record_layout(nq, k)
            self._nccl[key] = [{"local": torch.empty(rec, dtype=torch.uint8, device=dev),
                                "all": torch.empty(self.world * rec, dtype=torch.uint8, device=dev)} for _ in range(2)]
        return self._nccl[key][parity]

    # ------------------------------------------------------------------ search
    def search_device(self, q, k: int, out_d, out_l, allow=None, overlap: bool = False, certify: bool = True):
        """q [nq,dim], out_d [nq,k] f32, out_l [nq,k] i64: device tensors; asynchronous.

        overlap=False: everything is ordered on the current stream.  overlap=True: the local scan runs on the
        current stream, the exchange + merge on the side stream; call wait() before reading out_d / out_l.
        certify=True: the local searches report queries whose coarse-stage margin did not cover the error
        bound into a device counter (``uncertified()`` reads it; ``repair()`` redoes them exactly)."""
        import torch
        import torch.distributed as dist
        _lib = self._lib()
        lib = _lib.load()
        dev = q.device
        if certify and self._uncert is None:
            self._uncert = torch.zeros(1, dtype=torch.int32, device=dev)
        cnt = self._uncert if certify else None
        if self.world == 1:
            self.index.search_device(q, k, out_d, out_l, allow=allow, uncert_count=cnt)
            return
        nq = q.shape[0]
        cur = torch.cuda.current_stream()
        self._turn ^= 1
        parity = self._turn
        if self._done[parity] is not None:
            cur.wait_event(self._done[parity])  # the exchange that last read this parity's local record is done
        loff, rec = record_layout(nq, k)
        if self.exchange == "p2p":
            self._ensure_p2p(nq, k, dev)
            pd, pl = C.c_void_p(), C.c_void_p()
            _lib.check(lib.lb_exchange_slot(self._ex, nq, k, C.byref(pd), C.byref(pl)))
            self.index.search_device(q, k, pd.value, pl.value, allow=allow, uncert_count=cnt)
        else:
            s = self._nccl_slot(nq, k, dev, parity)
            base = s["local"].data_ptr()
            self.index.search_device(q, k, base, base + loff, allow=allow, uncert_count=cnt)
        if overlap:
            if self._comm is None:
                self._comm = torch.cuda.Stream(device=dev)
            scanned = torch.cuda.Event()
            scanned.record(cur)
            ex_stream = self._comm
            ex_stream.wait_event(scanned)
        else:
            ex_stream = cur
        with torch.cuda.stream(ex_stream):
            if self.exchange == "p2p":
                _lib.check(lib.lb_exchange_all_gather_merge(self._ex, nq, k, k, out_d.data_ptr(), out_l.data_ptr(),
                                                            ex_stream.cuda_stream))
            else:
                dist.all_gather_into_tensor(s["all"], s["local"], group=self.group)
                _lib.check(lib.lb_merge_topk_packed_device(self.device, s["all"].data_ptr(), rec, loff, self.world, nq,
                                                           k, k, out_d.data_ptr(), out_l.data_ptr(),
                                                           ex_stream.cuda_stream))
            done = torch.cuda.Event()
            done.record(ex_stream)
        self._done[parity] = done
        self._last_done = done

    def wait(self):
        """Make the current stream wait for every exchange issued so far (overlap=True)."""
        import torch
        if self._last_done is not None:
            torch.cuda.current_stream().wait_event(self._last_done)

    def uncertified(self) -> int:
        """Queries (summed over this rank's searches so far) whose coarse stage could not be certified.
        Synchronises."""
        return 0 if self._uncert is None else int(self._uncert.item())

    def check_exchange(self):
        """Raise if a merge ever gave up waiting for a peer's record (p2p mode).  Synchronises."""
        if self._ex is not None:
            _lib = self._lib()
            _lib.check(_lib.load().lb_exchange_error(self._ex))

    def close(self):
        if self._ex is not None:
            self._lib().load().lb_exchange_free(self._ex)
            self._ex = None
        self.index.close()
