#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (contract in the task statement, section 4).

Workload (BASELINE.json configs[1], "C2"): brute-force cosine k=100 over 1M x 768 fp16 unit-norm
embeddings, query batch 1024.  A step = one batch of 1024 queries answered (distance + top-k +
exact re-score).  N>1: the SAME database is row-sharded over the ranks (strong scaling), per-GPU
top-k lists are all-gathered over NCCL and merged on every rank.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os

# Libraries (NCCL's version banner, torch warnings) write to stdout; the contract is ONE JSON line there.  File
# descriptor 1 is pointed at stderr for the whole run and restored just before the result is printed.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(obj):
    sys.stdout.flush()
    os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(obj), flush=True)
    os.dup2(2, 1)
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ROWS, DIM, NQ, K = 1_000_000, 768, 1024, 100
METRIC_NAME = "k-NN QPS (cosine, k=100, 1M x 768 fp16, query batch 1024)"


def make_data(n_rows, nq, n_batches, device=None):
    """Seeded synthetic inputs: N(0,1) -> L2-normalised -> fp16; three all-zero
    rows exercise the cosine '== 0 -> 1.0' rule.  Generated with torch on the CPU so both arms see
    identical bytes."""
    import torch
    g = torch.Generator().manual_seed(2001)
    db = torch.empty((n_rows, DIM), dtype=torch.float16)
    chunk = 65536
    for lo in range(0, n_rows, chunk):
        x = torch.randn((min(chunk, n_rows - lo), DIM), generator=g)
        x /= x.norm(dim=1, keepdim=True)
        db[lo:lo + x.shape[0]] = x.half()
    for r in (3, n_rows // 2, n_rows - 1):
        db[r] = 0
    g2 = torch.Generator().manual_seed(2002)
    qs = torch.randn((n_batches, nq, DIM), generator=g2)
    qs /= qs.norm(dim=2, keepdim=True)
    return db, qs.half()


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, dev_index):
        super().__init__(daemon=True)
        self.dev_index, self.stop_flag = dev_index, False
        self.sm, self.reasons, self.max_mhz = [], set(), None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.dev_index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                     "sw_thermal_slowdown": 0x20, "hw_power_brake": 0x80, "sync_boost": 0x10}
            while not self.stop_flag:
                util = nv.nvmlDeviceGetUtilizationRates(h).gpu
                mhz = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.sm.append((mhz, util, time.time(), r))
                for n, bit in names.items():
                    if r & bit:
                        self.reasons.add(n)
                time.sleep(0.01)
        except Exception as e:  # NVML missing: report nothing rather than die
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    NAMES = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
             "sw_thermal_slowdown": 0x20, "hw_power_brake": 0x80, "sync_boost": 0x10}

    def summary(self):
        vals = [s[0] for s in self.sm]
        return {"sm_mhz": statistics.median(vals) if vals else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(vals)}

    def window(self, t0, t1):
        """Clock record of one config's timed region [t0, t1] (wall clock)."""
        sel = [s for s in self.sm if t0 <= s[2] <= t1]
        reasons = set()
        for s in sel:
            for n, bit in self.NAMES.items():
                if s[3] & bit:
                    reasons.add(n)
        vals = [s[0] for s in sel]
        return {"sm_mhz": statistics.median(vals) if vals else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(reasons), "samples": len(vals)}


def cpu_baseline(db_np, q_np, k):
    """The reference's CPU path (oracle port, AVX-512 + OpenMP, all host threads) on a bounded
    sample of the same workload: a few queries per core against the full database."""
    from oracle import oracle
    oracle.build()
    cores = oracle.fast_use_all_cores()  # not OMP_NUM_THREADS: torchrun sets it to 1
    oracle.search(oracle.COSINE, db_np, q_np[:cores], k, impl="fast")  # warm-up (page in the DB)
    nq = min(q_np.shape[0], 4 * cores)
    t0 = time.perf_counter()
    oracle.search(oracle.COSINE, db_np, q_np[:nq], k, impl="fast")
    dt = time.perf_counter() - t0
    return {"value": nq / dt, "unit": "queries/s", "cores": cores, "kind": "port",
            "sample": f"{nq} queries x full {db_np.shape[0]} x {DIM} fp16 DB, k={k}, {oracle.fast_isa()} + OpenMP, "
                      f"{dt:.2f} s wall"}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (C restatement of its
    SIMD kernels + heap; Go is not installable here) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle
    oracle.build()
    cores = oracle.fast_use_all_cores()  # not OMP_NUM_THREADS: torchrun sets it to 1
    db, qs = make_data(N_ROWS, NQ, 1)
    db_np = db.numpy()
    q_np = qs[0].numpy()
    nq = max(8, cores)  # one bounded sample per step
    for _ in range(args.warmup):
        oracle.search(oracle.COSINE, db_np, q_np[:nq], K, impl="fast")
    t0 = time.perf_counter()
    for s in range(args.steps):
        lo = (s * nq) % (NQ - nq + 1)
        oracle.search(oracle.COSINE, db_np, q_np[lo:lo + nq], K, impl="fast")
    dt = time.perf_counter() - t0
    qps = nq * args.steps / dt
    sample = f"{nq} queries/step x full {N_ROWS} x {DIM} fp16 DB, k={K}, {oracle.fast_isa()} + OpenMP"
    emit({
        "impl": "reference", "metric": METRIC_NAME, "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        # the same workload (and workload string) as the GPU arm; what differs is stated in `note` / `sample`
        "config": {"workload": "C2: brute-force cosine k=100, 1M x 768 fp16 unit-norm embeddings, query batch 1024",
                   "rows": N_ROWS, "dim": DIM, "queries_per_step": NQ, "k": K,
                   "note": "CPU restatement of the reference's SIMD path (Go toolchain absent); each step is a bounded "
                           f"sample of the batch ({nq} of its {NQ} queries against the full database)"},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--rows", type=int, default=N_ROWS, help="debug only: a smaller DB invalidates the number")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--configs", default="all", help="extra BASELINE configs in the `configs` array: all | none | C1,C3,...")
    args = ap.parse_args()
    wd = int(os.environ.get("LB_WATCHDOG", "0"))
    if wd > 0:  # debug: dump all Python stacks and exit if the run wedges
        import faulthandler
        faulthandler.dump_traceback_later(wd, exit=True)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from longbow_b200 import _lib
    from longbow_b200.shard import ShardedIndex, shard_range

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    for name_, val_ in os.environ.items():  # debug: LB_OPT_<option>=<int> -> lb_set_option (A/B experiments)
        if name_.startswith("LB_OPT_"):
            _lib.set_option(name_[7:], int(val_))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    warm = max(args.warmup, 3)
    steps = args.steps
    n_rows = args.rows

    def log(msg):
        if os.environ.get("LB_VERBOSE"):
            print(f"[bench r{rank}] {msg}", file=sys.stderr, flush=True)

    db, qs = make_data(n_rows, NQ, warm + steps)
    log("data ready")
    lo, hi = shard_range(n_rows, rank, world)
    sidx = ShardedIndex(DIM, np.float16, _lib.METRIC_COSINE, n_rows, rank, world, local)
    sidx.index.reserve(hi - lo)
    sidx.add_local_device(db[lo:hi].to(dev))
    d_qs = qs.to(dev)
    out_d = torch.empty((NQ, K), dtype=torch.float32, device=dev)
    out_l = torch.empty((NQ, K), dtype=torch.int64, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm ("value").  Independent batches alternate between two CUDA streams so the short
    # tail kernels of batch s (selection, merge, exact re-score, NCCL exchange) overlap the scan of batch s + 1;
    # every batch is still one complete search (scan + top-k + re-score [+ all-gather + merge]).
    log("index resident")
    # N > 1: the exchange (peer-memory push + signal + merge, csrc/exchange.cu) of batch s runs on ONE side stream,
    # in batch order on every rank, under the scan of batch s + 1.
    n_streams = int(os.environ.get("LB_BENCH_STREAMS", "2"))
    overlap = world > 1
    streams = [torch.cuda.Stream(device=dev) for _ in range(n_streams)]
    outs = [(out_d, out_l)] + [(torch.empty_like(out_d), torch.empty_like(out_l)) for _ in range(n_streams - 1)]

    def run_steps(first, count):
        for s in range(first, first + count):
            with torch.cuda.stream(streams[s % n_streams]):
                sidx.search_device(d_qs[s % d_qs.shape[0]], K, outs[s % n_streams][0], outs[s % n_streams][1],
                                   overlap=overlap)

    def join_streams():
        cur = torch.cuda.current_stream()
        for st_ in streams:
            cur.wait_stream(st_)
        sidx.wait()

    run_steps(0, warm)
    join_streams()
    barrier()
    log("warm-up done")
    sampler = ClockSampler(local)
    sampler.start()
    _lib.prof_read(reset=True)
    _lib.prof_enable(True)
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for st_ in streams:
        st_.wait_event(e0)
    run_steps(warm, steps)
    join_streams()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() - l0
    _lib.prof_enable(False)
    scan_ms, scan_n, scan_units = _lib.prof_read(reset=True)
    log(f"timed region done: {ms:.2f} ms")
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        mine = torch.tensor([ms, scan_ms / max(scan_n, 1)], device=dev)  # this rank's own region / scan-kernel time
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"region_ms": [round(float(x[0]), 3) for x in allr],
                    "scan_kernel_ms": [round(float(x[1]), 4) for x in allr]}
        ms = float(t.item())
    check_l = outs[(warm + steps - 1) % n_streams][1].cpu().numpy()  # the last timed batch
    check_d = outs[(warm + steps - 1) % n_streams][0].cpu().numpy()
    if ms < 600.0:
        # timed region too short for the 10 ms NVML sampler: keep the same load running for ~1 s more.  The number
        # of extra batches is derived from the all-reduced time, so every rank issues the same exchanges.
        extra = int(min(20000, max(4, 1000.0 / max(ms / steps, 1e-3))))
        extra += (-extra) % n_streams
        for _ in range(extra // n_streams):
            run_steps(warm + steps, n_streams)
        join_streams()
        torch.cuda.synchronize()
    # ---- the dominant kernel on its own: the same batches on ONE stream.  In the timed region above two streams
    # overlap, and the events bracketing a scan launch then also cover the time it queues behind (and runs beside)
    # the other stream's kernels -- consecutive scans even overlap each other at their edges -- so that bracket is not
    # the kernel's own duration.  Both are reported; the roofline uses this one.
    _lib.prof_read(reset=True)
    _lib.prof_enable(True)
    for s in range(steps):
        sidx.search_device(d_qs[(warm + s) % d_qs.shape[0]], K, outs[0][0], outs[0][1])
    torch.cuda.synchronize()
    _lib.prof_enable(False)
    scan1_ms, scan1_n, scan1_units = _lib.prof_read(reset=True)
    main_clocks = sampler.summary()
    uncertified = sidx.uncertified()
    if world > 1:
        sidx.check_exchange()

    # ---- end-to-end arm: host buffers through the C ABI, H2D + D2H inside the timed region
    e2e = None
    checks = {}  # result cross-checks between the arms (reported, and shouted about on stderr if one fails)
    if world == 1:
        hq = qs.pin_memory().numpy()
        hd = torch.empty((NQ, K), dtype=torch.float32).pin_memory().numpy()
        hl = torch.empty((NQ, K), dtype=torch.int64).pin_memory().numpy()
        for s in range(warm):
            sidx.index.search_into(hq[s], K, hd, hl)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for s in range(steps):
            sidx.index.search_into(hq[warm + s], K, hd, hl)
        dt1 = time.perf_counter() - t0
        checks["host_api_equals_device_api"] = bool(np.array_equal(hl, check_l))
        # The reference serves searches from many goroutines under a read lock (internal/gpu/faiss_gpu.go:108);
        # the C ABI is thread-safe the same way.  Two concurrent callers let one call's PCIe copies overlap the
        # other's kernels.  Every step still does its own H2D of the queries and D2H of the results.
        callers = int(os.environ.get("LB_BENCH_CALLERS", "2"))
        bufs = [(torch.empty((NQ, K), dtype=torch.float32).pin_memory().numpy(),
                 torch.empty((NQ, K), dtype=torch.int64).pin_memory().numpy()) for _ in range(callers)]
        last = [None] * callers

        def worker(c, lo_s, hi_s):
            torch.cuda.set_device(local)
            for s in range(lo_s, hi_s):
                sidx.index.search_into(hq[s], K, bufs[c][0], bufs[c][1])
                last[c] = s

        def run_callers(first, count):
            per = [(first + count * c // callers, first + count * (c + 1) // callers) for c in range(callers)]
            ths = [threading.Thread(target=worker, args=(c, a, b)) for c, (a, b) in enumerate(per)]
            t_ = time.perf_counter()
            for th in ths:
                th.start()
            for th in ths:
                th.join()
            return time.perf_counter() - t_
        run_callers(0, warm)
        dt = run_callers(warm, steps)
        # the last step of caller 0 must equal a fresh single-caller answer for the same batch
        sidx.index.search_into(hq[last[0]], K, hd, hl)
        checks["concurrent_callers_agree"] = bool(np.array_equal(hl, bufs[0][1]) and np.array_equal(hd, bufs[0][0]))
        sidx.index.search_into(hq[0], K, hd, hl)
        c2_first_d, c2_first_l = hd.copy(), hl.copy()
        e2e = {"value": NQ * steps / dt, "unit": "queries/s", "h2d_bytes_per_step": NQ * DIM * 2,
               "d2h_bytes_per_step": NQ * K * 12, "ms_per_step": dt / steps * 1e3, "callers": callers,
               "single_caller_value": NQ * steps / dt1, "single_caller_ms_per_step": dt1 / steps * 1e3}
    else:
        hq = qs.pin_memory()
        hd = torch.empty((NQ, K), dtype=torch.float32).pin_memory()
        hl = torch.empty((NQ, K), dtype=torch.int64).pin_memory()
        dq = torch.empty((NQ, DIM), dtype=torch.float16, device=dev)

        def step(s):
            dq.copy_(hq[s], non_blocking=True)
            sidx.search_device(dq, K, out_d, out_l)  # ordered: this step's result is read back below
            if rank == 0:
                hd.copy_(out_d, non_blocking=True)
                hl.copy_(out_l, non_blocking=True)
            torch.cuda.synchronize()
        for s in range(warm):
            step(s)
        barrier()
        t0 = time.perf_counter()
        for s in range(steps):
            step(warm + s)
        barrier()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        e2e = {"value": NQ * steps / dt, "unit": "queries/s", "h2d_bytes_per_step": NQ * DIM * 2 * world,
               "d2h_bytes_per_step": NQ * K * 12, "ms_per_step": dt / steps * 1e3}
        sidx.check_exchange()
        # The merged multi-GPU answer of the last TIMED batch against one full-size single-GPU index searched
        # through the certified host entry point (the 1.5 GB database fits rank 0's GPU next to its shard).
        if rank == 0:
            from longbow_b200 import gpu as _gpu
            full = _gpu.DenseIndex(DIM, np.float16, _lib.METRIC_COSINE, local)
            full.reserve(n_rows)
            full.add_device(db.to(dev))
            ref_d, ref_l = full.search(qs[warm + steps - 1].numpy(), K)
            checks["multi_gpu_equals_single"] = bool(np.array_equal(ref_l, check_l) and np.array_equal(ref_d, check_d))
            checks["single_index_uncertified"] = int(full.last_uncertified())
            full.close()

    # ---- the HBM-bound end of the same path: one query per call (gpu.Index.Search as the reference calls it),
    # device-resident, streaming scan; reported as scan GB/s against the measured copy bandwidth
    hbm_scan = None
    if world == 1:
        q1 = d_qs[0, :1].contiguous()
        o1d = torch.empty((1, K), dtype=torch.float32, device=dev)
        o1l = torch.empty((1, K), dtype=torch.int64, device=dev)
        for _ in range(40):   # the e2e arm before this leaves the GPU half idle: let the clocks settle under this load
            sidx.search_device(q1, K, o1d, o1l)
        torch.cuda.synchronize()
        _lib.prof_read(reset=True)
        _lib.prof_enable(True)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        n1 = 200
        for i in range(n1):
            sidx.search_device(d_qs[(i % (warm + steps)), :1], K, o1d, o1l)
        s1.record()
        torch.cuda.synchronize()
        _lib.prof_enable(False)
        k_ms, k_n, k_units = _lib.prof_read(reset=True)
        l1 = _lib.launch_count()
        if k_n:
            rows_scanned = k_units / k_n
            hbm_scan = {"queries_per_call": 1, "ms_per_call": s0.elapsed_time(s1) / n1,
                        "scan_kernel_ms": k_ms / k_n, "bytes_per_launch": rows_scanned * DIM * 2,
                        "achieved_gbs": rows_scanned * DIM * 2 / (k_ms / k_n * 1e-3) / 1e9}

    # ---- the other BASELINE configs (C1, C3, C5 at N = 1; C4 at every N, row-sharded at N > 1): bench_extra.py
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    extra = []
    if args.configs != "none" and n_rows == N_ROWS:
        import bench_extra
        from longbow_b200 import gpu as _gpu2, pq as _pq2
        sidx.close()  # free the C2 shard before the large configs
        del d_qs
        torch.cuda.empty_cache()
        ctx = {"torch": torch, "_lib": _lib, "gpu": _gpu2, "pq": _pq2, "dev": dev, "local": local, "rank": rank,
               "world": world, "steps": max(6, steps // 2), "peaks": peaks, "cpu": (not args.no_cpu) and world == 1,
               "timer": bench_extra.Timer(torch, _lib, sampler, adaptive=(world == 1)),
               "only": None if args.configs == "all" else [c.strip().upper() for c in args.configs.split(",")]}
        extra = bench_extra.run_all(ctx)
        for r in extra:
            for name, ok in (r.get("checks") or {}).items():
                checks[f"{r['name']}.{name}"] = ok
    sampler.stop_flag = True
    sampler.join(timeout=5)
    log("clock sampler joined")

    if rank == 0:
        peak_tf = peaks.get("bf16_tflops", 1590.0)
        peak_src = "measured burst (MEASURED_PEAKS.json bf16_tflops)" if peaks else "fallback 1.59 PFLOP/s"
        rows_local = hi - lo
        # algorithmic flops of the bracketed launches: 2 * D per (query, row) pair they scanned
        flops_per_launch = 2.0 * DIM * scan1_units / max(scan1_n, 1)
        avg_scan_ms = scan1_ms / max(scan1_n, 1)            # single-stream pass: the kernel's own duration
        avg_scan_ms_timed = scan_ms / max(scan_n, 1)        # two-stream timed region: includes queueing / overlap
        achieved = flops_per_launch / (avg_scan_ms * 1e-3) / 1e12 if scan1_n else None
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r2_roofline_traffic.json")))
            for kname, rec in tj.items():
                if kname.startswith("dense_scan_tc") and world == 1 and n_rows == N_ROWS:
                    traffic = rec["dram_bytes_read"] + rec["dram_bytes_write"]
        except Exception:
            pass
        sustained = peaks.get("bf16_tflops_sustained")
        if hbm_scan is not None:
            hbm_scan["peak_gbs"] = hbm_peak = peaks.get("hbm_gbs", 6650.0)
            hbm_scan["frac"] = hbm_scan["achieved_gbs"] / hbm_peak
        roof = {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": (achieved / peak_tf) if achieved else None, "traffic": traffic, "peak_source": peak_src,
                "traffic_source": ("ncu --set full capture of this kernel on this workload (dram__bytes_read.sum + "
                                   "dram__bytes_write.sum per launch), profiles/r2_roofline_traffic.json; a recorded "
                                   "constant, not measured in this run") if traffic else None,
                "frac_of_sustained_peak": (achieved / sustained) if (achieved and sustained) else None,
                "algorithmic_bytes_per_launch": scan_units / max(scan_n, 1) / NQ * DIM * 2,
                "algorithmic_flops_per_launch": flops_per_launch,
                "hbm_bound_single_query_scan": hbm_scan,
                "kernel": "coarse distance scan + fused top-k (dense_scan)", "avg_launch_ms": avg_scan_ms,
                "avg_launch_ms_source": ("CUDA events around every main-scan launch of a single-stream pass over the timed "
                                         "region's batches, run right after it (same process, same clocks)"),
                "avg_launch_ms_timed_region": avg_scan_ms_timed,
                "frac_timed_region": (flops_per_launch / (avg_scan_ms_timed * 1e-3) / 1e12 / peak_tf) if scan_n else None,
                "timed_region_note": ("two streams overlap there: a launch's bracket also covers queueing behind and running "
                                      "beside the other stream's kernels, and consecutive scans overlap at their edges (the "
                                      "next scan's CTAs start on SMs the previous one has left), so share_of_step can exceed 1"),
                "launches_timed": scan1_n, "share_of_step": (avg_scan_ms * scan_n) / ms if ms else None,
                "rows_per_launch": scan_units / max(scan_n, 1) / NQ,
                "scan_gbs": scan_units / max(scan_n, 1) / NQ * DIM * 2 / (avg_scan_ms * 1e-3) / 1e9 if scan_n else None,
                "hbm_peak_gbs": hbm_peak}
        out = {
            "metric": METRIC_NAME, "value": NQ * steps / (ms * 1e-3), "unit": "queries/s", "n_gpus": world,
            "steps": steps, "warmup": warm, "ms_per_step": ms / steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": "C2: brute-force cosine k=100, 1M x 768 fp16 unit-norm embeddings, query batch 1024",
                       "rows": n_rows, "dim": DIM, "queries_per_step": NQ, "k": K,
                       "sharding": (f"rows/{world} per GPU; per-batch all-gather of the top-k records over NVLink peer memory "
                                    f"({sidx.exchange}) + merge on every rank") if world > 1 else "single GPU",
                       "l2_policy": "inputs larger than L2 (1.5 GB DB streamed per step; fresh query batch each step)",
                       "streams": ("batches alternate between 2 CUDA streams (tail kernels overlap the next scan)"
                                   + ("" if world == 1 else "; exchange (peer-memory push + signal + merge) in batch "
                                      "order on one side stream"))},
            "e2e": e2e, "gpu_launches": launches, "clocks": main_clocks, "roofline": roof,
            "checks": checks, "uncertified": uncertified, "configs": extra,
        }
        if world > 1:
            out["per_rank"] = per_rank
        failed = [name for name, ok in checks.items() if ok is False]
        for name in failed:
            print(f"[bench] CHECK FAILED: {name}", file=sys.stderr, flush=True)
        if not args.no_cpu and world == 1:
            out["cpu_baseline"] = cpu_baseline(db.numpy(), qs[0].numpy(), K)
            # the CPU sample doubles as a checker: the GPU's answers for the sample's queries must equal the port's
            # (O-exact, the reference's lane order: ids and distances bit-equal)
            from oracle import oracle as _o
            nqs = 16
            wd_, wl_ = _o.search(_o.COSINE, db.numpy(), qs[0].numpy()[:nqs], K)
            out["checks"]["C2.equals_exact_oracle_16q"] = bool(np.array_equal(c2_first_l[:nqs], wl_) and
                                                               np.array_equal(c2_first_d[:nqs], wd_))
            failed = [name for name, ok in out["checks"].items() if ok is False]
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0 and failed:
        raise SystemExit(f"bench: result checks failed: {failed}")  # a wrong answer is not a benchmark result


if __name__ == "__main__":
    main()
