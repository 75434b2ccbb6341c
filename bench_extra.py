"""bench_extra.py -- the other BASELINE.json configs (C1, C3, C4, C5) as driver-visible entries of bench.py's JSON
line (`configs`).  bench.py's headline stays C2; each entry here carries its own value (QPS, device-resident,
CUDA events), e2e (host buffers through the C ABI), roofline (algorithmic bytes / flops per SURVEY.md 8(d) over
the CUDA-event time of the dominant kernel), clocks sampled during its timed region, a bounded CPU baseline
(oracle port, all host cores) and a parity check of the GPU answers for the CPU sample's queries.

Synthetic inputs follow SURVEY.md 8(d) (distributions, sizes); they are generated on the GPU with seeded torch
generators and copied to the host only where the CPU sample needs them.
"""
import os
import time

import numpy as np


def _peaks(root):
    import json
    try:
        return json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


class Timer:
    """W warm-up + K timed calls of fn with CUDA events on the current stream; dominant-kernel time from lb_prof."""

    def __init__(self, torch, _lib, sampler, adaptive=True):
        self.torch, self._lib, self.sampler, self.adaptive = torch, _lib, sampler, adaptive

    def run(self, fn, steps, warm=3, min_ms=300.0, fin=None):
        """Timed region of at least `min_ms` (so the 10 ms NVML clock sampler sees it) when adaptive: the step count
        is raised after a calibration pass.  Multi-rank runs keep the given count (every rank must issue the same
        number of exchanges)."""
        torch, _lib = self.torch, self._lib
        for _ in range(warm):
            fn()
        if fin is not None:
            fin()
        torch.cuda.synchronize()
        if self.adaptive:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            one = max(e0.elapsed_time(e1), 1e-3)
            steps = int(min(5000, max(steps, min_ms / one)))
        _lib.prof_read(reset=True)
        _lib.prof_read_aux(reset=True)
        _lib.prof_enable(True)
        l0 = _lib.launch_count()
        t0 = time.time()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if fin is not None:
            fin()   # e.g. make the timing stream wait for work issued on side streams
        e1.record()
        torch.cuda.synchronize()
        t1 = time.time()
        _lib.prof_enable(False)
        k_ms, k_n, k_units = _lib.prof_read(reset=True)
        a_ms, a_n, a_units = _lib.prof_read_aux(reset=True)
        return {"ms": e0.elapsed_time(e1) / steps, "steps": steps, "kernel_ms": (k_ms / k_n) if k_n else None,
                "kernel_launches": k_n, "kernel_ms_per_step": k_ms / steps, "aux_ms_per_step": a_ms / steps,
                "aux_bytes_per_step": a_units / steps, "launches": _lib.launch_count() - l0,
                "clocks": self.sampler.window(t0, t1) if self.sampler else None}


def _host_timer(fn, steps, warm=2):
    for _ in range(warm):
        fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    return (time.perf_counter() - t0) / steps * 1e3


def _roof_hbm(bytes_per_launch, kernel_ms, peaks, kernel, traffic=None):
    peak = peaks.get("hbm_gbs", 6650.0)
    ach = bytes_per_launch / (kernel_ms * 1e-3) / 1e9 if kernel_ms else None
    return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": (ach / peak) if ach else None,
            "traffic": traffic, "kernel": kernel, "avg_launch_ms": kernel_ms,
            "algorithmic_bytes_per_launch": bytes_per_launch,
            "peak_source": "measured copy bandwidth (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6.65 TB/s"}


def _roof_tensor(flops_per_launch, kernel_ms, peaks, kernel, note=None):
    peak = peaks.get("bf16_tflops", 1590.0)
    ach = flops_per_launch / (kernel_ms * 1e-3) / 1e12 if kernel_ms else None
    out = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": (ach / peak) if ach else None,
           "traffic": None, "kernel": kernel, "avg_launch_ms": kernel_ms,
           "algorithmic_flops_per_launch": flops_per_launch,
           "peak_source": "measured burst bf16 (MEASURED_PEAKS.json bf16_tflops)" if peaks else "fallback 1.59 PFLOP/s"}
    if note:
        out["note"] = note
    return out


# ----------------------------------------------------------------------------------------------- C1
def config_c1(ctx):
    torch, _lib, gpu, dev = ctx["torch"], ctx["_lib"], ctx["gpu"], ctx["dev"]
    N, D, Q, K = 100_000, 128, 1000, 10
    g = torch.Generator(device=dev).manual_seed(1001)
    db = torch.rand((N, D), generator=g, device=dev)
    g2 = torch.Generator(device=dev).manual_seed(1002)
    qs = torch.rand((Q, D), generator=g2, device=dev)
    idx = gpu.DenseIndex(D, np.float32, _lib.METRIC_L2, ctx["local"])
    idx.add_device(db)
    od = torch.empty((Q, K), dtype=torch.float32, device=dev)
    ol = torch.empty((Q, K), dtype=torch.int64, device=dev)
    t = ctx["timer"].run(lambda: idx.search_device(qs, K, od, ol), ctx["steps"])
    hq = qs.cpu().numpy()
    hd, hl = np.empty((Q, K), np.float32), np.empty((Q, K), np.int64)
    e2e_ms = _host_timer(lambda: idx.search_into(hq, K, hd, hl), max(3, ctx["steps"] // 2))
    out = {"name": "C1", "workload": "brute-force L2 k=10, 100k x 128 fp32 U[0,1), 1000 queries (the reference's CPU-runnable case)",
           "value": Q / (t["ms"] * 1e-3), "unit": "queries/s", "ms_per_step": t["ms"], "steps": t["steps"], "dtype": "f32",
           "e2e": {"value": Q / (e2e_ms * 1e-3), "unit": "queries/s", "ms_per_step": e2e_ms,
                   "h2d_bytes_per_step": Q * D * 4, "d2h_bytes_per_step": Q * K * 12},
           "roofline": _roof_tensor(2.0 * Q * N * D, t["kernel_ms"], ctx["peaks"], "dense_scan_tc<tf32 x3> (3xTF32 split: "
                                    "three MMAs per k-step; flops counted once)",
                                    "51 MB database is L2-resident.  fp32 rows cost three kind::tf32 MMAs per k-step and "
                                    "tf32 runs at half the bf16 rate, so this method's ceiling is peak / 6 "
                                    "(method_ceiling_tflops); the search is 5 kernels of 20-130 us each"),
           "gpu_launches": t["launches"], "clocks": t["clocks"], "checks": {}}
    if out["roofline"]["achieved"]:
        out["roofline"]["method_ceiling_tflops"] = out["roofline"]["peak"] / 6.0
        out["roofline"]["frac_of_method_ceiling"] = out["roofline"]["achieved"] / (out["roofline"]["peak"] / 6.0)
    if ctx["cpu"]:
        from oracle import oracle
        cores = oracle.fast_use_all_cores()
        dbh = db.cpu().numpy()
        t0 = time.perf_counter()
        wd, wl = oracle.search(oracle.L2, dbh, hq, K, impl="fast")
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": Q / dt, "unit": "queries/s", "cores": cores, "kind": "port",
                               "sample": f"the full config: {Q} queries x {N} x {D} fp32, {oracle.fast_isa()} + OpenMP, {dt:.2f} s"}
        ed, el = oracle.search(oracle.L2, dbh, hq[:64], K)  # O-exact on 64 queries
        out["checks"]["equals_exact_oracle_64q"] = bool(np.array_equal(hl[:64], el) and np.array_equal(hd[:64], ed))
    idx.close()
    return out


# ----------------------------------------------------------------------------------------------- C3
def config_c3(ctx):
    torch, _lib, gpu, pq, dev = ctx["torch"], ctx["_lib"], ctx["gpu"], ctx["pq"], ctx["dev"]
    N, D, M, K, KP, Q = ctx.get("c3_rows", 10_000_000), 768, 96, 10, 100, 256
    g = torch.Generator(device=dev).manual_seed(3001)
    codes = torch.randint(0, 256, (N, M), generator=g, device=dev, dtype=torch.uint8)
    g = torch.Generator(device=dev).manual_seed(3002)
    cb = torch.randn((M, 256, D // M), generator=g, device=dev)
    g = torch.Generator(device=dev).manual_seed(3004)
    qs = torch.randn((Q, D), generator=g, device=dev)
    enc = pq.PQEncoder(D, M, 256, cb.cpu().numpy(), ctx["local"])
    enc.add_codes_device(codes)
    # raw vectors for the fp32 re-rank: decode(codes) + N(0, 0.05)  (30.7 GB at 10 M rows)
    raw = gpu.DenseIndex(D, np.float32, _lib.METRIC_L2, ctx["local"])
    raw.reserve(N)
    g = torch.Generator(device=dev).manual_seed(3003)
    step = 250_000
    for lo in range(0, N, step):
        hi = min(N, lo + step)
        c = codes[lo:hi].long()
        v = torch.stack([cb[m][c[:, m]] for m in range(M)], dim=1).reshape(hi - lo, D)
        v += 0.05 * torch.randn(v.shape, generator=g, device=dev)
        raw.add_device(v.contiguous())
    del v, c
    enc.attach_raw(raw)
    od = torch.empty((Q, K), dtype=torch.float32, device=dev)
    ol = torch.empty((Q, K), dtype=torch.int64, device=dev)
    steps = max(2, ctx["steps"] // 4)
    t = ctx["timer"].run(lambda: enc.search_device(qs, K, KP, od, ol), steps, warm=1)
    # one query per pass: the HBM-bound form of the scan (N * M code bytes per pass)
    q1 = qs[:1].contiguous()
    o1d = torch.empty((1, K), dtype=torch.float32, device=dev)
    o1l = torch.empty((1, K), dtype=torch.int64, device=dev)
    t1 = ctx["timer"].run(lambda: enc.search_device(q1, K, KP, o1d, o1l), ctx["steps"])
    hq = qs.cpu().numpy()
    hd, hl = np.empty((Q, K), np.float32), np.empty((Q, K), np.int64)
    e2e_ms = _host_timer(lambda: enc.search_into(hq, K, KP, hd, hl), 2, warm=1)
    lookups = float(Q) * N * M
    out = {"name": "C3", "workload": f"PQ ADC scan M=96 nbits=8 over {N} x 768 codes + fp32 re-rank k'=100 -> k=10, 256 queries",
           "value": Q / (t["ms"] * 1e-3), "unit": "queries/s", "ms_per_step": t["ms"], "steps": t["steps"], "dtype": "u8 codes; coarse f16 x f16 -> f32 (batch) or u16/u32 table sums (one query); exact stage f32",
           "e2e": {"value": Q / (e2e_ms * 1e-3), "unit": "queries/s", "ms_per_step": e2e_ms,
                   "h2d_bytes_per_step": Q * D * 4, "d2h_bytes_per_step": Q * K * 12, "uncertified": enc.last_uncertified()},
           # dominant kernel of the batch: the dense tensor-core scan over the decoded fp16 slabs (2 Q N D flop per step);
           # the decode that feeds it is HBM-bound and reported beside it, as is the one-query look-up pass
           "roofline": _roof_tensor(2.0 * Q * N * D, t["kernel_ms_per_step"], ctx["peaks"],
                                    "dense_scan_tc<f16,L2> over decoded slabs (3 launches per step; time = their sum)",
                                    "batched PQ coarse stage: decode to fp16 (pq_decode_kernel) + tensor-core scan + exact fp32 "
                                    "table sums of the candidates + certification (csrc/pq_gemm.cu)"),
           "decode": _roof_hbm(t["aux_bytes_per_step"], t["aux_ms_per_step"], ctx["peaks"],
                               "pq_decode_kernel<8> (per step: N*M code bytes in, N*D*2 bytes out)"),
           "single_query_roofline": _roof_hbm(float(N) * M, t1["kernel_ms"], ctx["peaks"],
                                              "adc_coarse_kernel<1,3> (one query per pass: N*M code bytes)"),
           "single_query": {"ms_per_call": t1["ms"], "scan_kernel_ms": t1["kernel_ms"]},
           "gpu_launches": t["launches"], "clocks": t["clocks"], "checks": {}}
    if ctx["cpu"]:
        from oracle import oracle
        cores = oracle.fast_use_all_cores()
        nqs = min(Q, max(8, cores))
        ch, cbh = codes.cpu().numpy(), cb.cpu().numpy()
        t0 = time.perf_counter()
        wd, wl = oracle.pq_search(cbh, ch, None, hq[:nqs], K, 0, impl="fast")
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": nqs / dt, "unit": "queries/s", "cores": cores, "kind": "port",
                               "sample": f"{nqs} queries x full {N} x 96 codes, ADC scan + top-10 without the fp32 re-rank "
                                         f"(raw vectors stay on the GPU), {oracle.fast_isa()} gather + OpenMP, {dt:.2f} s"}
        # checker: O-exact (sequential fp32 sum in j, simd.go:345-355) on 8 queries, bit-equal ids and distances
        enc.attach_raw(None)
        gd, gl = enc.search(hq[:8], K)          # 8 queries: the look-up path
        ed, el = oracle.pq_search(cbh, ch, None, hq[:8], K, 0)
        out["checks"]["adc_topk_equals_exact_oracle_8q"] = bool(np.array_equal(gl, el) and np.array_equal(gd, ed))
        gdb, glb = enc.search(hq, K)            # the whole batch: decode + tensor-core coarse stage
        out["checks"]["adc_topk_batch_path_equals_exact_oracle_8q"] = bool(np.array_equal(glb[:8], el) and
                                                                           np.array_equal(gdb[:8], ed))
        out["batch_path_uncertified"] = enc.last_uncertified()
    enc.close()
    raw.close()
    return out


# ----------------------------------------------------------------------------------------------- C4
def _c4_shard(torch, dev, rank, rows):
    g = torch.Generator(device=dev).manual_seed(4001 + rank)
    return torch.randint(-128, 128, (rows, 128), generator=g, device=dev, dtype=torch.int8)


def config_c4(ctx):
    torch, _lib, gpu, dev = ctx["torch"], ctx["_lib"], ctx["gpu"], ctx["dev"]
    world, rank = ctx["world"], ctx["rank"]
    ROWS, D, Q, K = ctx.get("c4_rows", 12_500_000), 128, 1024, 10
    from longbow_b200.shard import ShardedIndex
    sidx = ShardedIndex(D, np.int8, _lib.METRIC_DOT, ROWS * world, rank, world, ctx["local"])
    sidx.index.reserve(ROWS)
    sidx.add_local_device(_c4_shard(torch, dev, rank, ROWS))
    g = torch.Generator(device=dev).manual_seed(4999)   # (not 4001 + rank: the queries must not be rows of a shard)
    nb = 4
    qs = torch.randint(-128, 128, (nb, Q, D), generator=g, device=dev, dtype=torch.int8)
    od = torch.empty((Q, K), dtype=torch.float32, device=dev)
    ol = torch.empty((Q, K), dtype=torch.int64, device=dev)
    steps = max(4, ctx["steps"] // 2) if world == 1 else 80  # 80 x ~4 ms: long enough for the clock sampler
    it = [0]

    # N > 1: the exchange (peer-memory push + signal + wait + merge) of batch s runs on the side stream under the
    # scan of batch s + 1, as in the C2 arm; every batch is still one complete sharded search
    def step():
        sidx.search_device(qs[it[0] % nb], K, od, ol, overlap=world > 1)
        it[0] += 1
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    t = ctx["timer"].run(step, steps, fin=sidx.wait if world > 1 else None)
    ms = t["ms"]
    per_rank = None
    if world > 1:
        tt = torch.tensor([ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
        # every rank's own step and scan-kernel time (a slow GPU of the box paces all of them through the exchange)
        mine = torch.tensor([t["ms"], t["kernel_ms"] or 0.0, float((t["clocks"] or {}).get("sm_mhz") or 0)], device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"step_ms": [round(float(x[0]), 4) for x in allr], "scan_kernel_ms": [round(float(x[1]), 4) for x in allr],
                    "sm_mhz": [int(x[2]) for x in allr]}
    last_q = qs[(it[0] - 1) % nb]
    got_d, got_l = od.cpu().numpy(), ol.cpu().numpy()
    out = {"name": "C4", "workload": f"int8 dot-product k=10 over {ROWS * world} x 128 int8 "
                                     f"({'one GPU shard of the 100M-row config' if world == 1 else f'row-sharded over {world} GPUs, peer-memory all-gather + merge'}), 1024 queries",
           "value": Q / (ms * 1e-3), "unit": "queries/s", "ms_per_step": ms, "steps": t["steps"], "dtype": "i8 (s32 accumulate)",
           "n_gpus": world, "scaling": "weak (12.5 M rows per GPU)",
           "roofline": _roof_tensor(2.0 * Q * ROWS * D, t["kernel_ms"], ctx["peaks"], "dense_scan_tc<i8,dot> (integer MMA, kind::i8)",
                                    "fraction quoted against the measured bf16 peak; the int8 dense peak is 2x that. 128-byte "
                                    "rows: one k-block per tile, bound by the TMEM drain of the epilogue (DESIGN.md)"),
           "gpu_launches": t["launches"], "clocks": t["clocks"], "checks": {}, "uncertified": sidx.uncertified()}
    if world > 1:
        sidx.check_exchange()
        out["per_rank"] = per_rank
    if world == 1:
        hq = last_q.cpu().numpy()
        hd, hl = np.empty((Q, K), np.float32), np.empty((Q, K), np.int64)
        e2e_ms = _host_timer(lambda: sidx.index.search_into(hq, K, hd, hl), max(3, steps // 2))
        out["e2e"] = {"value": Q / (e2e_ms * 1e-3), "unit": "queries/s", "ms_per_step": e2e_ms,
                      "h2d_bytes_per_step": Q * D, "d2h_bytes_per_step": Q * K * 12}
        out["checks"]["host_api_equals_device_api"] = bool(np.array_equal(hl, got_l) and np.array_equal(hd, got_d))
        # one query per call: HBM-bound pass over the shard
        q1 = qs[0, :1].contiguous()
        o1d = torch.empty((1, K), dtype=torch.float32, device=dev)
        o1l = torch.empty((1, K), dtype=torch.int64, device=dev)
        t1 = ctx["timer"].run(lambda: sidx.index.search_device(q1, K, o1d, o1l), ctx["steps"])
        out["single_query"] = {"ms_per_call": t1["ms"],
                               "roofline": _roof_hbm(float(ROWS) * D, t1["kernel_ms"], ctx["peaks"], "dense_scan_stream<int8,dot>")}
        if ctx["cpu"]:
            from oracle import oracle
            cores = oracle.fast_use_all_cores()
            nqs = max(8, cores)
            dbh = _c4_shard(torch, dev, rank, ROWS).cpu().numpy()
            t0 = time.perf_counter()
            wd, wl = oracle.search(oracle.DOT, dbh, hq[:nqs], K, impl="fast")
            dt = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": nqs / dt, "unit": "queries/s", "cores": cores, "kind": "port",
                                   "sample": f"{nqs} queries x the full {ROWS} x 128 int8 shard, {oracle.fast_isa()} + OpenMP, {dt:.2f} s"}
            out["checks"]["equals_oracle"] = bool(np.array_equal(hl[:nqs], wl) and np.array_equal(hd[:nqs], wd))
    elif rank == 0:
        # the merged answer of the last timed batch against ONE index holding every rank's rows (12.8 GB at 8 GPUs)
        full = gpu.DenseIndex(D, np.int8, _lib.METRIC_DOT, ctx["local"])
        full.reserve(ROWS * world)
        for r in range(world):
            full.add_device(_c4_shard(torch, dev, r, ROWS))
        rd, rl = full.search(last_q.cpu().numpy(), K)
        out["checks"]["multi_gpu_equals_single"] = bool(np.array_equal(rl, got_l) and np.array_equal(rd, got_d))
        full.close()
    sidx.close()
    return out


# ----------------------------------------------------------------------------------------------- C5
def config_c5(ctx):
    torch, _lib, gpu, dev = ctx["torch"], ctx["_lib"], ctx["gpu"], ctx["dev"]
    N, D, Q, C, K = ctx.get("c5_rows", 10_000_000), 384, 4096, 128, 10
    idx = gpu.DenseIndex(D, np.float32, _lib.METRIC_L2, ctx["local"])
    idx.reserve(N)
    g = torch.Generator(device=dev).manual_seed(5001)
    db = torch.empty((N, D), dtype=torch.float32, device=dev)  # kept for the CPU sample's row gather (15.4 GB)
    step = 1_000_000
    for lo in range(0, N, step):
        db[lo:lo + step] = torch.randn((min(step, N - lo), D), generator=g, device=dev)
    idx.add_device(db)
    g = torch.Generator(device=dev).manual_seed(5002)
    qs = torch.randn((Q, D), generator=g, device=dev)
    # candidate lists shaped like an ef=128 frontier: the exact 64 nearest rows (what a converged
    # walk holds) + 64 random ids (what it visited on the way), shuffled
    top_d = torch.empty((Q, 64), dtype=torch.float32, device=dev)
    top_l = torch.empty((Q, 64), dtype=torch.int64, device=dev)
    idx.search_device(qs, 64, top_d, top_l)
    g = torch.Generator(device=dev).manual_seed(5003)
    rnd = torch.randint(0, N, (Q, C - 64), generator=g, device=dev, dtype=torch.int64)
    cand = torch.cat([top_l, rnd], dim=1)
    perm = torch.argsort(torch.rand((Q, C), generator=g, device=dev), dim=1)
    cand = torch.gather(cand, 1, perm).to(torch.uint32).contiguous()
    g = torch.Generator(device=dev).manual_seed(5004)
    tomb = (torch.rand(N, generator=g, device=dev) < 0.05)
    g = torch.Generator(device=dev).manual_seed(5005)
    allow = (torch.rand(N, generator=g, device=dev) < 0.30)
    tomb_h, allow_h = tomb.cpu().numpy(), allow.cpu().numpy()
    idx.set_tombstones(tomb_h)
    allow_packed = gpu.pack_bitmap(allow_h)
    allow_d = torch.from_numpy(allow_packed.view(np.int64)).to(dev)
    od = torch.empty((Q, K), dtype=torch.float32, device=dev)
    ol = torch.empty((Q, K), dtype=torch.int64, device=dev)
    t = ctx["timer"].run(lambda: idx.rerank_device(qs, cand, K, od, ol, allow=allow_d), ctx["steps"])
    live = (~tomb & allow)[cand.long()].float().mean().item()
    hq, hc = qs.cpu().numpy(), cand.cpu().numpy()
    hd, hl = np.empty((Q, K), np.float32), np.empty((Q, K), np.int64)
    lib = _lib.load()

    def host_call():
        _lib.check(lib.lb_index_rerank(idx._h, hq.ctypes.data, Q, hc.ctypes.data, C, K, allow_packed.ctypes.data,
                                       hd.ctypes.data, hl.ctypes.data))
    e2e_ms = _host_timer(host_call, max(3, ctx["steps"] // 3))
    bytes_gathered = Q * C * live * D * 4 + Q * C * 4
    out = {"name": "C5", "workload": f"HNSW-shaped re-rank: batch 4096 x ef=128 candidate ids (exact top-64 + 64 random), {N} x 384 fp32, "
                                     "tombstones 5% + predicate allow-bitmap 30% applied in-kernel, k=10",
           "value": Q / (t["ms"] * 1e-3), "unit": "queries/s", "ms_per_step": t["ms"], "steps": t["steps"], "dtype": "f32",
           "e2e": {"value": Q / (e2e_ms * 1e-3), "unit": "queries/s", "ms_per_step": e2e_ms,
                   "h2d_bytes_per_step": Q * D * 4 + Q * C * 4 + allow_packed.nbytes, "d2h_bytes_per_step": Q * K * 12},
           "roofline": _roof_hbm(bytes_gathered, t["ms"], ctx["peaks"],
                                 "rescore_coop_kernel<float,L2> (gather of the live candidates' rows + ids)"),
           "live_candidate_fraction": live, "gpu_launches": t["launches"], "clocks": t["clocks"],
           "checks": {"host_api_equals_device_api": bool(np.array_equal(hl, ol.cpu().numpy()) and np.array_equal(hd, od.cpu().numpy()))}}
    # without bitmaps: every candidate row is gathered (805 MB per batch)
    idx.set_tombstones(None)
    t2 = ctx["timer"].run(lambda: idx.rerank_device(qs, cand, K, od, ol), ctx["steps"])
    out["all_candidates_live"] = {"ms_per_step": t2["ms"], "value": Q / (t2["ms"] * 1e-3),
                                  "roofline": _roof_hbm(Q * C * D * 4.0 + Q * C * 4, t2["ms"], ctx["peaks"], "rescore_coop_kernel<float,L2>")}
    if ctx["cpu"]:
        from oracle import oracle
        cores = oracle.fast_use_all_cores()
        nqs = 256
        # the CPU sample re-ranks against a compact copy of just the rows its candidates name (same arithmetic)
        ids = cand[:nqs].long().reshape(-1)
        uniq, inv = torch.unique(ids, return_inverse=True)
        subh = db[uniq].cpu().numpy()
        candh = inv.reshape(nqs, C).cpu().numpy().astype(np.int64)
        tomb_s = gpu.pack_bitmap(tomb_h[uniq.cpu().numpy()])
        allow_s = gpu.pack_bitmap(allow_h[uniq.cpu().numpy()])
        t0 = time.perf_counter()
        wd, wl = oracle.rerank(oracle.L2, subh, hq[:nqs], candh, K, tomb=tomb_s, allow=allow_s, impl="fast")
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": nqs / dt, "unit": "queries/s", "cores": cores, "kind": "port",
                               "sample": f"{nqs} queries x 128 candidates re-ranked against a compact copy of the rows they name "
                                         f"(bitmaps remapped), {oracle.fast_isa()} + OpenMP, {dt:.3f} s"}
        # checker: O-exact (reference lane order) on the same sample, bit-equal ids and distances
        ed, el = oracle.rerank(oracle.L2, subh, hq[:nqs], candh, K, tomb=tomb_s, allow=allow_s)
        uq = uniq.cpu().numpy()
        el_global = np.where(el >= 0, uq[np.clip(el, 0, None)], -1)
        out["checks"]["equals_exact_oracle_256q"] = bool(np.array_equal(hl[:nqs], el_global) and np.array_equal(hd[:nqs], ed))
    idx.close()
    del db
    return out


# ----------------------------------------------------------------------------------------------- C5w
def config_c5_walk(ctx):
    """Config 5 with the REAL walk: a navigable graph (24 exact nearest neighbours + 8 random long-range edges per
    node, built on this GPU with the library's own brute-force search) over 1 M x 384 fp32 rows; batch 4096, ef=128;
    the GPU runs ArrowHNSW.searchLayer for every query (csrc/hnsw.cu) and re-ranks the frontier with tombstones +
    predicate bitmap in-kernel.  The 10 M-row graph of the config cannot be BUILT inside a benchmark run; the walk's
    cost per query depends on ef and the degree, not on N, so the 1 M graph is the bounded stand-in (stated)."""
    torch, _lib, gpu, dev = ctx["torch"], ctx["_lib"], ctx["gpu"], ctx["dev"]
    from longbow_b200 import store
    N, D, Q, EF, K, DEG, KNN = ctx.get("c5w_rows", 1_000_000), 384, 4096, 128, 10, 32, 24
    g = torch.Generator(device=dev).manual_seed(5101)
    db = torch.randn((N, D), generator=g, device=dev)
    idx = gpu.DenseIndex(D, np.float32, _lib.METRIC_L2, ctx["local"])
    idx.reserve(N)
    idx.add_device(db)
    nbrs = torch.empty((N, DEG), dtype=torch.int64, device=dev)
    od = torch.empty((4096, KNN + 1), dtype=torch.float32, device=dev)
    ol = torch.empty((4096, KNN + 1), dtype=torch.int64, device=dev)
    t0 = time.time()
    for lo in range(0, N, 4096):
        hi = min(N, lo + 4096)
        idx.search_device(db[lo:hi], KNN + 1, od[:hi - lo], ol[:hi - lo])
        nbrs[lo:hi, :KNN] = ol[:hi - lo, 1:]  # rank 0 is the row itself
    nbrs[:, KNN:] = torch.randint(0, N, (N, DEG - KNN), generator=g, device=dev)
    torch.cuda.synchronize()
    build_s = time.time() - t0
    graph = store.HNSWGraph(idx, DEG)
    nb32 = nbrs.to(torch.uint32).contiguous()
    _lib.check(_lib.load().lb_graph_set_layer_device(graph._h, nb32.data_ptr(), None, N,
                                                     torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    g = torch.Generator(device=dev).manual_seed(5102)
    qs = torch.randn((Q, D), generator=g, device=dev)
    entries = torch.zeros(Q, dtype=torch.int32, device=dev)
    g = torch.Generator(device=dev).manual_seed(5104)
    tomb = (torch.rand(N, generator=g, device=dev) < 0.05).cpu().numpy()
    g = torch.Generator(device=dev).manual_seed(5105)
    allow = (torch.rand(N, generator=g, device=dev) < 0.30).cpu().numpy()
    idx.set_tombstones(tomb)
    allow_packed = gpu.pack_bitmap(allow)
    allow_d = torch.from_numpy(allow_packed.view(np.int64)).to(dev)
    out_d = torch.empty((Q, K), dtype=torch.float32, device=dev)
    out_l = torch.empty((Q, K), dtype=torch.int64, device=dev)
    fail = torch.zeros(1, dtype=torch.int32, device=dev)
    t = ctx["timer"].run(lambda: graph.search_device(qs, entries, EF, K, out_d, out_l, fail, allow=allow_d), max(3, ctx["steps"] // 2), warm=2)
    hq, he = qs.cpu().numpy(), np.zeros(Q, np.uint32)
    e2e_ms = _host_timer(lambda: graph.Search(hq, he, EF, K, allow=allow_packed), 3, warm=1)
    gi, gd, gv = graph.SearchLayer(hq[:512], he[:512], EF)
    visited = float(gv.mean())
    out = {"name": "C5w", "workload": f"HNSW search ef=128, batch 4096: GPU graph walk (searchLayer, degree 32) + GPU re-rank with "
                                      f"tombstones 5% + allow-bitmap 30%, k=10, {N} x 384 fp32 (graph of the 10 M config scaled to "
                                      f"what can be built inside the run: {build_s:.1f} s on this GPU)",
           "value": Q / (t["ms"] * 1e-3), "unit": "queries/s", "ms_per_step": t["ms"], "steps": t["steps"], "dtype": "f32",
           "e2e": {"value": Q / (e2e_ms * 1e-3), "unit": "queries/s", "ms_per_step": e2e_ms,
                   "h2d_bytes_per_step": Q * D * 4 + Q * 4 + allow_packed.nbytes, "d2h_bytes_per_step": Q * K * 12},
           "roofline": _roof_hbm(Q * visited * D * 4.0, t["ms"], ctx["peaks"],
                                 "hnsw_search_layer_kernel<float,L2> (row gathers of the visited nodes: "
                                 f"{visited:.0f} distance evaluations per query) + rescore_coop_kernel"),
           "distance_evaluations_per_query": visited, "walk_failures": int(fail.item()),
           "gpu_launches": t["launches"], "clocks": t["clocks"], "checks": {"no_walk_failures": int(fail.item()) == 0}}
    if ctx["cpu"]:
        from oracle import oracle
        cores = oracle.fast_use_all_cores()
        nqs = 256
        dbh, nbh = db.cpu().numpy(), nb32.cpu().numpy()
        t0 = time.perf_counter()
        wi, wd, wv = oracle.hnsw_search_layer(oracle.L2, dbh, nbh, None, hq[:nqs], he[:nqs], EF)
        cand = wi.astype(np.int64)
        cand[wi == 0xFFFFFFFF] = -1
        ed, el = oracle.rerank(oracle.L2, dbh, hq[:nqs], cand, K, tomb=gpu.pack_bitmap(tomb), allow=allow_packed)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": nqs / dt, "unit": "queries/s", "cores": cores, "kind": "port",
                               "sample": f"{nqs} queries: searchLayer restatement (heaps, per-neighbour distance) + re-rank on the same graph, "
                                         f"OpenMP over queries, {dt:.3f} s"}
        gdk, glk = graph.Search(hq[:nqs], he[:nqs], EF, K, allow=allow_packed)
        out["checks"]["frontier_equals_cpu_walk_256q"] = bool(np.array_equal(gi[:nqs], wi) and np.array_equal(gd[:nqs], wd))
        out["checks"]["topk_equals_exact_oracle_256q"] = bool(np.array_equal(glk, el) and np.array_equal(gdk, ed))
    graph.Close()
    idx.close()
    return out


def run_all(ctx):
    """Every extra config the rank count allows: C4 always (row-sharded at N > 1); C1 / C3 / C5 at N = 1."""
    out = []
    names = ctx.get("only") or (["C1", "C3", "C4", "C5", "C5W"] if ctx["world"] == 1 else ["C4"])
    fns = {"C1": config_c1, "C3": config_c3, "C4": config_c4, "C5": config_c5, "C5W": config_c5_walk}
    for n in names:
        t0 = time.time()
        try:
            r = fns[n](ctx)
        except Exception as e:  # one config failing must not take the headline down with it -- but it is reported
            import traceback
            traceback.print_exc()
            r = {"name": n, "error": f"{type(e).__name__}: {e}", "checks": {"ran": False}}
        r["wall_s"] = round(time.time() - t0, 1)
        ctx["torch"].cuda.empty_cache()
        out.append(r)
    return out
