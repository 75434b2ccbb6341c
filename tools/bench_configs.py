"""Device-resident timing of every BASELINE.json config on ONE GPU (bench.py reports only the headline,
C2).  Each line: config, ms per batch, QPS, algorithmic bytes / flops and the roofline fraction against
MEASURED_PEAKS.json.  Sizes are the configs' own unless noted (C4 = one GPU's 12.5M-row shard of the
100M-row database; C3's raw fp32 vectors for the re-rank are limited by what fits beside the codes).

  python tools/bench_configs.py [c1 c2 c3 c4 c5 ...] > profiles/...
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from longbow_b200 import _lib, gpu, pq

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    PEAKS = {}
HBM = PEAKS.get("hbm_gbs", 6650.0)
TF = PEAKS.get("bf16_tflops", 1590.0)
dev = torch.device("cuda", 0)
STEPS = int(os.environ.get("STEPS", "10"))
for _opt in ("tc_pair", "dense_scan", "tc_boot_tiles", "pq_scan", "pq_ahead", "pq_ring"):
    if os.environ.get(_opt.upper()):
        _lib.set_option(_opt, int(os.environ[_opt.upper()]))


def timed(fn, steps=STEPS, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _lib.prof_read(True)
    _lib.prof_enable(True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    _lib.prof_enable(False)
    scan_ms, scan_n, _u = _lib.prof_read(True)
    return e0.elapsed_time(e1) / steps, (scan_ms / scan_n if scan_n else None)


def report(name, ms, scan_ms, nq, bytes_pass, flops, note=""):
    out = {"config": name, "ms_per_batch": round(ms, 4), "qps": round(nq / (ms * 1e-3), 1),
           "scan_kernel_ms": None if scan_ms is None else round(scan_ms, 4)}
    t = (scan_ms or ms) * 1e-3
    if bytes_pass:
        out["scan_GBps"] = round(bytes_pass / t / 1e9, 1)
        out["hbm_frac_of_measured"] = round(bytes_pass / t / 1e9 / HBM, 4)
    if flops:
        out["scan_TFLOPs"] = round(flops / t / 1e12, 1)
        out["tensor_frac_of_measured_bf16"] = round(flops / t / 1e12 / TF, 4)
    if note:
        out["note"] = note
    print(json.dumps(out), flush=True)


def c1():
    g = torch.Generator(device=dev).manual_seed(1001)
    N, D, Q, K = 100_000, 128, 1000, 10
    db = torch.rand((N, D), generator=g, device=dev)
    qs = torch.rand((Q, D), generator=g, device=dev)
    idx = gpu.DenseIndex(D, np.float32, _lib.METRIC_L2)
    idx.add_device(db)
    od = torch.empty((Q, K), dtype=torch.float32, device=dev)
    ol = torch.empty((Q, K), dtype=torch.int64, device=dev)
    ms, sm = timed(lambda: idx.search_device(qs, K, od, ol))
    report("C1 brute-force L2 k=10, 100k x 128 fp32, 1000 queries", ms, sm, Q, N * D * 4, 2.0 * Q * N * D,
           "fp32 SIMT scan (L2-resident DB: compute-bound, see DESIGN.md)")
    idx.close()


def c2():
    g = torch.Generator(device=dev).manual_seed(2001)
    N, D, Q, K = 1_000_000, 768, 1024, 100
    db = torch.randn((N, D), generator=g, device=dev)
    db = (db / db.norm(dim=1, keepdim=True)).half()
    qs = torch.randn((Q, D), generator=g, device=dev)
    qs = (qs / qs.norm(dim=1, keepdim=True)).half()
    idx = gpu.DenseIndex(D, np.float16, _lib.METRIC_COSINE)
    idx.reserve(N)
    idx.add_device(db)
    od = torch.empty((Q, K), dtype=torch.float32, device=dev)
    ol = torch.empty((Q, K), dtype=torch.int64, device=dev)
    ms, sm = timed(lambda: idx.search_device(qs, K, od, ol))
    report("C2 brute-force cosine k=100, 1M x 768 fp16, 1024 queries", ms, sm, Q, N * D * 2, 2.0 * Q * N * D)
    for nq in (1, 2, 4, 8, 32):
        qn = qs[:nq].contiguous()
        odn = torch.empty((nq, K), dtype=torch.float32, device=dev)
        oln = torch.empty((nq, K), dtype=torch.int64, device=dev)
        ms, sm = timed(lambda: idx.search_device(qn, K, odn, oln))
        report(f"C2 database, {nq} quer{'y' if nq == 1 else 'ies'} per call (HBM-bound: streaming scan at 1 query, one padded tensor-core query block above)", ms, sm, nq,
               N * D * 2, None)
    idx.close()


def c3():
    g = torch.Generator(device=dev).manual_seed(3001)
    N, D, M, K, KP = int(os.environ.get("C3_N", 10_000_000)), 768, 96, 10, 100
    Q = int(os.environ.get("C3_Q", 256))
    cb = torch.randn((M, 256, D // M), generator=g, device=dev)
    codes = torch.randint(0, 256, (N, M), generator=g, device=dev, dtype=torch.uint8)
    qs = torch.randn((Q, D), generator=g, device=dev)
    enc = pq.PQEncoder(D, M, 256, cb.cpu().numpy())
    enc.add_codes_device(codes)
    od = torch.empty((Q, K), dtype=torch.float32, device=dev)
    ol = torch.empty((Q, K), dtype=torch.int64, device=dev)
    for nq1 in (1, 4):
        q1 = qs[:nq1].contiguous()
        o1d = torch.empty((nq1, K), dtype=torch.float32, device=dev)
        o1l = torch.empty((nq1, K), dtype=torch.int64, device=dev)
        ms, sm = timed(lambda: enc.search_device(q1, K, 0, o1d, o1l), steps=STEPS, warm=2)
        report(f"C3 PQ ADC scan, {nq1} quer{'y' if nq1 == 1 else 'ies'} per pass over {N} x 96 codes (HBM-bound: N*M bytes per pass)",
               ms, sm, nq1, N * M, None)
    ms, sm = timed(lambda: enc.search_device(qs, K, 0, od, ol), steps=max(2, STEPS // 3), warm=1)
    report(f"C3 PQ ADC scan M=96 over {N} x 96 codes, {Q} queries, k=10 (no re-rank)", ms, sm, Q, None, None,
           f"ADC lookups/s = {Q * N * M / (ms * 1e-3):.3e}; code bytes per pass {N * M / 1e6:.0f} MB; "
           f"per-query pass rate {Q * N * M / (ms * 1e-3) / 1e9:.1f} GB/s of codes consumed")
    # with the fp32 re-rank of k'=100 candidates (raw vectors: decode(codes) + noise)
    NR = int(os.environ.get("C3_RAW", N))
    if NR < N:  # the re-rank needs the raw vector of every coded row
        enc.close()
        return
    raw = gpu.DenseIndex(D, np.float32, _lib.METRIC_L2)
    raw.reserve(NR)
    step = 500_000
    for lo in range(0, NR, step):
        hi = min(NR, lo + step)
        c = codes[lo:hi].long()
        v = torch.stack([cb[m][c[:, m]] for m in range(M)], dim=1).reshape(hi - lo, D)
        v += 0.05 * torch.randn(v.shape, generator=g, device=dev)
        raw.add_device(v.contiguous())
    del v, c
    enc.attach_raw(raw)
    ms, sm = timed(lambda: enc.search_device(qs, K, KP, od, ol), steps=max(2, STEPS // 3), warm=1)
    report(f"C3 PQ ADC scan + fp32 re-rank k'=100 -> k=10, {N} rows, {Q} queries", ms, sm, Q, None, None,
           f"ADC lookups/s = {Q * N * M / (ms * 1e-3):.3e}")
    enc.close()
    raw.close()


def c4():
    g = torch.Generator(device=dev).manual_seed(4001)
    N, D, Q, K = 12_500_000, 128, 1024, 10
    db = torch.randint(-128, 128, (N, D), generator=g, device=dev, dtype=torch.int8)
    qs = torch.randint(-128, 128, (Q, D), generator=g, device=dev, dtype=torch.int8)
    idx = gpu.DenseIndex(D, np.int8, _lib.METRIC_DOT)
    idx.reserve(N)
    idx.add_device(db)
    od = torch.empty((Q, K), dtype=torch.float32, device=dev)
    ol = torch.empty((Q, K), dtype=torch.int64, device=dev)
    ms, sm = timed(lambda: idx.search_device(qs, K, od, ol))
    report("C4 int8 dot k=10, one GPU's shard 12.5M x 128 of 100M, 1024 queries", ms, sm, Q, N * D, 2.0 * Q * N * D,
           "integer MMA (kind::i8); tensor fraction quoted against the bf16 peak (int8 dense peak is 2x)")
    q1 = qs[:1].contiguous()
    od1 = torch.empty((1, K), dtype=torch.float32, device=dev)
    ol1 = torch.empty((1, K), dtype=torch.int64, device=dev)
    ms, sm = timed(lambda: idx.search_device(q1, K, od1, ol1))
    report("C4 single query (HBM-bound pass over the shard)", ms, sm, 1, N * D, None)
    for nq in (2, 4, 8, 32):
        qn = qs[:nq].contiguous()
        odn = torch.empty((nq, K), dtype=torch.float32, device=dev)
        oln = torch.empty((nq, K), dtype=torch.int64, device=dev)
        ms, sm = timed(lambda: idx.search_device(qn, K, odn, oln))
        report(f"C4 shard, {nq} queries per call (one padded tensor-core query block)", ms, sm, nq, N * D, None)
    idx.close()


def c5():
    g = torch.Generator(device=dev).manual_seed(5001)
    N, D, Q, C, K = int(os.environ.get("C5_N", 10_000_000)), 384, 4096, 128, 10
    idx = gpu.DenseIndex(D, np.float32, _lib.METRIC_L2)
    idx.reserve(N)
    step = 1_000_000
    for lo in range(0, N, step):
        idx.add_device(torch.randn((min(step, N - lo), D), generator=g, device=dev))
    qs = torch.randn((Q, D), generator=g, device=dev)
    cand = torch.randint(0, N, (Q, C), generator=g, device=dev, dtype=torch.int64).to(torch.uint32)
    tomb = (torch.rand(N, generator=g, device=dev) < 0.05).cpu().numpy()
    allow = (torch.rand(N, generator=g, device=dev) < 0.30).cpu().numpy()
    idx.set_tombstones(tomb)
    allow_d = torch.from_numpy(gpu.pack_bitmap(allow).view(np.int64)).to(dev)
    od = torch.empty((Q, K), dtype=torch.float32, device=dev)
    ol = torch.empty((Q, K), dtype=torch.int64, device=dev)
    ms, sm = timed(lambda: idx.rerank_device(qs, cand, K, od, ol, allow=allow_d))
    report("C5 HNSW re-rank: 4096 queries x 128 candidate ids, 10M x 384 fp32, tombstones 5% + allow 30%", ms, sm, Q,
           Q * C * D * 4 * 0.285 + Q * C * 4, None,
           "bytes = rows actually gathered (28.5% of candidates pass both bitmaps) + ids; "
           f"unfiltered-equivalent rate {Q * C * D * 4 / (ms * 1e-3) / 1e9:.0f} GB/s")
    idx.set_tombstones(None)
    ms, sm = timed(lambda: idx.rerank_device(qs, cand, K, od, ol))
    report("C5 re-rank without bitmaps (every candidate gathered)", ms, sm, Q, Q * C * D * 4 + Q * C * 4, None)
    idx.close()


if __name__ == "__main__":
    which = sys.argv[1:] or ["c1", "c2", "c3", "c4", "c5"]
    for w in which:
        t0 = time.time()
        {"c1": c1, "c2": c2, "c3": c3, "c4": c4, "c5": c5}[w]()
        torch.cuda.empty_cache()
        print(f"# {w} done in {time.time() - t0:.1f} s", file=sys.stderr, flush=True)
