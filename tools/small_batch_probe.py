"""Probe: 1..8 queries per call on the C2 database -- streaming scan (dense_scan=3) vs the auto policy."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from longbow_b200 import _lib, gpu
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
kind = sys.argv[1] if len(sys.argv) > 1 else "c2"
if kind == "c2":
    N, D, K, metric, npdt = 1_000_000, 768, 100, _lib.METRIC_COSINE, np.float16
    db = torch.randn((N, D), generator=g, device=dev); db = (db / db.norm(dim=1, keepdim=True)).half()
    qs = torch.randn((8, D), generator=g, device=dev); qs = (qs / qs.norm(dim=1, keepdim=True)).half()
else:
    N, D, K, metric, npdt = 12_500_000, 128, 10, _lib.METRIC_DOT, np.int8
    db = torch.randint(-128, 128, (N, D), generator=g, device=dev, dtype=torch.int8)
    qs = torch.randint(-128, 128, (8, D), generator=g, device=dev, dtype=torch.int8)
idx = gpu.DenseIndex(D, npdt, metric); idx.reserve(N); idx.add_device(db)
for mode in (0, 3):
    _lib.set_option("dense_scan", mode)
    for nq in (1, 2, 4, 8):
        q = qs[:nq].contiguous()
        od = torch.empty((nq, K), dtype=torch.float32, device=dev); ol = torch.empty((nq, K), dtype=torch.int64, device=dev)
        for _ in range(3): idx.search_device(q, K, od, ol)
        torch.cuda.synchronize()
        _lib.prof_read(True); _lib.prof_enable(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): idx.search_device(q, K, od, ol)
        e1.record(); torch.cuda.synchronize(); _lib.prof_enable(False)
        ms, n, _u = _lib.prof_read(True)
        print(kind, "mode", mode, "nq", nq, "ms/call", round(e0.elapsed_time(e1) / 20, 4), "scan_ms", round(ms / max(n, 1), 4))
_lib.set_option("dense_scan", 0)
