"""Timing probes of the tensor-core scan (debug knobs make results invalid; timing only)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from longbow_b200 import _lib, gpu

N, D, Q, K = 1_000_000, 768, 1024, int(os.environ.get("K", "100"))
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
db = torch.randn((N, D), generator=g, device=dev)
db = (db / db.norm(dim=1, keepdim=True)).half()
qs = torch.randn((Q, D), generator=g, device=dev)
qs = (qs / qs.norm(dim=1, keepdim=True)).half()
idx = gpu.DenseIndex(D, np.float16, _lib.METRIC_COSINE)
idx.reserve(N); idx.add_device(db)
od = torch.empty((Q, K), dtype=torch.float32, device=dev); ol = torch.empty((Q, K), dtype=torch.int64, device=dev)
for mode in [int(x) for x in (sys.argv[1:] or ["0"])]:
    _lib.set_option("tc_debug", mode)
    for _ in range(3): idx.search_device(qs, K, od, ol)
    torch.cuda.synchronize()
    _lib.prof_read(True); _lib.prof_enable(True)
    for _ in range(10): idx.search_device(qs, K, od, ol)
    torch.cuda.synchronize(); _lib.prof_enable(False)
    ms, n, _u = _lib.prof_read(True)
    print(f"tc_debug={mode}: scan {ms/n:.3f} ms  ({2*Q*N*D/(ms/n*1e-3)/1e12:.0f} TFLOP/s)", flush=True)
_lib.set_option("tc_debug", 0)
