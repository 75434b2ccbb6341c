"""Probe: one query per call on the C2 database with a large k -- the ordinary path (streaming scan, kc candidates
per warp) against the exhaustive exact chain (lb_set_option("exhaustive_k", 1) forces it), answers compared; and the
cost of k beyond the fused selector (k > 704 always takes the exhaustive chain).  This is SearchHybrid's k * 10
candidate call (internal/store/hnsw_gpu.go:85)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from longbow_b200 import _lib, gpu
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
N, D = 1_000_000, 768
db = torch.randn((N, D), generator=g, device=dev); db = (db / db.norm(dim=1, keepdim=True)).half()
qs = torch.randn((4, D), generator=g, device=dev); qs = (qs / qs.norm(dim=1, keepdim=True)).half()
idx = gpu.DenseIndex(D, np.float16, _lib.METRIC_COSINE); idx.reserve(N); idx.add_device(db)

def run(K, nq, reps=8):
    q = qs[:nq].contiguous()
    od = torch.empty((nq, K), dtype=torch.float32, device=dev); ol = torch.empty((nq, K), dtype=torch.int64, device=dev)
    for _ in range(2): idx.search_device(q, K, od, ol)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): idx.search_device(q, K, od, ol)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, od.cpu().numpy(), ol.cpu().numpy()

for K in (100, 160, 224, 288, 352, 448, 512, 704):
    _lib.set_option("exhaustive_k", 100000)   # never (below the selector's limit)
    t0, d0, l0 = run(K, 1)
    _lib.set_option("exhaustive_k", 1)        # always
    t1, d1, l1 = run(K, 1)
    print("k", K, "ordinary ms", round(t0, 3), "exhaustive ms", round(t1, 3), "equal", bool(np.array_equal(d0, d1) and np.array_equal(l0, l1)), flush=True)
_lib.set_option("exhaustive_k", 0)
for K, nq in ((1000, 1), (2048, 1)):
    print("k", K, "queries", nq, "ms/call", round(run(K, nq, 4)[0], 3), flush=True)
