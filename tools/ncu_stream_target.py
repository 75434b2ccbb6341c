"""ncu driver: single-query searches (streaming scan) on the C2 database."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from longbow_b200 import _lib, gpu
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
N, D, K = 1_000_000, 768, 100
db = torch.randn((N, D), generator=g, device=dev)
db = (db / db.norm(dim=1, keepdim=True)).half()
q = torch.randn((1, D), generator=g, device=dev)
q = (q / q.norm(dim=1, keepdim=True)).half()
idx = gpu.DenseIndex(D, np.float16, _lib.METRIC_COSINE)
idx.reserve(N); idx.add_device(db)
od = torch.empty((1, K), dtype=torch.float32, device=dev); ol = torch.empty((1, K), dtype=torch.int64, device=dev)
for _ in range(3):
    idx.search_device(q, K, od, ol)
torch.cuda.synchronize()
print("ok", int(ol[0, 0]))
