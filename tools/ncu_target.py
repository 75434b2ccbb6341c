"""Small driver for ncu captures: builds the C2 (or C4 / PQ) working set on the device and runs a few searches."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from longbow_b200 import _lib, gpu

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
if cfg.startswith("c2"):
    N, D, Q, K, metric, npdt = 1_000_000, 768, 1024, 100, _lib.METRIC_COSINE, np.float16
    if ":" in cfg:
        N = int(cfg.split(":")[1])  # c2:125000 = one rank's shard at 8 GPUs
    db = torch.randn((N, D), generator=g, device=dev)
    db = (db / db.norm(dim=1, keepdim=True)).half()
    qs = torch.randn((Q, D), generator=g, device=dev)
    qs = (qs / qs.norm(dim=1, keepdim=True)).half()
elif cfg == "c4":
    N, D, Q, K, metric, npdt = 12_500_000, 128, 1024, 10, _lib.METRIC_DOT, np.int8
    db = torch.randint(-128, 128, (N, D), generator=g, device=dev, dtype=torch.int8)
    qs = torch.randint(-128, 128, (Q, D), generator=g, device=dev, dtype=torch.int8)
else:
    raise SystemExit("c2 | c4")
idx = gpu.DenseIndex(D, npdt, metric)
idx.reserve(N)
idx.add_device(db)
od = torch.empty((Q, K), dtype=torch.float32, device=dev)
ol = torch.empty((Q, K), dtype=torch.int64, device=dev)
for _ in range(reps):
    idx.search_device(qs, K, od, ol)
torch.cuda.synchronize()
print("ok", int(ol[0, 0]))
