"""ncu target: batched PQ search at C3 size through the decode + tensor-core coarse stage (Q queries, two searches).
  python tools/ncu_pq_batch_target.py [Q] [N]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from longbow_b200 import _lib, pq

dev = torch.device("cuda", 0)
Q = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
g = torch.Generator(device=dev).manual_seed(3001)
M, D, K = 96, 768, 10
cb = torch.randn((M, 256, D // M), generator=g, device=dev)
codes = torch.randint(0, 256, (N, M), generator=g, device=dev, dtype=torch.uint8)
qs = torch.randn((Q, D), generator=g, device=dev)
enc = pq.PQEncoder(D, M, 256, cb.cpu().numpy())
enc.add_codes_device(codes)
od = torch.empty((Q, K), dtype=torch.float32, device=dev)
ol = torch.empty((Q, K), dtype=torch.int64, device=dev)
for _ in range(2):
    enc.search_device(qs, K, 100, od, ol)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    enc.search_device(qs, K, 100, od, ol)
e1.record(); torch.cuda.synchronize()
print("ok ms/batch", e0.elapsed_time(e1) / 3, "uncertified", enc.last_uncertified())
