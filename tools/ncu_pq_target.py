"""ncu target: PQ ADC coarse scan at C3 size (10M x 96 codes): three single-query searches and one 4-query search.
  python tools/ncu_pq_target.py [N]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from longbow_b200 import _lib, pq

dev = torch.device("cuda", 0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
g = torch.Generator(device=dev).manual_seed(3001)
M, D, K = 96, 768, 10
cb = torch.randn((M, 256, D // M), generator=g, device=dev)
codes = torch.randint(0, 256, (N, M), generator=g, device=dev, dtype=torch.uint8)
qs = torch.randn((8, D), generator=g, device=dev)
enc = pq.PQEncoder(D, M, 256, cb.cpu().numpy())
enc.add_codes_device(codes)
for nq in (1, 1, 1, 4):
    od = torch.empty((nq, K), dtype=torch.float32, device=dev)
    ol = torch.empty((nq, K), dtype=torch.int64, device=dev)
    enc.search_device(qs[:nq].contiguous(), K, 0, od, ol)
torch.cuda.synchronize()
print("ok", od[0, :3].tolist(), ol[0, :3].tolist())
