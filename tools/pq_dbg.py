import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from longbow_b200 import _lib, pq
dev = torch.device("cuda", 0)
N, M, D = 2_000_000, 96, 768
g = torch.Generator(device=dev).manual_seed(3001)
codes = torch.randint(0, 256, (N, M), generator=g, device=dev, dtype=torch.uint8)
cb = torch.randn((M, 256, D // M), generator=g, device=dev)
qs = torch.randn((64, D), generator=g, device=dev).cpu().numpy()
enc = pq.PQEncoder(D, M, 256, cb.cpu().numpy())
enc.add_codes_device(codes)
for mode in (2, 3):
    _lib.set_option("pq_scan", mode)
    for k in (10, 100, 200):
        d, l = enc.search(qs, k)
        print("mode", mode, "k", k, "uncertified", enc.last_uncertified(), "d[0,:3]", d[0, :3], "gap", d[0, k - 1] - d[0, 0])
