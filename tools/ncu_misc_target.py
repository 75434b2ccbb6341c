"""ncu driver for the secondary kernels: PQ ADC scan (2M x 96 codes, 64 queries), int8 tensor-core scan
(2M x 128, 1024 queries) and the cooperative re-rank (1M x 384 fp32, 1024 x 128 candidates)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from longbow_b200 import _lib, gpu, pq
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(7)
# PQ
N, D, M = 2_000_000, 768, 96
cb = torch.randn((M, 256, D // M), generator=g, device=dev).cpu().numpy()
codes = torch.randint(0, 256, (N, M), generator=g, device=dev, dtype=torch.uint8)
enc = pq.PQEncoder(D, M, 256, cb)
enc.add_codes_device(codes)
qs = torch.randn((64, D), generator=g, device=dev)
od = torch.empty((64, 10), dtype=torch.float32, device=dev); ol = torch.empty((64, 10), dtype=torch.int64, device=dev)
for _ in range(2):
    enc.search_device(qs, 10, 0, od, ol)
torch.cuda.synchronize()
# int8
N8 = 2_000_000
db8 = torch.randint(-128, 128, (N8, 128), generator=g, device=dev, dtype=torch.int8)
q8 = torch.randint(-128, 128, (1024, 128), generator=g, device=dev, dtype=torch.int8)
i8 = gpu.DenseIndex(128, np.int8, _lib.METRIC_DOT); i8.reserve(N8); i8.add_device(db8)
od8 = torch.empty((1024, 10), dtype=torch.float32, device=dev); ol8 = torch.empty((1024, 10), dtype=torch.int64, device=dev)
for _ in range(2):
    i8.search_device(q8, 10, od8, ol8)
torch.cuda.synchronize()
# re-rank
NR = 1_000_000
f32 = gpu.DenseIndex(384, np.float32, _lib.METRIC_L2); f32.reserve(NR)
f32.add_device(torch.randn((NR, 384), generator=g, device=dev))
qr = torch.randn((1024, 384), generator=g, device=dev)
cand = torch.randint(0, NR, (1024, 128), generator=g, device=dev, dtype=torch.int64).to(torch.uint32)
odr = torch.empty((1024, 10), dtype=torch.float32, device=dev); olr = torch.empty((1024, 10), dtype=torch.int64, device=dev)
for _ in range(2):
    f32.rerank_device(qr, cand, 10, odr, olr)
torch.cuda.synchronize()
print("ok")
