"""ncu target: the GPU HNSW walk (csrc/hnsw.cu) on a 200k x 384 fp32 graph (24 kNN + 8 random edges), 2048 queries,
ef=128, then the re-rank.  python tools/ncu_hnsw_target.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from longbow_b200 import _lib, gpu, store

dev = torch.device("cuda", 0)
N, D, Q, EF, K, DEG, KNN = 200_000, 384, int(os.environ.get("Q", "2048")), 128, 10, 32, 24
g = torch.Generator(device=dev).manual_seed(5101)
db = torch.randn((N, D), generator=g, device=dev)
idx = gpu.DenseIndex(D, np.float32, _lib.METRIC_L2)
idx.add_device(db)
nbrs = torch.empty((N, DEG), dtype=torch.int64, device=dev)
od = torch.empty((4096, KNN + 1), dtype=torch.float32, device=dev)
ol = torch.empty((4096, KNN + 1), dtype=torch.int64, device=dev)
for lo in range(0, N, 4096):
    hi = min(N, lo + 4096)
    idx.search_device(db[lo:hi], KNN + 1, od[:hi - lo], ol[:hi - lo])
    nbrs[lo:hi, :KNN] = ol[:hi - lo, 1:]
nbrs[:, KNN:] = torch.randint(0, N, (N, DEG - KNN), generator=g, device=dev)
graph = store.HNSWGraph(idx, DEG)
nb32 = nbrs.to(torch.uint32).contiguous()
_lib.check(_lib.load().lb_graph_set_layer_device(graph._h, nb32.data_ptr(), None, N, torch.cuda.current_stream().cuda_stream))
qs = torch.randn((Q, D), generator=g, device=dev)
entries = torch.zeros(Q, dtype=torch.int32, device=dev)
out_d = torch.empty((Q, K), dtype=torch.float32, device=dev)
out_l = torch.empty((Q, K), dtype=torch.int64, device=dev)
fail = torch.zeros(1, dtype=torch.int32, device=dev)
for _ in range(2):
    graph.search_device(qs, entries, EF, K, out_d, out_l, fail)
torch.cuda.synchronize()
print("ok", int(fail.item()), out_l[0, :3].tolist())
