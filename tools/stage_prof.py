"""Warm per-kernel timings of one search step (torch.profiler / CUPTI sees every kernel of the process,
including the ones liblongbow_b200.so launches).  Usage: python tools/stage_prof.py [c2|c4|c1] [debug modes...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

from longbow_b200 import _lib, gpu

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
modes = [int(x) for x in sys.argv[2:]] or [0]
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
if cfg == "c2":
    N, D, Q, K, metric, npdt = int(os.environ.get("N", 1_000_000)), 768, 1024, 100, _lib.METRIC_COSINE, np.float16
    db = torch.randn((N, D), generator=g, device=dev)
    db = (db / db.norm(dim=1, keepdim=True)).half()
    qs = torch.randn((Q, D), generator=g, device=dev)
    qs = (qs / qs.norm(dim=1, keepdim=True)).half()
elif cfg == "c4":
    N, D, Q, K, metric, npdt = int(os.environ.get("N", 12_500_000)), 128, 1024, 10, _lib.METRIC_DOT, np.int8
    db = torch.randint(-128, 128, (N, D), generator=g, device=dev, dtype=torch.int8)
    qs = torch.randint(-128, 128, (Q, D), generator=g, device=dev, dtype=torch.int8)
elif cfg == "c1":
    N, D, Q, K, metric, npdt = 100_000, 128, 1000, 10, _lib.METRIC_L2, np.float32
    db = torch.rand((N, D), generator=g, device=dev)
    qs = torch.rand((Q, D), generator=g, device=dev)
else:
    raise SystemExit("config: c1 | c2 | c4")
idx = gpu.DenseIndex(D, npdt, metric)
idx.reserve(N)
idx.add_device(db)
od = torch.empty((Q, K), dtype=torch.float32, device=dev)
ol = torch.empty((Q, K), dtype=torch.int64, device=dev)
STEPS = 10
if os.environ.get("BOOT"):
    _lib.set_option("tc_boot_tiles", int(os.environ["BOOT"]))
for mode in modes:
    _lib.set_option("tc_debug", mode)
    for _ in range(3):
        idx.search_device(qs, K, od, ol)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(STEPS):
        idx.search_device(qs, K, od, ol)
    e1.record()
    torch.cuda.synchronize()
    print(f"== {cfg} tc_debug={mode}: {e0.elapsed_time(e1) / STEPS:.3f} ms/step", flush=True)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(STEPS):
            idx.search_device(qs, K, od, ol)
        torch.cuda.synchronize()
    rows = []
    for ev in prof.key_averages():
        t = getattr(ev, "device_time_total", None)
        if t is None:
            t = getattr(ev, "cuda_time_total", 0)
        if t:
            rows.append((t / STEPS, ev.count / STEPS, ev.key[:90]))
    for t, c, name in sorted(rows, reverse=True):
        print(f"   {t:9.1f} us/step  x{c:4.1f}  {name}")
_lib.set_option("tc_debug", 0)
