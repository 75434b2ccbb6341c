"""Power / clock probe of the C2 scan variants: each variant runs back to back for ~SECS seconds while NVML is sampled.
Prints ms per batch, scan-kernel ms, median SM clock and median board power -- to tell a cycle-bound kernel from one
that sits on the power cap (where fewer cycles only lower the clock)."""
import os, sys, time, threading, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pynvml as nv
from longbow_b200 import _lib, gpu
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
SECS = float(os.environ.get("SECS", "2.5"))
WORK = os.environ.get("WORK", "c2")
if WORK == "c2":
    N, D, NQ, K = int(os.environ.get("ROWS", "1000000")), 768, 1024, 100
    db = torch.randn((N, D), generator=g, device=dev); db = (db / db.norm(dim=1, keepdim=True)).half()
    qs = torch.randn((NQ, D), generator=g, device=dev); qs = (qs / qs.norm(dim=1, keepdim=True)).half()
    idx = gpu.DenseIndex(D, np.float16, _lib.METRIC_COSINE)
else:  # c4: one rank's shard of the int8 dot-product config
    N, D, NQ, K = 12_500_000, 128, 1024, 10
    db = torch.randint(-128, 128, (N, D), generator=g, device=dev, dtype=torch.int8)
    qs = torch.randint(-128, 128, (NQ, D), generator=g, device=dev, dtype=torch.int8)
    idx = gpu.DenseIndex(D, np.int8, _lib.METRIC_DOT)
idx.reserve(N); idx.add_device(db)
od = torch.empty((NQ, K), dtype=torch.float32, device=dev); ol = torch.empty((NQ, K), dtype=torch.int64, device=dev)
nv.nvmlInit(); h = nv.nvmlDeviceGetHandleByIndex(0)
class S(threading.Thread):
    def __init__(s): super().__init__(daemon=True); s.stop = False; s.mhz = []; s.w = []
    def run(s):
        while not s.stop:
            s.mhz.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)); s.w.append(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
            time.sleep(0.02)
# cases: tc_debug values (0 full, 1 no epilogue, 32 filter only, 128 round-1 release form) or "mm" (cuBLAS bf16 8192^3)
cases = (sys.argv[1] if len(sys.argv) > 1 else "0,1,32,128,mm,0").split(",")
a = torch.randn((8192, 8192), device=dev, dtype=torch.bfloat16); b = torch.randn((8192, 8192), device=dev, dtype=torch.bfloat16)
for dbg in cases:
    mm = dbg == "mm"
    if dbg.startswith("boot"):      # "boot8": tc_boot_tiles = 8 (0 = automatic)
        _lib.set_option("tc_debug", 0); _lib.set_option("tc_boot_tiles", int(dbg[4:]))
    elif dbg.startswith("rsb"):     # "rsb1": rescore_block = 1 (block-per-query exact stage), "rsb0": warp-per-query
        _lib.set_option("tc_debug", 0); _lib.set_option("rescore_block", int(dbg[3:]))
    elif not mm:
        _lib.set_option("tc_debug", int(dbg))
    run = (lambda: torch.matmul(a, b)) if mm else (lambda: idx.search_device(qs, K, od, ol))
    for _ in range(5): run()
    torch.cuda.synchronize()
    s = S(); s.start()
    _lib.prof_read(True); _lib.prof_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time(); n = 0
    e0.record()
    while time.time() - t0 < SECS:
        for _ in range(20): run()
        n += 20
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize(); _lib.prof_enable(False); s.stop = True; s.join()
    ms, cnt, _u = _lib.prof_read(True)
    half = len(s.mhz) // 2   # second half: settled
    per = e0.elapsed_time(e1) / n
    extra = f"TF/s {2 * 8192**3 / per / 1e9:.0f}" if mm else f"scan_ms {ms / max(cnt, 1):.4f}"
    print(f"case {dbg} ms/iter {per:.4f} {extra} sm_mhz {statistics.median(s.mhz[half:])} "
          f"power_w {statistics.median(s.w[half:]):.0f} max_w {max(s.w):.0f}", flush=True)
_lib.set_option("tc_debug", 0); _lib.set_option("tc_boot_tiles", 0); _lib.set_option("rescore_block", 0)
