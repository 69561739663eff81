"""Time hd_majority (window 11) on 18000^2 rasters of different content: which part of the kernel costs what.
Usage (GPU box): python tools/majority_time.py [n]"""
import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
from hydrodem_b200 import _lib, device as dev
from hydrodem_b200.filters import custom_filters as cf
from hydrodem_b200.synth import DeviceMosaic
n = int(sys.argv[1]) if len(sys.argv) > 1 else 18000
mos = DeviceMosaic(n, n, 1005)
src = dev.empty(n, n, _lib.F32)
t = src.tensor()
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): r = fn()
    b.record(); torch.cuda.synchronize(); return r, a.elapsed_time(b) / reps
def fill(kind):
    for r0 in range(0, n, 3000):
        r1 = min(n, r0 + 3000)
        if kind == "synthetic hsheds":
            mos.band(r0, r1, out={"hsheds": t[r0:r1]})
        elif kind == "textured, no lagoons":
            mos._fill(r0, r1, None, t[r0:r1])
        elif kind == "constant":
            t[r0:r1] = 7.0
        elif kind == "all distinct":
            t[r0:r1] = torch.rand((r1 - r0, n), device='cuda')
        elif kind == "two values 50/50":
            t[r0:r1] = (torch.rand((r1 - r0, n), device='cuda') < 0.5).float() + 1
        elif kind == "two values 90/10":
            t[r0:r1] = (torch.rand((r1 - r0, n), device='cuda') < 0.1).float() + 1
f = cf.MajorityFilter(window_size=11)
for kind in ("synthetic hsheds", "textured, no lagoons", "constant", "all distinct", "two values 50/50", "two values 90/10"):
    fill(kind)
    out, ms = timed(lambda: f.run_device(src))
    nz = float((out.tensor() != 0).float().mean())
    print(f"{kind:24s} {ms:7.3f} ms   {n * n / ms / 1e6:7.1f} Gcells/s   non-zero results {nz:.4f}", flush=True)
