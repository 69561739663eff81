import sys, os, torch, numpy as np
sys.path.insert(0, '/root/repo')
from hydrodem_b200 import _lib, device as dev
from hydrodem_b200.filters import custom_filters as cf
n = 17990
src = dev.empty(n, n, _lib.F32)
t = src.tensor()
g = torch.Generator(device='cuda').manual_seed(1)
for r0 in range(0, n, 2048):
    r1 = min(n, r0 + 2048)
    a = torch.randn((r1 - r0, n), generator=g, device='cuda'); b = torch.randn((r1 - r0, n), generator=g, device='cuda')
    t[r0:r1] = torch.sqrt(a * a + b * b) * 100
lib = _lib.load()
flags = torch.empty(int(lib.hd_hollow_tile_count(n, n)), dtype=torch.uint8, device='cuda')
def timed(fn):
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); r = fn(); b.record(); torch.cuda.synchronize(); return r, a.elapsed_time(b)
for rep in range(2):
    (mask, mod), t1 = timed(lambda: cf._hollow_pass(src, None, 55, flags_out=flags))
    (m2, _), t2 = timed(lambda: cf._hollow_pass(mod, mask, 55, last=True, flags_in=flags))
    (m3, _), t3 = timed(lambda: cf._hollow_pass(mod, mask, 55, last=True))
    print(f"pass1 {t1:.2f} ms  pass2 flagged {t2:.2f} ms  pass2 dense {t3:.2f} ms  flagged tiles {int(flags.sum())} of {flags.numel()}  hits {int(mask.tensor().sum())}  equal {bool(torch.equal(m2.tensor(), m3.tensor()))}")
