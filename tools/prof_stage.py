"""Run one stage of the chain a few times on a 3601x3601 synthetic tile (for ncu captures)."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from hydrodem_b200 import device as dev
from hydrodem_b200.filters import custom_filters as cf, new_filters as nf
from hydrodem_b200.synth import SynthScene

stage = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3601
sc = SynthScene(n, n, 1002)
if stage == "fft":
    r = dev.upload(sc.srtm())
    for _ in range(3):
        cf.FourierInitial().run_device(r)
elif stage == "fill":
    r = dev.upload(np.round(sc.srtm()))
    for _ in range(2):
        f = nf.SinkFill(); f.run_device(r); print("sweeps", f.sweeps)
elif stage == "quadratic":
    r = dev.upload(sc.srtm())
    for _ in range(3):
        cf.QuadraticFilter(window_size=15).run_device(r)
elif stage == "majority":
    r = dev.upload(sc.hsheds())
    for _ in range(3):
        cf.MajorityFilter(window_size=11).run_device(r)
torch.cuda.synchronize()
print("done", stage)
