import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from hydrodem_b200 import device as dev
from hydrodem_b200.filters import new_filters as nf
from hydrodem_b200.synth import SynthScene
sc = SynthScene(3601, 3601, 1002)
r = dev.upload(np.round(sc.srtm()))
f = nf.SinkFill(want_stats=False)
for _ in range(3): f.run_device(r)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5): f.run_device(r)
b.record(); torch.cuda.synchronize()
print("fill ms", a.elapsed_time(b) / 5)
