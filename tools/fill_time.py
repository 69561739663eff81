"""Time the sink-fill alone on the chain's own final DEM (HD_FILL_TRACE=1 prints the per-phase cycle split)."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from hydrodem_b200 import _lib, device as dev
from hydrodem_b200.filters import new_filters as nf
from hydrodem_b200.pipeline import ConditioningChain
from hydrodem_b200.synth import SynthScene
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3601
sc = SynthScene(n, n, 1002)
chain = ConditioningChain(with_hydrology=False)
res = chain.run_device(*chain.upload_inputs(sc.srtm(), sc.groves(), sc.hsheds()))
r = dev.convert(res.rasters["final"], _lib.F32, np.float32)
f = nf.SinkFill(want_stats=False)
for _ in range(3): f.run_device(r)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5): f.run_device(r)
b.record(); torch.cuda.synchronize()
print("fill ms", a.elapsed_time(b) / 5)
