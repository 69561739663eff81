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
elif stage == "daf":
    r = dev.upload(sc.srtm())
    for _ in range(2):
        cf.DetectApplyFourier().run_device(r)
elif stage == "chainfill":
    from hydrodem_b200 import _lib
    from hydrodem_b200.pipeline import ConditioningChain
    chain = ConditioningChain(with_hydrology=False)
    res = chain.run_device(*chain.upload_inputs(sc.srtm(), sc.groves(), sc.hsheds()))
    z = dev.convert(res.rasters["final"], _lib.F32, np.float32)
    for _ in range(2):
        nf.SinkFill(want_stats=False).run_device(z)
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
elif stage == "stencils":
    from hydrodem_b200.filters import extension_filters as ef
    from hydrodem_b200 import _lib
    srtm = dev.upload(sc.srtm()); hs = dev.upload(sc.hsheds())
    maj = cf.MajorityFilter(window_size=11).run_device(hs)
    dem64 = dev.convert(srtm, _lib.F64)
    for _ in range(2):
        cf.ExpandFilter(window_size=13).run_device(maj)
        ef.BinaryErosion(iterations=2).run_device(maj)
        ef.GreyDilation(size=(7, 7)).run_device(maj)
        cf.PostProcessingFinal().run_device(dem64)
        cf.PostProcessingFinal().run_device(srtm)
        nf.D8FlowDirection().run_device(srtm)
        nf.MedianFilter(window_size=3).run_device(srtm)
        cf.CorrectNANValues().run_device(hs)
        cf.QuadraticFilter(window_size=15).run_device(srtm)
        cf.DetectBlanksFourier().run_device(srtm.sub(0, 1790, 0, 1790))
torch.cuda.synchronize()
print("done", stage)
