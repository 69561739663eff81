"""Small cases for compute-sanitizer (racecheck / memcheck / synccheck): the asynchronous sink-fill worklist (inter-CTA
polling protocol), the TMA tile_loop kernels, the cluster / DSMEM FFT, the river wavefront and the streaming slots, on a
700 x 900 raster.

    compute-sanitizer --tool racecheck python tools/sanitize_case.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hydrodem_b200.filters import custom_filters as cf, new_filters as nf     # noqa: E402
from hydrodem_b200.pipeline import ConditioningChain                           # noqa: E402
from hydrodem_b200.synth import SynthScene                                     # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
sc = SynthScene(700, 900, 31)
srtm, groves, hsheds = sc.srtm(), sc.groves(), sc.hsheds()
if which in ("all", "fill"):
    z = np.round(srtm)
    z[100:104, 200:230] = np.nan
    f = nf.SinkFillD8(want_stats=True)
    filled, d8 = f.apply(z)
    print("fill: visits", f.sweeps, "status", f.status(), "checksum", float(np.nansum(filled)), int(d8.sum()))
if which in ("all", "chain"):
    out = ConditioningChain().apply(srtm, groves, hsheds.copy())
    print("chain: final checksum", float(out.final.sum()), "d8", int(out.d8.sum()))
if which in ("all", "longfft"):
    a = np.random.default_rng(1).normal(100, 5, (12, 9000)).astype(np.float32)      # 9000 = 2 x 4500: cluster + DSMEM path
    from hydrodem_b200.filters import extension_filters as ef
    print("long fft:", complex(ef.FourierTransform().apply(a)[3, 17]))
if which in ("all", "rivers"):
    mask = (np.random.default_rng(2).random((700, 900)) < 0.3).astype(np.float32)
    print("rivers:", float(cf.RouteRivers(window_size=3, dem=hsheds).apply(mask).sum()))
if which in ("all", "stream"):
    chain = ConditioningChain()
    n = 0
    for r in chain.stream([(srtm, groves, hsheds)] * 3, depth=2):
        n += int(r["d8"].sum())
    print("stream: d8 checksum", n)
