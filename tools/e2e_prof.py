import cProfile, pstats, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from hydrodem_b200 import device as dev
from hydrodem_b200.pipeline import ConditioningChain
from hydrodem_b200.synth import SynthScene
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3601
sc = SynthScene(n, n, 1002)
host = {}
for k, a in dict(srtm=sc.srtm(), groves=sc.groves(), hsheds=sc.hsheds()).items():
    p = dev.pinned_empty(a.shape, a.dtype); p[...] = a; host[k] = p
chain = ConditioningChain()
def tiles(m):
    for _ in range(m):
        yield (host["srtm"], host["groves"], host["hsheds"])
for r in chain.stream(tiles(4)):
    del r
torch.cuda.synchronize()
for depth in (1, 2, 3, 4):
    for r in chain.stream(tiles(4), depth=depth):
        del r
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for r in chain.stream(tiles(20), depth=depth):
        del r
    torch.cuda.synchronize()
    print("depth", depth, "ms/step", (time.perf_counter() - t0) / 20 * 1e3)
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
for r in chain.stream(tiles(20)):
    del r
torch.cuda.synchronize()
pr.disable()
import os; print("cpus", len(os.sched_getaffinity(0)), os.cpu_count()); print("ms/step", (time.perf_counter() - t0) / 20 * 1e3)
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
