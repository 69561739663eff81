import cProfile, pstats, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from hydrodem_b200 import device as dev
from hydrodem_b200.pipeline import ConditioningChain
from hydrodem_b200.synth import SynthScene
n = 3601
DEPTH = int(sys.argv[1]) if len(sys.argv) > 1 else 2
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
for r in chain.stream(tiles(20), depth=DEPTH):
    del r
torch.cuda.synchronize()
pr.disable()
import os; print("cpus", len(os.sched_getaffinity(0)), os.cpu_count()); print("ms/step", (time.perf_counter() - t0) / 20 * 1e3)
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)

chain._trace = []
for r in chain.stream(tiles(12), depth=DEPTH):
    del r
torch.cuda.synchronize()
tr = chain._trace
print("compute span per tile (ms):", [round(a.elapsed_time(b), 2) for a, b, _ in tr])
for _, _, m in tr[4:8]:
    print("   segments fourier/groves/combine/hydrology (ms):", [round(m[i].elapsed_time(m[i + 1]), 2) for i in range(len(m) - 1)])
print("gap to next tile (ms):", [round(tr[i][1].elapsed_time(tr[i + 1][0]), 2) for i in range(len(tr) - 1)])
