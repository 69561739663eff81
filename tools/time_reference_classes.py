"""Time the reference's OWN filter classes (unmodified, imported from /root/reference) on a small synthetic crop of
the bench workload, in the BUILD container (the reference tree does not exist on the GPU box), and record the result
for bench.py's cpu_baseline.reference_classes.

    python tools/time_reference_classes.py [size]        -> profiles/r2_reference_classes_cpu.json

The chain is HydroDEMProcess.start (hydro_dem_process.py:122-153) without the GDAL file handling: DetectApplyFourier,
BinaryClosing, GrovesCorrectionsIter(3), LagoonsDetection, the sum of the final terms, PostProcessingFinal.
"""
import json
import os
import platform
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference/cguerrero/hydrodem")
from filters.custom_filters import (DetectApplyFourier, GrovesCorrectionsIter, LagoonsDetection,      # noqa: E402
                                    PostProcessingFinal)
from filters.extension_filters import BinaryClosing                                                 # noqa: E402
from hydrodem_b200.synth import SynthScene                                                          # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 256
sc = SynthScene(size, size, 1005)
srtm, groves, hsheds = sc.srtm(), sc.groves(), sc.hsheds()
stages = {}
t_all = time.perf_counter()
with np.errstate(all="ignore"):
    t = time.perf_counter(); dem = DetectApplyFourier().apply(srtm); stages["DetectApplyFourier"] = time.perf_counter() - t
    t = time.perf_counter(); gc = BinaryClosing(structure=np.ones((3, 3))).apply(groves)
    dem = GrovesCorrectionsIter(gc, iterations=3).apply(dem); stages["GrovesCorrectionsIter"] = time.perf_counter() - t
    t = time.perf_counter(); ld = LagoonsDetection(); mask_l = ld.apply(hsheds.copy()); stages["LagoonsDetection"] = time.perf_counter() - t
    t = time.perf_counter()
    complete = dem * (1 - mask_l) + ld.lagoons_values                   # _prepare_final_terms without rivers (:60-91)
    final = PostProcessingFinal().apply(complete); stages["combine+PostProcessingFinal"] = time.perf_counter() - t
total = time.perf_counter() - t_all
rec = {"mcells_s": size * size / total / 1e6, "seconds": total, "stages_s": stages,
       "sample": f"the reference's own classes (cguerrero/hydrodem/filters, unmodified) on one {size}x{size} synthetic tile, "
                 "1 process (they are single-threaded pure Python)",
       "where": f"build container, {platform.processor() or platform.machine()}, {os.cpu_count()} vCPU; "
                "recorded by tools/time_reference_classes.py", "final_checksum": float(final.sum())}
print(json.dumps(rec, indent=1))
with open(os.path.join(REPO, "profiles", "r2_reference_classes_cpu.json"), "w") as f:
    json.dump(rec, f, indent=1)
