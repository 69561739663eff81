"""Per-stage throughput on one GPU: Mcells/s and % of the measured HBM peak (BASELINE.md metric, per stage).

    python tools/stage_bench.py [--sizes 3601 14400] [--out profiles/r1_stages.json]

Inputs are resident in HBM; each stage is timed with CUDA events over `reps` back-to-back launches after a
warm-up (inputs > L2 at the larger size; at 3601^2 the 256 MB flush buffer is written between repetitions).
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hydrodem_b200 import _lib, device as dev                                  # noqa: E402
from hydrodem_b200.filters import custom_filters as cf, extension_filters as ef, new_filters as nf   # noqa: E402
from hydrodem_b200.synth import SynthScene                                     # noqa: E402

PEAK = 6550.4
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:      # noqa: BLE001
    pass


def timed(fn, reps, flush):
    fn(); fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        if flush is not None:
            flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", type=int, nargs="+", default=[3601, 14400])
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    results = []
    for n in args.sizes:
        sc = SynthScene(n, n, 1002)
        srtm = dev.upload(sc.srtm())
        hs = dev.upload(sc.hsheds())
        groves = ef.BinaryClosing(structure=np.ones((3, 3))).run_device(dev.upload(sc.groves()))
        maj = cf.MajorityFilter(window_size=11).run_device(hs)
        mask = ef.BinaryErosion(iterations=2).run_device(maj)
        dem64 = dev.convert(srtm, _lib.F64)
        rounded = dev.upload(np.round(sc.srtm()))
        filled = nf.SinkFill(want_stats=False).run_device(rounded)
        cells = n * n
        gc = cf.GrovesCorrection(groves)
        fill = nf.SinkFill(want_stats=True)
        stages = [
            ("expand 7 (u8 -> u8)", 2, lambda: cf.ExpandFilter(window_size=7).run_device(mask)),
            ("expand 13 (f32 -> u8)", 5, lambda: cf.ExpandFilter(window_size=13).run_device(maj)),
            ("majority 11", 8, lambda: cf.MajorityFilter(window_size=11).run_device(hs)),
            ("nan-correction 3", 8, lambda: cf.CorrectNANValues().run_device(hs)),
            ("quadratic 15 (isotropic)", 8, lambda: cf.QuadraticFilter(window_size=15).run_device(srtm)),
            ("groves correction (quadratic 15 + tail)", 9, lambda: gc.run_device(srtm, out_dtype=_lib.F32)),
            ("median 3", 8, lambda: nf.MedianFilter(window_size=3).run_device(srtm)),
            ("median 5", 8, lambda: nf.MedianFilter(window_size=5).run_device(srtm)),
            ("binary erosion x2 (f32 -> u8)", 5, lambda: ef.BinaryErosion(iterations=2).run_device(maj)),
            ("grey dilation 7x7 (max)", 8, lambda: ef.GreyDilation(size=(7, 7)).run_device(maj)),
            ("mean 3x3 + round (f64)", 16, lambda: cf.PostProcessingFinal().run_device(dem64)),
            ("mean 3x3 + round (f32)", 8, lambda: cf.PostProcessingFinal().run_device(srtm)),
            ("d8", 5, lambda: nf.D8FlowDirection().run_device(filled)),
            ("sink-fill (async worklist)", None, lambda: fill.run_device(rounded)),
        ]
        if n <= 8192:
            daf = cf.DetectApplyFourier()
            stages += [("fft2 + shift + abs", 32, lambda: cf.FourierInitial().run_device(srtm)),
                       ("stripe removal (fft2, mask, ifft2)", 64 + 20, lambda: daf.run_device(srtm))]
        for name, bpc, fn in stages:
            ms = timed(fn, args.reps, flush if cells * 8 < (200 << 20) else None)
            row = {"size": n, "stage": name, "ms": ms, "mcells_s": cells / ms / 1e3}
            if name.startswith("sink-fill"):
                visits = fill.sweeps
                row["tile_visits"] = visits
                bpc_eff = 12.0 * visits * 4096 / cells
                row["bytes_per_cell"] = bpc_eff
                row["gbs"] = bpc_eff * cells / ms / 1e6
            else:
                row["bytes_per_cell"] = bpc
                row["gbs"] = bpc * cells / ms / 1e6
            row["pct_hbm"] = 100.0 * row["gbs"] / PEAK
            results.append(row)
            print(f"{n:6d}  {name:42s} {ms:9.3f} ms  {row['mcells_s']:10.0f} Mcells/s  {row['gbs']:8.0f} GB/s  "
                  f"{row['pct_hbm']:5.1f} % HBM", flush=True)
        del srtm, hs, groves, maj, mask, dem64, rounded, filled
        torch.cuda.empty_cache()
    if args.out:
        json.dump({"peak_gbs": PEAK, "stages": results}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
