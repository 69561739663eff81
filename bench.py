#!/usr/bin/env python
"""bench.py -- Mcells/s of the full HydroDEM conditioning chain on B200.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): the full conditioning chain on one synthetic 3601 x 3601 SRTM 1-arcsec
tile per GPU -- Fourier stripe removal, groves correction x3, lagoon detection, recombination, 3x3 mean + round,
sink-fill, D8.  A "step" is one pass of the chain over one tile.  At N > 1 every rank conditions its own tile
(seed 1002 + rank): tiles are independent, there is no data-path collective, scaling is weak.

Numbers on the JSON line
  value     whole-job Mcells/s with the inputs resident in HBM (CUDA events on the launching stream, max over ranks)
  e2e       the same through the public API with HOST buffers: `for out in ConditioningChain.stream(tiles)` -- every
            step uploads its three input rasters from pinned host memory and hands back final DEM / filled DEM / D8
            as host arrays, all inside the timed region; copies of neighbouring steps overlap the kernels.  The
            latency of ONE tile through ConditioningChain.apply_to_host is reported next to it.
  roofline  dominant kernel of the step: algorithmic bytes per launch / average launch time, measured live with a
            CUDA event pair around every launch (hd_profile_*), against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the CPU oracle port (oracle/chain.py) timed on rank 0, one core, on a bounded sample tile
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

TILE = 3601
SEED = 1002
METRIC = "Mcells/s, full conditioning chain"
UNIT = "Mcells/s"

# Algorithmic HBM bytes per cell and per launch of each kernel in this workload.
# fft_rows: the four row passes of one forward + one masked inverse 2-D transform of an odd x odd raster are
#   real->c64 (4+8), c64->c64 on the Hermitian half (8+8)/2, masked c64->c64 on half the rows (8+1+8)/2,
#   c64 pairs->|re| (8+4)  = 40.5 B/cell over 4 launches.
# fill_async: 12 B per cell of every tile VISIT (z + W read, W written) -- bench.py multiplies by the visit count.
ALGO_BYTES_PER_CELL = {
    "fft_rows_kernel": 40.5 / 4.0,
    "transpose_kernel": 16.0,
    "transpose_real_kernel": 8.0,
    "quadratic_kernel": 9.0,
    "majority_kernel": 8.0,
    "fill_sweep_kernel": 12.0,
    "fill_async_kernel": 12.0,
    "hollow_kernel": 9.0 * 0.25,        # runs on a spectrum quarter
    "expand_kernel": 2.0,
    "morph_kernel": 2.0,
    "maxfilter_kernel": 8.0,
    "fix3_kernel": 8.0,
    "conv3_kernel": 16.0,
    "final_terms_kernel": 20.0,
    "elementwise_kernel": 9.0,
    "d8_kernel": 5.0,
}


def env_int(name, default):
    return int(os.environ.get(name, default))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self):
        """Drop what was sampled so far (warm-up); keep sampling."""
        self.first = len(self.lines)

    def count(self):
        return len(self.lines) - getattr(self, "first", 0) if self.proc else 1 << 30

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.lines[getattr(self, "first", 0):]:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); smax.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except (OSError, KeyError, ValueError):
        return 6650.0, "fallback"          # /opt/skills/guides/B200_PROFILING.md


# ---------------------------------------------------------------------------------------------------------
def cpu_chain_sample(size, seed):
    """One pass of the CPU oracle chain over a size x size sample tile -> seconds."""
    from hydrodem_b200.synth import SynthScene
    from oracle import chain
    sc = SynthScene(size, size, seed)
    a, g, h = sc.srtm(), sc.groves(), sc.hsheds()
    t0 = time.perf_counter()
    with np.errstate(all="ignore"):
        chain.conditioning_chain(a, g, h)
    return time.perf_counter() - t0


def _cpu_worker(args):
    return cpu_chain_sample(*args)


def run_reference(args):
    """The reference arm: the CPU implementation of the path (the oracle port of the reference's NumPy/SciPy
    filters -- the reference itself is Python + GDAL and cannot be installed here) on all host cores."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle import clib
    clib.build()
    cores = len(os.sched_getaffinity(0)) or 1          # the cores this process may actually use
    size = args.cpu_sample
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for _ in range(args.warmup if args.warmup < 1 else 1):
            pool.map(_cpu_worker, [(256, SEED + i) for i in range(cores)])
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_cpu_worker, [(size, SEED + i) for i in range(cores)])     # one sample tile per core per step
        dt = time.perf_counter() - t0
    cells = size * size * cores * args.steps
    value = cells / dt / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"full conditioning chain, synthetic {TILE}x{TILE} SRTM 1-arcsec tile per GPU "
                               "(BASELINE.json configs[1])",
                   "sample": f"{cores} x {size}x{size} tiles per step (cost is linear in cells)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"oracle/chain.py on {cores} processes x {size}x{size} synthetic tiles per step"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from hydrodem_b200 import _lib, device as dev
    from hydrodem_b200.pipeline import ConditioningChain
    from hydrodem_b200.synth import SynthScene

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the conditioning path has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL writes its banner ("NCCL version ...") and any NCCL_DEBUG output to stdout (NCCL_DEBUG_FILE is ignored at
        # NCCL_DEBUG=VERSION): stdout carries exactly one JSON line, so file descriptor 1 points at stderr while the
        # communicator is created and the first collective runs
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    lib = _lib.load()
    ny = nx = args.size

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # synthetic inputs of this rank's tile, in pinned host memory
    scene = SynthScene(ny, nx, SEED + rank)
    host = {}
    for name, arr in (("srtm", scene.srtm()), ("groves", scene.groves()), ("hsheds", scene.hsheds())):
        pin = dev.pinned_empty(arr.shape, arr.dtype)
        pin[...] = arr
        host[name] = pin
    chain = ConditioningChain()
    d_in = chain.upload_inputs(host["srtm"], host["groves"], host["hsheds"])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")       # > 126 MB L2

    captured = None if args.no_graph else chain.capture(*d_in)             # one CUDA graph per pass

    def step_resident():
        return captured.replay() if captured is not None else chain.run_device(*d_in)

    # ---- resident-input timing ------------------------------------------------------------------------
    sampler = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ else local)
    if rank == 0:
        sampler.start()                    # nvidia-smi needs ~0.5 s to start streaming: launch it before the warm-up
    for _ in range(args.warmup):
        step_resident()
    barrier()
    if rank == 0:
        sampler.mark()                     # only samples taken from here on (the timed region) are reported
    lib.hd_reset_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sweeps = []
    barrier()
    for k in range(args.steps):
        flush.fill_(k & 0xff)                                              # evict L2 between timed steps (untimed)
        ev[k][0].record()
        res = step_resident()
        ev[k][1].record()
        sweeps.append(res.info.get("fill_sweeps"))
    barrier()
    launches = int(lib.hd_launch_count()) if captured is None else captured.launches * args.steps
    # nvidia-smi delivers a sample every 20-100 ms; a timed region of K x 2.5 ms can end before the first one.  If so,
    # the identical steps keep running (untimed) until a few samples under the same load exist, and the line says so.
    extra = 0
    if rank == 0:
        t_end = time.perf_counter() + 1.5
        while sampler.count() < 4 and time.perf_counter() < t_end:
            step_resident()
            torch.cuda.synchronize()
            extra += 1
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        clocks["sampled"] = ("during the timed steps" if extra == 0 else
                             f"during the timed steps and {extra} identical untimed steps run right after them")
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    total_ms = max_over_ranks(total_ms)
    ms_per_step = total_ms / args.steps
    cells = ny * nx
    value = world * cells / (ms_per_step * 1e-3) / 1e6

    stats = ConditioningChain(fill_stats=True).run_device(*d_in)          # untimed: tile visits of the fill worklist
    sweeps = [stats.info.get("fill_sweeps")]
    # ---- per-kernel profile of one more step (event pair around every launch) -----------------------------
    lib.hd_profile_enable(1)
    flush.fill_(1)
    chain.run_device(*d_in)                                               # eager launches: an event pair around each
    import ctypes
    cbuf = ctypes.create_string_buffer(1 << 16)
    lib.hd_profile_report(cbuf, 1 << 16)
    lib.hd_profile_enable(0)
    kernels = {}
    for line in cbuf.value.decode().splitlines():
        name, cnt, ms = line.split()
        kernels[name] = {"launches": int(cnt), "total_ms": float(ms)}
    prof_total = sum(k["total_ms"] for k in kernels.values()) or 1.0
    top = max(kernels, key=lambda n: kernels[n]["total_ms"])
    peak, peak_kind = measured_peak()
    top_avg_ms = kernels[top]["total_ms"] / kernels[top]["launches"]
    algo_bytes = ALGO_BYTES_PER_CELL.get(top, 8.0) * cells
    if top == "fill_async_kernel" and sweeps and sweeps[-1]:
        algo_bytes = 12.0 * 64 * 64 * sweeps[-1]                          # bytes actually staged: tile visits x 64 x 64 cells
    achieved = algo_bytes / (top_avg_ms * 1e-3) / 1e9
    # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full captures (profiles/)
    # dram__bytes_read.sum + dram__bytes_write.sum per launch from the `ncu --set full` captures of this round
    # (profiles/r1_fft_rows_final_raw.csv: four launches 100.3 / 60.8 / 69.3 / 121.8 MB; profiles/r1_fill_async_raw.csv)
    ncu_traffic = {"fft_rows_kernel": 88.0e6, "fill_async_kernel": 130.9e6}
    roofline = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
                "frac": achieved / peak, "traffic": ncu_traffic.get(top), "launches_per_step": kernels[top]["launches"],
                "avg_launch_ms": top_avg_ms, "share_of_step": kernels[top]["total_ms"] / prof_total,
                "algorithmic_bytes_per_launch": algo_bytes}
    breakdown = {n: round(k["total_ms"], 4) for n, k in sorted(kernels.items(), key=lambda kv: -kv[1]["total_ms"])}

    # ---- end to end through the public API, host buffers ---------------------------------------------------
    # every step uploads its three input rasters from pinned host memory and reads its three results back into
    # host arrays; ConditioningChain.stream overlaps the copies of neighbouring steps with the kernels
    e2e_steps = max(1, args.e2e_steps if args.e2e_steps > 0 else args.steps)
    def tiles(n):
        for _ in range(n):
            yield (host["srtm"], host["groves"], host["hsheds"])

    for r in chain.stream(tiles(max(8, args.warmup)), depth=args.stream_depth):
        del r                                                              # pinned result buffers go back to the cache
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for r in chain.stream(tiles(e2e_steps), depth=args.stream_depth):
        outs = (r["final"], r["filled"], r["d8"])                          # host arrays (pinned), copies complete
        del r, outs
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    barrier()
    h2d, d2h = chain.last_transfer_bytes                                   # what crossed PCIe (final goes as float32)
    e2e_ms = max_over_ranks(t_e2e / e2e_steps * 1e3)
    e2e_value = world * cells / (e2e_ms * 1e-3) / 1e6
    # latency of ONE tile through the same API (nothing to overlap with)
    t_single = []
    for k in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = chain.apply_to_host(host["srtm"], host["groves"], host["hsheds"])
        torch.cuda.synchronize()
        t_single.append(time.perf_counter() - t0)
        del r
    single_ms = max_over_ranks(min(t_single[1:]) * 1e3)
    barrier()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"full conditioning chain, synthetic {ny}x{nx} SRTM 1-arcsec tile per GPU "
                                   "(BASELINE.json configs[1])",
                       "stages": "fft2+peak mask+ifft2, groves x3 (quadratic 15), nanfix, majority 11, erode2, expand 7, "
                                 "max 7x7, combine, mean3+round, sink-fill, D8",
                       "tile": [ny, nx], "seed": SEED, "parallelism": f"tile-parallel x{world}, no collective",
                       "l2": "256 MB flush buffer written between timed steps (untimed)",
                       "launch": "eager launches" if captured is None else f"one CUDA graph of {captured.launches} kernels per step",
                       "fill_tile_visits": sweeps[-1] if sweeps else None},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms, "steps": e2e_steps,
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "api": "for out in hydrodem_b200.pipeline.ConditioningChain.stream(tiles): out = {final, filled, d8} "
                           f"ndarrays; {args.stream_depth} slots, copies of neighbouring steps overlap the kernels",
                    "single_tile_latency_ms": single_ms,
                    "host": {"cpus": len(os.sched_getaffinity(0)), "ranks_on_box": env_int("LOCAL_WORLD_SIZE", 1)},
                    "single_tile_api": "ConditioningChain.apply_to_host(srtm, groves, hsheds)"},
            "gpu_launches": launches, "launches_per_step": launches / args.steps,
            "roofline": roofline, "kernel_ms": breakdown, "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            from oracle import clib
            clib.build()
            dt = cpu_chain_sample(args.cpu_sample, SEED)
            line["cpu_baseline"] = {"value": args.cpu_sample ** 2 / dt / 1e6, "unit": UNIT, "cores": 1, "kind": "port",
                                    "seconds": dt,
                                    "sample": f"oracle/chain.py (NumPy/SciPy restatement of the reference filters) on "
                                              f"one {args.cpu_sample}x{args.cpu_sample} synthetic tile, 1 process"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=TILE, help="tile edge (default 3601 = BASELINE.json configs[1])")
    ap.add_argument("--cpu-sample", type=int, default=1024, help="edge of the CPU baseline sample tile")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = same as --steps")
    ap.add_argument("--stream-depth", type=int, default=3, help="tiles in flight in the e2e streaming loop")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="issue the kernels one by one instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
