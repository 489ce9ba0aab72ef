#!/usr/bin/env python3
"""bench.py -- headline benchmark of the core transform profile (BASELINE.json configs[1]).

  python bench.py --gpus N --steps K --warmup W            # our CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (host cores)

A "step" is one pass of the hot path (5x5 Gaussian blur + make_mask + masked ROI letterbox +
RGB/HSV/LAB histograms) over one synthetic batch of 4096 256x256x3 uint8 leaf-like images PER GPU
(weak scaling: the batch shards by image, no data-path collective; with N > 1 the dataset-level colour
histogram is accumulated on each device and merged by ONE NCCL allreduce per pass, inside the timed region).  Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "images/sec (256x256 transform+augment)"
UNIT = "images/s"
# core transform profile, algorithmic HBM bytes per 256x256 image
#   read 3N; write blur 3N, mask N, ROI canvas 3*256*256, hist 9*256*4, bbox/counters 80 B
ALGO_BYTES_PER_IMAGE = lambda n, roi: 7 * n + 3 * roi * roi + 9 * 256 * 4 + 80  # noqa: E731


def _gen_chunk(args):
    from leaffliction_b200 import synth
    start, n, h, w, seed = args
    return synth.leaf_batch(n, h, w, seed, start)


def make_images(n, h, w, seed, pool):
    import numpy as np
    per = 64
    jobs = [(s, min(per, n - s), h, w, seed) for s in range(0, n, per)]
    return np.concatenate(pool.map(_gen_chunk, jobs))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_rate(images, cores, pool):
    """Reference CPU path (oracle/refcalls.py = the reference's own OpenCV/NumPy calls on arrays, no
    JPEG I/O) on `cores` worker processes.  Returns (images/s, seconds)."""
    from oracle import refcalls
    t0 = time.perf_counter()
    n = refcalls.core_transform_pool(list(images), pool, cores)
    dt = time.perf_counter() - t0
    return n / dt, dt


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = len(os.sched_getaffinity(0))
    sample = min(args.batch, max(cores * 128, 256))      # ~1-2 s of work on all host cores per step
    with mp.get_context("fork").Pool(cores) as pool:
        imgs = make_images(sample, args.size, args.size, 1234, pool)
        for _ in range(max(args.warmup, 1)):
            cpu_reference_rate(imgs[: max(cores * 4, 8)], cores, pool)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cpu_reference_rate(imgs, cores, pool)
        dt = time.perf_counter() - t0
    rate = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"core transform profile (blur+mask+ROI+histograms), {args.size}x{args.size}x3 uint8, "
                               f"bounded sample of {sample} images per step of the {args.batch}-image batch",
                   "parity_profile": "P1 (mask_strategy hsv_h, grabcut_refine false, no upscale)"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} images x {args.steps} steps; oracle/refcalls.py = the reference's OpenCV/"
                                   f"NumPy/SciPy calls on in-memory arrays (no JPEG I/O), one process per core"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="images per GPU per step")
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    from leaffliction_b200 import engine as eng
    from leaffliction_b200 import ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, S = args.batch, args.size
    N = S * S
    cores = len(os.sched_getaffinity(0))
    gen_procs = max(1, cores // max(1, min(world, 8)))
    with mp.get_context("fork").Pool(gen_procs) as pool:
        imgs_np = make_images(B, S, S, 1234 + 100000 * rank, pool)
    host_in = torch.from_numpy(imgs_np).pin_memory()
    x = host_in.to(dev, non_blocking=True)
    engine = eng.TransformEngine(S, S, ops.mask_cfg("hsv_h"), 1.5, (256, 256), dev, chunk=512)
    out = ops.alloc_core_outputs(B, S, S, (256, 256), dev)
    ds_hist = torch.zeros((9 * 256,), dtype=torch.int64, device=dev)

    def step():
        engine.run_device(x, out)
        if world > 1:  # dataset-level colour histogram: accumulated on the device batch by batch ...
            ds_hist.add_(out.hist9.sum(dim=0, dtype=torch.int64).view(-1))

    def merge():
        if world > 1:  # ... and merged across ranks by ONE allreduce per dataset pass, inside the timed region
            dist.all_reduce(ds_hist)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    merge()
    ds_hist.zero_()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    merge()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * B * args.steps / (ms / 1e3)

    # ---- roofline of the dominant kernel.  The whole step is ONE launch of the fused kernel k_core (plus a
    # 256-byte memset node that resets its image queue): time it alone with CUDA events on the launching stream.
    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    reps = max(3, min(args.steps, 10))
    k_ms = {"k_core": timed(lambda: engine.run_device(x, out), reps)}
    algo_bytes = ALGO_BYTES_PER_IMAGE(N, 256)   # 664,656 B per 256x256 image
    peak, peak_src = peak_hbm()
    achieved = algo_bytes * B / (k_ms["k_core"] / 1e3) / 1e9
    traffic, traffic_note = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")   # dram bytes per image from the committed ncu --set full capture
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic = float(tj["k_core_dram_bytes_per_image"]) * B
            traffic_note = (f"dram__bytes_read+write of k_core from profiles/{tj.get('report', 'ncu report')} "
                            f"({tj.get('images_in_launch')} images in the profiled launch), scaled per image to this launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "k_core (fused blur+mask+ROI+histograms, one block per image)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_note, "peak_source": peak_src,
                "kernel_ms": {k: round(v, 4) for k, v in k_ms.items()},
                "algo_bytes_per_launch": algo_bytes * B,
                "pipeline_algo_bytes_per_image": algo_bytes,
                "pipeline_achieved_gbs": algo_bytes * (B * args.steps / (ms / 1e3)) / 1e9,
                "pipeline_frac_of_peak": algo_bytes * (B * args.steps / (ms / 1e3)) / 1e9 / peak,
                "frac_of_nominal_8TBs": achieved / 8000.0}

    # ---- end to end: host buffers in, host buffers out, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        host_out = eng.alloc_host_outputs(B, S, S, (256, 256))
        for _ in range(2):
            engine.run_host(host_in, host_out)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            engine.run_host(host_in, host_out)
        b.record()
        barrier()
        ems = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([ems], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ems = float(t.item())
        e2e = {"value": world * B * args.steps / (ems / 1e3), "unit": UNIT,
               "h2d_bytes_per_step": int(host_in.numel()) * world, "d2h_bytes_per_step": int(host_out.nbytes()) * world,
               "api": "leaffliction_b200.engine.TransformEngine.run_host (pinned host in/out, 3-stream chunked overlap)"}
    clocks = sampler.stop() if rank == 0 else None

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        passes = 4                                        # the whole batch, 4 passes: 10-20 s of work on all host cores
        with mp.get_context("fork").Pool(cores) as pool:
            cpu_reference_rate(imgs_np[: max(cores * 2, 8)], cores, pool)
            t0 = time.perf_counter()
            for _ in range(passes):
                cpu_reference_rate(imgs_np, cores, pool)
            dt = time.perf_counter() - t0
        rate = passes * B / dt
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"the whole {B}-image batch x {passes} passes, {dt:.1f} s wall on {cores} cores; oracle/refcalls.py (the "
                         f"reference's OpenCV/NumPy/SciPy calls on in-memory arrays, no JPEG I/O), one process per core"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"core transform profile (blur+mask+ROI+histograms): {B} x {S}x{S}x3 uint8 leaf-like "
                                   f"images per GPU, resident in HBM",
                       "parity_profile": "P1 (mask_strategy hsv_h, grabcut_refine false, no upscale)",
                       "l2_policy": f"inputs larger than L2 ({B * N * 3 / 1e6:.0f} MB per step vs 126 MB L2)",
                       "images_per_gpu": B, "parallelism": f"image-sharded x{world}"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
            "gpu_launches": 1 * args.steps,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
