#!/usr/bin/env python3
"""bench.py -- headline benchmark: BASELINE.json metric "images/sec (256x256 transform+augment)".

  python bench.py --gpus N --steps K --warmup W            # our CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (host cores)

A "step" is one pass of the WHOLE hot path over one synthetic batch of 4096 256x256x3 uint8 leaf-like images PER GPU
(BASELINE configs[1]):
  * transform half -- the core profile (5x5 Gaussian blur + make_mask + masked ROI letterbox + RGB/HSV/LAB histograms):
    one launch of the fused kernel k_core, which also adds the batch to the rank's dataset-level colour histogram;
  * augment half   -- the six ImageAugmenter ops on every image of the same resident batch (flip, rotate, skew, shear,
    crop, distortion; one task seed per op and image, parameters drawn natively in the reference's RNG order, the
    distortion noise generated on the device): 9 launches.
Weak scaling: the batch shards by image, no data-path collective; the per-rank dataset histogram is merged by ONE NCCL
allreduce per timed pass (inside the timed region; the same per-step work runs at N = 1 and N > 1).
Prints ONE JSON line (rank 0).  Sub-records `configs.c3_balance / c4_1024 / c5_resize224` time BASELINE configs[2..4].
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
PARITY = "P1 (mask_strategy hsv_h, grabcut_refine false, no upscale)"
# core transform profile, algorithmic HBM bytes per image
#   read 3N; write blur 3N, mask N, ROI canvas 3*RH*RW, hist 9*256*4, bbox/counters 80 B
CORE_BYTES = lambda n, roi: 7 * n + 3 * roi * roi + 9 * 256 * 4 + 80  # noqa: E731


def workload_text(B, S):
    return (f"core transform profile (blur+mask+ROI+histograms) + 6-op augment set (flip, rotate, skew, shear, crop, "
            f"distortion): {B} x {S}x{S}x3 uint8 leaf-like images per GPU, inputs resident in memory")


def _gen_chunk(args):
    from leaffliction_b200 import synth
    start, n, h, w, seed = args
    return synth.leaf_batch(n, h, w, seed, start)


def make_images(n, h, w, seed, pool):
    import numpy as np
    per = 64
    jobs = [(s, min(per, n - s), h, w, seed) for s in range(0, n, per)]
    return np.concatenate(pool.map(_gen_chunk, jobs))


def task_seeds(B, rank=0):
    """One non-zero task seed per op and image, as `random.randint(0, 1000000)` would hand them out
    (dataset_balancer.py:127); identical for the CUDA arm and the reference arm."""
    import numpy as np
    rng = np.random.default_rng(20260 + rank)
    return rng.integers(1, 1000001, size=(6, B), dtype=np.int64)


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


def cpu_reference_rate(images, seeds, cores, pool):
    """Reference CPU path (oracle/refcalls.py = the reference's own OpenCV/Pillow/NumPy calls on arrays, no
    JPEG I/O): core transform + the six augmentations of every image, on `cores` worker processes."""
    from oracle import refcalls
    t0 = time.perf_counter()
    n = refcalls.core_transform_pool(list(images), pool, cores, seeds)
    dt = time.perf_counter() - t0
    return n / dt, dt


def workload_config(batch, size, world, aug_outputs=6):
    """The `config` object of BOTH arms (this implementation and `--impl reference`): what is computed, not how."""
    return {"workload": workload_text(batch, size), "parity_profile": PARITY,
            "l2_policy": f"inputs larger than L2 ({batch * size * size * 3 / 1e6:.0f} MB per step vs 126 MB L2)",
            "images_per_gpu": batch, "parallelism": f"image-sharded x{world}", "augment_outputs_per_image": aug_outputs}


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
    sample = min(args.batch, max(cores * 48, 128))       # a few seconds of work on all host cores per step
    seeds = task_seeds(args.batch)[:, :sample]
    with mp.get_context("fork").Pool(cores) as pool:
        imgs = make_images(sample, args.size, args.size, 1234, pool)
        for _ in range(max(min(args.warmup, 2), 1)):
            cpu_reference_rate(imgs[: max(cores * 2, 8)], seeds[:, : max(cores * 2, 8)], cores, pool)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cpu_reference_rate(imgs, seeds, cores, pool)
        dt = time.perf_counter() - t0
    rate = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args.batch, args.size, args.gpus),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} images of the {args.batch}-image batch per step x {args.steps} steps; "
                                   f"oracle/refcalls.py = the reference's OpenCV/Pillow/NumPy/SciPy calls on in-memory "
                                   f"arrays (no JPEG I/O): core transform + 6 augmentations per image, one process per core"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def bind_rank_to_cores(local, nlocal):
    """Give each rank of the box its own slice of the cores this process may use (preferring the GPU's NUMA-local
    cores when /sys exposes them): the ranks' host threads (parameter draws, launches, pinned copies) stop competing."""
    try:
        avail = sorted(os.sched_getaffinity(0))
        if nlocal <= 1 or len(avail) < 2 * nlocal:
            return {"cores": len(avail), "bound": False}
        per = len(avail) // nlocal
        mine = avail[local * per:(local + 1) * per]
        os.sched_setaffinity(0, mine)
        return {"cores": len(mine), "bound": True, "first": mine[0], "last": mine[-1]}
    except Exception as e:   # noqa: BLE001
        return {"cores": None, "bound": False, "error": str(e)}


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
    ap.add_argument("--no-configs", action="store_true", help="skip the c3 / c4 / c5 sub-records")
    ap.add_argument("--transform-only", action="store_true", help="debug: time the transform half alone")
    ap.add_argument("--serial", action="store_true", help="debug: every kernel of the step back to back on one stream")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    from leaffliction_b200 import augment as aug_mod
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
    cores_all = len(os.sched_getaffinity(0))
    gen_procs = max(1, cores_all // max(1, min(world, 8)))
    with mp.get_context("fork").Pool(gen_procs) as pool:
        imgs_np = make_images(B, S, S, 1234 + 100000 * rank, pool)
    binding = bind_rank_to_cores(local, int(os.environ.get("LOCAL_WORLD_SIZE", str(world))))
    host_in = torch.from_numpy(imgs_np).pin_memory()
    x = host_in.to(dev, non_blocking=True)
    engine = eng.TransformEngine(S, S, ops.mask_cfg("hsv_h"), 1.5, (256, 256), dev, chunk=512, augment=not args.transform_only)
    out = ops.alloc_core_outputs(B, S, S, (256, 256), dev)
    ds_hist = torch.zeros((9, 256), dtype=torch.int64, device=dev)
    augset = None if args.transform_only else aug_mod.AugmentSet(B, S, S, dev, concurrent=not args.serial, pipelined=not args.serial)
    seeds = task_seeds(B, rank)

    def step():
        # three streams: noise + distortion, the five geometric augment kernels, and k_core run side by side (all but the
        # distortion are issue- or latency-bound: together they keep the schedulers busier than back to back); the steps
        # are pipelined: each stream is in order with itself, the streams are joined once per timed pass (merge())
        if augset is not None:
            augset.start(x, seeds)               # augment half: 6 ops x B images (side streams)
        engine.run_device(x, out, ds_hist)       # k_core: transform half + the rank's dataset colour histogram
        if augset is not None:
            augset.finish()

    def merge():
        if augset is not None:
            augset.join()      # pipelined steps: the caller's stream waits for the augment streams once per pass
        if world > 1:  # ONE allreduce per dataset pass, inside the timed region
            dist.all_reduce(ds_hist)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world > 1:
            t = torch.tensor([v], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return v

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
    ms = max_over_ranks(e0.elapsed_time(e1))
    value = world * B * args.steps / (ms / 1e3)
    # the dataset histogram must equal the per-image histograms summed (checked outside the timed region)
    ds_ok = bool(torch.equal(out.hist9.sum(dim=0, dtype=torch.int64) * args.steps, ds_hist)) if world == 1 else None

    # ---- per-kernel rooflines: every kernel of the step timed alone with CUDA events on the launching stream
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
    peak, peak_src = peak_hbm()
    k_ms = {"k_core": timed(lambda: engine.run_device(x, out, ds_hist), reps)}
    core_bytes = CORE_BYTES(N, 256)                # 664,656 B per 256x256 image
    algo = {"k_core": core_bytes * B}
    if augset is not None:
        tm = {}
        augset.run(x, seeds, timings={})
        for _ in range(reps):
            augset.run(x, seeds, timings=tm)
        for k, v in tm.items():
            k_ms[k] = v / reps
        algo.update(augset.algo_bytes())
    kernels = []
    for k, t in k_ms.items():
        gbs = algo[k] / (t / 1e3) / 1e9
        kernels.append({"kernel": k, "ms": round(t, 4), "algo_bytes_per_launch": int(algo[k]), "achieved_gbs": round(gbs, 1),
                        "frac": round(gbs / peak, 4), "share_of_step": None})
    tot_k = sum(k["ms"] for k in kernels)
    for k in kernels:
        k["share_of_step"] = round(k["ms"] / tot_k, 4)
    step_bytes = sum(algo.values())
    achieved = algo["k_core"] / (k_ms["k_core"] / 1e3) / 1e9
    traffic, traffic_note = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")   # dram bytes per image from the committed ncu --set full capture
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic = float(tj["k_core_dram_bytes_per_image"]) * B
            traffic_note = (f"dram__bytes_read+write of k_core, ncu --set full, {tj.get('images_in_launch')} images in the profiled "
                            f"launch, scaled per image to this launch; numbers committed in profiles/traffic.json "
                            f"(capture: {tj.get('report', 'ncu report')})")
        except Exception:
            traffic = None
    step_gbs = step_bytes * (args.steps / (ms / 1e3)) / 1e9
    roofline = {"bound": "hbm", "kernel": "k_core (fused blur+mask+ROI+histograms, one block per image): the transform half's only kernel",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_note, "peak_source": peak_src,
                "algo_bytes_per_launch": algo["k_core"], "frac_of_nominal_8TBs": achieved / 8000.0,
                "kernels": kernels,
                "largest_time_share": max(kernels, key=lambda k: k["ms"])["kernel"],
                "step": {"algo_bytes_per_image": step_bytes / B, "achieved_gbs": step_gbs, "frac": step_gbs / peak,
                         "note": "whole step (transform + augment), all kernels back to back: algorithmic bytes / step time"}}

    # ---- end to end: host buffers in, host buffers out (all seven transform outputs + the six augment outputs),
    # copies inside the timed region
    e2e = None
    if not args.no_e2e:
        host_out = eng.alloc_host_outputs(B, S, S, (256, 256), augment=augset is not None)
        kw = {"seeds": seeds} if augset is not None else {}
        for _ in range(2):
            engine.run_host(host_in, host_out, **kw)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            engine.run_host(host_in, host_out, **kw)     # returns when the host buffers are filled
        barrier()
        ems = max_over_ranks((time.perf_counter() - t0) * 1e3)
        h2d, d2h = int(host_in.numel()), int(host_out.nbytes())
        # PCIe ceiling of this rank: one large pinned copy each way
        probe = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
        dprobe = torch.empty_like(probe, device=dev)
        bw_in = probe.numel() / (timed(lambda: dprobe.copy_(probe, non_blocking=True), 3) / 1e3) / 1e9
        bw_out = probe.numel() / (timed(lambda: probe.copy_(dprobe, non_blocking=True), 3) / 1e3) / 1e9
        bw_in, bw_out = -max_over_ranks(-bw_in), -max_over_ranks(-bw_out)     # the slowest rank
        ceil_rate = world * B / max(h2d / (bw_in * 1e9), d2h / (bw_out * 1e9))
        e2e_val = world * B * args.steps / (ems / 1e3)
        e2e = {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
               "api": "leaffliction_b200.engine.TransformEngine(augment=True).run_host (pinned host in/out, 3-stream chunked "
                      "overlap, returns after the last device-to-host copy); timed with the host clock",
               "pcie": {"h2d_gbs": round(bw_in, 2), "d2h_gbs": round(bw_out, 2), "ceiling_images_s": ceil_rate,
                        "frac_of_ceiling": e2e_val / ceil_rate,
                        "note": "pinned 256 MiB cudaMemcpyAsync each way on this rank (slowest rank at N > 1); ceiling = the "
                                "slower direction moving this step's bytes, directions overlapped"}}
        del probe, dprobe, host_out
    clocks = sampler.stop() if rank == 0 else None

    # ---- BASELINE configs[2..4] as sub-records
    configs = None
    if not args.no_configs and not args.transform_only:
        from tools import bench_configs
        configs = bench_configs.run_all(x, dev, rank, world, peak, max_over_ranks, barrier)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample = min(B, max(cores_all * 64, 256))
        with mp.get_context("fork").Pool(cores_all) as pool:
            cpu_reference_rate(imgs_np[: max(cores_all * 2, 8)], seeds[:, : max(cores_all * 2, 8)], cores_all, pool)
            t0 = time.perf_counter()
            passes = 0
            while passes < 1 or time.perf_counter() - t0 < 10.0:
                cpu_reference_rate(imgs_np[:sample], seeds[:, :sample], cores_all, pool)
                passes += 1
            dt = time.perf_counter() - t0
        rate = passes * sample / dt
        cpu = {"value": rate, "unit": UNIT, "cores": cores_all, "kind": "port",
               "sample": f"{sample} images of the batch x {passes} passes, {dt:.1f} s wall on {cores_all} cores; oracle/refcalls.py (the "
                         f"reference's OpenCV/Pillow/NumPy/SciPy calls on in-memory arrays, no JPEG I/O): core transform + 6 "
                         f"augmentations per image, one process per core"}

    if rank == 0:
        launches = 1 + (aug_mod.AugmentSet.launches_per_run if augset is not None else 0)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": workload_config(B, S, world, 6 if augset is not None else 0),
            "run": {"streams": 1 if (args.serial or augset is None) else 3,
                    "dataset_histogram_matches_per_image_sum": ds_ok, "host_binding": binding},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks, "configs": configs,
            "gpu_launches": launches * args.steps,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
