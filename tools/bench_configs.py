"""BASELINE.json configs[2..4] as bounded sub-benchmarks, called by bench.py (sub-records `configs.*` of its JSON line)
and runnable alone.  Device-timed with CUDA events, max over ranks, inputs resident in HBM.

  c3_balance    class-balancing augmentation of a 64 Ki-image, 8-class imbalanced dataset (36,864 augment tasks),
                tasks sharded by index across the ranks (strong scaling), class histogram merged by one allreduce
  c4_1024       1024x1024x3 leaves: the augment warps (skew, shear, rotate), the 5x5 / 15x15 blur and the core transform
                profile, 256 images per GPU (weak scaling)
  c5_resize224  augment -> Lanczos 224x224 -> /255 float32 -> DLPack (train.py's leaf_cnn input), batch 32..1024 on one GPU
  f2_jpeg       JPEG bitstreams in -> nvJPEG decode -> core transform -> nvJPEG encode -> bitstreams out (one GPU)
  f3_overlays   the Analyze overlay and the ROI rectangle image of every image of a resident batch (one GPU)
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from leaffliction_b200 import augment, balance, ops, synth  # noqa: E402


def _timed(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def c3_balance(x, dev, rank, world, peak, max_over_ranks, barrier, images=65536):
    S = int(x.shape[1])
    counts = balance.synthetic_class_counts()
    scale = images / 65536.0
    names = [c for p in counts.values() for c in p]
    plants = {p: list(c) for p, c in counts.items()}
    per_class = [max(1, int(round(n * scale))) for p in counts.values() for n in p.values()]
    labels = np.repeat(np.arange(len(names)), per_class)
    N = len(labels)
    base = x[:128]
    data = base.repeat((N + 127) // 128, 1, 1, 1)[:N].contiguous()      # the dataset, resident in HBM
    part = np.bincount(labels[list(balance.shard(N, rank, world))], minlength=len(names)).astype(np.int64)
    merged, _ = balance.allreduce_histograms(part, device=dev)           # warm-up of the collective (communicator init)
    t0 = time.perf_counter()
    merged, _ = balance.allreduce_histograms(part, device=dev)           # the ONE collective of the pass
    assert merged.tolist() == per_class
    plan, all_tasks = balance.task_arrays_for_labels(labels, names, plants, seed=42)
    t_plan = time.perf_counter() - t0
    mine = all_tasks.shard(rank, world)
    augment.augment_device(data, mine.slice(0, None, max(1, len(mine) // 256)))      # warm-up: tables, allocator pools
    augment.augment_device(data, mine)
    barrier()
    passes = []
    for _ in range(2):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        res = augment.augment_device(data, mine)
        b.record()
        torch.cuda.synchronize()
        passes.append((a.elapsed_time(b), (time.perf_counter() - t0) * 1e3))
        del res
    dev_ms = max_over_ranks(min(p[0] for p in passes))
    wall_ms = max_over_ranks(min(p[1] for p in passes))
    n_tasks = len(all_tasks)
    del data
    torch.cuda.empty_cache()
    return {"workload": f"class balancing: {N} images {S}x{S}x3 in HBM, 8 classes (imbalanced), {n_tasks} augment tasks "
                        f"(6 ops, SURVEY 8d counts), tasks sharded by index over {world} GPU(s), one histogram allreduce",
            "scaling": "strong", "n_gpus": world, "tasks": n_tasks, "value": n_tasks / (wall_ms / 1e3), "unit": "augmented images/s",
            "ms_per_pass_wall": wall_ms, "ms_per_pass_device": dev_ms, "plan_histogram_tasklist_ms": t_plan * 1e3,
            "algo_bytes_per_task_mean": 2.41e6 / 6, "achieved_gbs": n_tasks * (2.41e6 / 6) / (wall_ms / 1e3) / 1e9,
            "frac": n_tasks * (2.41e6 / 6) / (wall_ms / 1e3) / 1e9 / peak / world}


def c4_1024(dev, rank, world, peak, max_over_ranks, barrier, B=256, S=1024):
    base = synth.leaf_batch(4, S, S, 4321 + rank)
    x = torch.from_numpy(base).to(dev).repeat(B // 4, 1, 1, 1).contiguous()
    N = S * S
    rng = np.random.default_rng(7 + rank)
    res = {}

    def add(name, fn, algo_bytes, reps=3):
        ms = max_over_ranks(_timed(fn, reps))
        gbs = algo_bytes * B / (ms / 1e3) / 1e9
        res[name] = {"ms": round(ms, 4), "algo_bytes_per_image": int(algo_bytes), "achieved_gbs": round(gbs, 1), "frac": round(gbs / peak, 4)}
        return ms

    out = torch.empty_like(x)
    sk = np.array([[1 + s, 0, -s * S, 0, 1 + s, -s * S, 0, 0] for s in rng.uniform(0.05, 0.15, B)])
    dsk, pe = torch.from_numpy(sk).to(dev), torch.ones(B, dtype=torch.int32, device=dev)
    t_skew = add("warp_bicubic_skew", lambda: ops.warp_bicubic(x, dsk, pe, out=out), 6 * N)
    sh = np.array([[1, k, 0, 0, 1, 0, 0, 0] if h else [1, 0, 0, k, 1, 0, 0, 0] for k, h in zip(rng.uniform(-0.2, 0.2, B), rng.random(B) < 0.5)], np.float64)
    dsh, pa = torch.from_numpy(sh).to(dev), torch.zeros(B, dtype=torch.int32, device=dev)
    t_shear = add("warp_bicubic_shear", lambda: ops.warp_bicubic(x, dsh, pa, out=out), 6 * N)
    params = np.zeros((B, 8), np.int32)
    px = 0
    for i, a in enumerate(rng.uniform(-30, 30, B)):
        m, nw, nh = augment.rotate_matrix(float(a), S, S)
        params[i, :6], params[i, 6:] = augment.fixed_affine(m), (nw, nh)
        px += nw * nh
    slab, _ = ops.rotate_nn(x, params)
    dpar = torch.from_numpy(params).to(dev)
    t_rot = add("rotate_nn", lambda: ops.rotate_nn(x, params, 255, dpar, slab), 3 * N + 3 * px / B)
    del slab
    t_g5 = add("gauss_u8_5x5", lambda: ops.gauss_u8(x, 5, 1.5), 6 * N)
    add("gauss_u8_15x15", lambda: ops.gauss_u8(x, 15, 0.0), 6 * N)
    cfg = ops.mask_cfg("hsv_h")
    co = ops.alloc_core_outputs(B, S, S, (S, S), dev)
    add("pipeline_core", lambda: ops.pipeline_core(x, cfg, 1.5, (S, S), co), 7 * N + 3 * N + 9 * 256 * 4 + 80)
    tot = t_skew + t_shear + t_rot + t_g5
    by = (6 + 6 + 6) * N + 3 * N + 3 * px / B
    del x, out, co
    torch.cuda.empty_cache()
    return {"workload": f"{B} x {S}x{S}x3 images per GPU in HBM: augment warps (skew, shear, rotate) + 5x5 blur, then 15x15 blur and the core profile",
            "scaling": "weak", "n_gpus": world, "value": world * B / (tot / 1e3), "unit": "images/s (skew+shear+rotate+blur5 per image)",
            "achieved_gbs": by * B / (tot / 1e3) / 1e9, "frac": by * B / (tot / 1e3) / 1e9 / peak, "ops": res}


def c5_resize224(x, dev, peak, chunk=512):
    """augment -> Lanczos 224x224 -> /255 float32 -> DLPack.  Pillow rounds the augmented image to uint8 before the resize
    (the reference writes the augmented JPEG, train.py reads it back: sequence.py:84-88), so the two stages stay two
    kernels; they run chunk by chunk (512 images = 100 MB of uint8 intermediate), so that the resize reads the
    intermediate from the 126 MB L2 and HBM sees one read of the source and one write of the float32 tensor."""
    S = int(x.shape[1])
    N = S * S
    by = 3 * N + 224 * 224 * 3 * 4                      # 798,720 B per image (read u8, write f32)
    rng = np.random.default_rng(5)
    out, per_op = {}, {}
    bmax = min(1024, int(x.shape[0]))
    mode = torch.zeros(bmax, dtype=torch.int32, device=dev)
    sk = torch.from_numpy(np.array([[1 + s, 0, -s * S, 0, 1 + s, -s * S, 0, 0] for s in rng.uniform(0.05, 0.15, bmax)])).to(dev)
    pe = torch.ones(bmax, dtype=torch.int32, device=dev)
    boxes = np.zeros((bmax, 4), np.int32)
    for i, r in enumerate(rng.uniform(0.8, 0.95, bmax)):
        nw = int(S * r)
        boxes[i] = (rng.integers(0, S - nw + 1), rng.integers(0, S - nw + 1), nw, nw)
    tmp = torch.empty((chunk, S, S, 3), dtype=torch.uint8, device=dev)
    o8 = torch.empty((bmax, 224, 224, 3), dtype=torch.uint8, device=dev)
    of = torch.empty((bmax, 224, 224, 3), dtype=torch.float32, device=dev)
    plans224, crop_plans = {}, {}

    def plan224(n):                 # parameter tables are built once per batch size, outside the timed region
        if n not in plans224:
            plans224[n] = ops.CropPlan(np.tile(np.array([0, 0, S, S], np.int32), (n, 1)), (224, 224), dev)
        return plans224[n]

    def crop_plan(c0, n):
        if (c0, n) not in crop_plans:
            crop_plans[(c0, n)] = ops.CropPlan(boxes[c0:c0 + n], (S, S), dev)
        return crop_plans[(c0, n)]

    def pipeline(op, b):
        for c0 in range(0, b, chunk):
            n = min(chunk, b - c0)
            xs, t = x[c0:c0 + n], tmp[:n]
            if op == "flip":
                ops.flip(xs, mode[c0:c0 + n], out=t)
            elif op == "skew":
                ops.warp_bicubic(xs, sk[c0:c0 + n], pe[c0:c0 + n], out=t)
            else:
                ops.crop_lanczos(xs, crop_plan(c0, n), out=t)
            ops.crop_lanczos(t, plan224(n), want_f32=True, out=o8[c0:c0 + n], outf=of[c0:c0 + n])
        return torch.utils.dlpack.to_dlpack(of[:b])     # zero-copy hand-over to the training framework

    for b in (32, 128, 512, 1024):
        if b > bmax:
            continue
        ms = _timed(lambda: pipeline("flip", b), reps=10, warm=2)
        out[str(b)] = {"ms": round(ms, 4), "images_per_s": round(b / (ms / 1e3)), "achieved_gbs": round(by * b / (ms / 1e3) / 1e9, 1),
                       "frac": round(by * b / (ms / 1e3) / 1e9 / peak, 4)}
    for op in ("flip", "skew", "crop"):
        ms = _timed(lambda: pipeline(op, bmax), reps=5, warm=2)
        per_op[op] = {"ms": round(ms, 4), "images_per_s": round(bmax / (ms / 1e3)), "frac": round(by * bmax / (ms / 1e3) / 1e9 / peak, 4)}
    best = max(out.values(), key=lambda r: r["images_per_s"])
    return {"workload": "augment (flip; also skew, crop) -> Lanczos 224x224 -> /255 float32 -> DLPack, 256x256x3 inputs in HBM, "
                        f"chunks of {chunk} images so that the uint8 intermediate stays in L2; batch sweep on one GPU",
            "n_gpus": 1, "value": best["images_per_s"], "unit": "images/s (flip, best batch)", "algo_bytes_per_image": by,
            "batches": out, f"ops_at_batch_{bmax}": per_op}


def p0_default_strategy(x, dev, rank, world, peak, max_over_ranks, barrier, n=2048):
    """The core transform profile with the reference YAML's default mask strategy (config.yaml:7 `inclusive`; parity profile
    P0: grabCut off, no upscale): front-end kernel + make_mask on its candidate + blur + ROI + statistics."""
    from leaffliction_b200 import engine as eng
    xs = x[:n]
    S = int(x.shape[1])
    e = eng.TransformEngine(S, S, ops.mask_cfg("hsv_h"), 1.5, (256, 256), dev, front="inclusive")
    out = ops.alloc_core_outputs(len(xs), S, S, (256, 256), dev)
    ms = max_over_ranks(_timed(lambda: e.run_device(xs, out), reps=3, warm=2))
    by = 7 * S * S + 3 * 256 * 256 + 9 * 256 * 4 + 80
    return {"workload": f"core transform profile, mask_strategy inclusive (the reference's default), {len(xs)} x {S}x{S}x3 images per GPU in HBM",
            "parity_profile": "P0 (config.yaml defaults, grabcut_refine false, no upscale)", "scaling": "weak", "n_gpus": world,
            "value": world * len(xs) / (ms / 1e3), "unit": "images/s", "ms": round(ms, 4), "launches": 2,
            "achieved_gbs": by * len(xs) / (ms / 1e3) / 1e9, "frac": by * len(xs) / (ms / 1e3) / 1e9 / peak}


def f2_jpeg(x, dev, n=1024):
    """SURVEY 8f rank 2, the file boundary: JPEG bitstreams in host memory -> nvJPEG decode into the device batch -> core
    transform profile -> nvJPEG encode of the blur and ROI outputs -> bitstreams in host memory (what folder mode does per
    image with Pillow / cv2.imwrite).  Next to it, the codec work alone through Pillow on every host core."""
    import io
    import os
    import time
    from concurrent.futures import ThreadPoolExecutor

    from PIL import Image

    from leaffliction_b200 import engine as eng
    from leaffliction_b200 import jpegio
    S = int(x.shape[1])
    imgs = x[:n].cpu().numpy()
    cores = len(os.sched_getaffinity(0))

    def enc(a):
        b = io.BytesIO()
        Image.fromarray(a).save(b, format="JPEG", quality=95)
        return b.getvalue()

    def dec(b):
        return np.asarray(Image.open(io.BytesIO(b)).convert("RGB"))
    with ThreadPoolExecutor(max_workers=cores) as ex:
        blobs = list(ex.map(enc, imgs))
        t0 = time.perf_counter()
        arrs = list(ex.map(dec, blobs))
        list(ex.map(enc, arrs))
        list(ex.map(enc, arrs))              # two outputs per image, as below
        t_host = time.perf_counter() - t0
    e = eng.TransformEngine(S, S, ops.mask_cfg("hsv_h"), 1.5, (256, 256), dev)
    out = ops.alloc_core_outputs(n, S, S, (256, 256), dev)
    xin = torch.empty((n, S, S, 3), dtype=torch.uint8, device=dev)

    def step():
        jpegio.decode_batch(blobs, S, S, out=xin)
        e.run_device(xin, out)
        a = jpegio.encode_batch(out.blur)
        b = jpegio.encode_batch(out.roi)
        return len(a) + len(b)
    step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    return {"workload": f"{n} JPEG files ({S}x{S}, q95) in host memory -> nvJPEG decode -> core transform profile -> nvJPEG encode of "
                        "blur + ROI -> bitstreams in host memory", "n_gpus": 1, "value": n / dt, "unit": "images/s", "ms": round(dt * 1e3, 2),
            "bytes_over_pcie_per_image": int(np.mean([len(b) for b in blobs])) * 3, "raw_bytes_per_image": 3 * S * S * 3,
            "host_codec_only": {"value": n / t_host, "unit": "images/s", "cores": cores,
                                "what": "Pillow decode + two Pillow encodes per image on all host cores, no transform"}}


def f3_overlays(x, dev, peak, n=2048):
    """the two overlay images of the reference's folder run (apply_analyze_filter's image, analyze.py:37-122;
    the rectangle view of apply_roi_filter, roi.py:43-44) for a resident batch, from the outputs of the core transform: masked
    image, contour trace + record, grey + Canny, the overlay kernel, the rectangle kernel (bit-identical to the OpenCV calls)."""
    from leaffliction_b200 import engine as eng
    xs = x[:n]
    S = int(x.shape[1])
    N = S * S
    e = eng.TransformEngine(S, S, ops.mask_cfg("hsv_h"), 1.5, (256, 256), dev)
    out = e.run_device(xs)
    ms = _timed(lambda: e.overlays_device(xs, out), reps=3, warm=2)
    masked = ops.apply_mask(xs, out.mask, 255)
    rec = ops.analyze_records(out.mask, out.info, 4096, 512)
    gray = ops.cvt_color(masked, "gray")
    edges = ops.canny(gray, 80, 160, True)
    stages = {"apply_mask": (lambda: ops.apply_mask(xs, out.mask, 255), 7 * N),
              "trace_contour+analyze_record": (lambda: ops.analyze_records(out.mask, out.info, 4096, 512), N),
              "grey": (lambda: ops.cvt_color(masked, "gray"), 4 * N),
              "canny": (lambda: ops.canny(gray, 80, 160, True), 2 * N),
              "analyze_overlay": (lambda: ops.analyze_overlay(masked, rec, edges, out.mask), 8 * N),
              "draw_rectangles": (lambda: ops.draw_rectangles(masked, out.info), 6 * N)}
    per = {}
    for k, (fn, by) in stages.items():
        t = _timed(fn, reps=3, warm=1)
        per[k] = {"ms": round(t, 4), "algo_bytes_per_image": by, "frac": round(by * len(xs) / (t / 1e3) / 1e9 / peak, 4)}
    by_all = sum(v[1] for v in stages.values())
    return {"workload": f"Analyze overlay + ROI rectangle image for {len(xs)} x {S}x{S}x3 images in HBM, from the core transform's mask / box "
                        "(OpenCV's drawing restated on the device, bit-identical)", "n_gpus": 1, "value": len(xs) / (ms / 1e3),
            "unit": "images/s (both overlay images per image)", "ms": round(ms, 4), "launches": 7, "algo_bytes_per_image": by_all,
            "achieved_gbs": by_all * len(xs) / (ms / 1e3) / 1e9, "frac": by_all * len(xs) / (ms / 1e3) / 1e9 / peak, "stages": per}


def run_all(x, dev, rank, world, peak, max_over_ranks, barrier):
    res = {}
    if world == 1:
        try:
            res["f2_jpeg"] = f2_jpeg(x, dev)
        except Exception as e:   # noqa: BLE001
            res["f2_jpeg"] = {"error": f"{type(e).__name__}: {e}"}
        try:
            res["f3_overlays"] = f3_overlays(x, dev, peak)
        except Exception as e:   # noqa: BLE001
            res["f3_overlays"] = {"error": f"{type(e).__name__}: {e}"}
    for name, fn in (("p0_default_strategy", lambda: p0_default_strategy(x, dev, rank, world, peak, max_over_ranks, barrier)),
                     ("c3_balance", lambda: c3_balance(x, dev, rank, world, peak, max_over_ranks, barrier)),
                     ("c4_1024", lambda: c4_1024(dev, rank, world, peak, max_over_ranks, barrier)),
                     ("c5_resize224", lambda: c5_resize224(x, dev, peak))):
        try:
            res[name] = fn()
        except Exception as e:   # noqa: BLE001 -- a sub-record must not take the headline line down with it
            res[name] = {"error": f"{type(e).__name__}: {e}"}
            if world > 1:
                raise          # ranks would fall out of step on the collectives: fail loudly instead
    return res
