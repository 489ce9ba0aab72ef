#!/usr/bin/env python3
"""Per-kernel device timing of every C-ABI op on the hot path (NOT the contract benchmark -- that is
bench.py): algorithmic HBM bytes per image / CUDA-event time, as a fraction of the
measured copy peak.  Inputs larger than L2, 3 warm-ups, events on the launching stream.

    python tools/bench_ops.py [--batch 4096] [--size 256] [--reps 5] > profiles/rNN_ops.json
"""
import argparse
import json
import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from leaffliction_b200 import augment, ops, synth  # noqa: E402


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    B, S = args.batch, args.size
    N = S * S
    dev = torch.device("cuda:0")
    base = synth.leaf_batch(min(B, 64), S, S)
    x = torch.from_numpy(np.concatenate([base] * ((B + len(base) - 1) // len(base)))[:B]).to(dev)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    cfg = ops.mask_cfg("hsv_h")
    rng = random.Random(1)
    res = {}

    def add(name, fn, algo_bytes):
        ms = timed(fn, args.reps)
        gbs = algo_bytes * B / (ms / 1e3) / 1e9
        res[name] = {"ms": round(ms, 4), "algo_bytes_per_image": int(algo_bytes), "GB/s": round(gbs, 1), "frac_of_measured_peak": round(gbs / peak, 4),
                     "images_per_s": round(B / (ms / 1e3))}

    # ---- colour / mask / filters
    add("cvt_color_hsv", lambda: ops.cvt_color(x, "hsv"), 6 * N)
    add("cvt_color_lab", lambda: ops.cvt_color(x, "lab"), 6 * N)
    add("cvt_color_gray", lambda: ops.cvt_color(x, "gray"), 4 * N)
    add("threshold_mask_hsv_h", lambda: ops.threshold_mask(x, cfg), 4 * N)
    mask, info = ops.make_mask(x, cfg)
    add("apply_mask", lambda: ops.apply_mask(x, mask, 255), 7 * N)
    add("gauss_u8_5x5", lambda: ops.gauss_u8(x, 5, 1.5), 6 * N)
    add("gauss_u8_15x15", lambda: ops.gauss_u8(x, 15, 0.0), 6 * N)
    add("make_mask_hsv_h", lambda: ops.make_mask(x, cfg), 4 * N + 32)
    add("roi_letterbox", lambda: ops.roi_letterbox(x, mask, info, (S, S)), 4 * N + 3 * N)
    add("color_stats", lambda: ops.color_stats(x, mask), 4 * N + 12 * 256 * 4)
    out = ops.alloc_core_outputs(B, S, S, (S, S), dev)
    add("pipeline_core(fused k_core)", lambda: ops.pipeline_core(x, cfg, 1.5, (S, S), out), 7 * N + 3 * N + 9 * 256 * 4 + 80)
    if N <= 65536:
        nsub = min(B, 1024)
        xs, ms_ = x[:nsub], mask[:nsub]
        gray = ops.cvt_color(xs, "gray")
        t = timed(lambda: ops.canny(gray, 50, 150, True), args.reps)
        res["canny(1024 img)"] = {"ms": round(t, 4), "images_per_s": round(nsub / (t / 1e3))}
        t = timed(lambda: ops.raw_mask_front_end(xs, "inclusive", cfg), args.reps)
        res["raw_mask_inclusive(1024 img)"] = {"ms": round(t, 4), "images_per_s": round(nsub / (t / 1e3))}
        t = timed(lambda: ops.saliency_blur(xs, ms_, cfg, 1.5), args.reps)
        res["saliency_blur(1024 img)"] = {"ms": round(t, 4), "images_per_s": round(nsub / (t / 1e3))}
        t = timed(lambda: ops.brown_spots(xs, ms_, cfg), args.reps)
        res["brown_spots(1024 img)"] = {"ms": round(t, 4), "images_per_s": round(nsub / (t / 1e3))}
        if max(S, S) == 256:   # k-means candidate (cv2.kmeans restated): the image stays in shared memory for every pass
            t = timed(lambda: ops.kmeans_raw(xs), args.reps)
            res["kmeans_raw(1024 img)"] = {"ms": round(t, 4), "images_per_s": round(nsub / (t / 1e3))}
    dsd = torch.arange(1, 6 * B + 1, dtype=torch.int32, device=dev)
    t = timed(lambda: ops.seed_words(dsd, 16), args.reps)
    res[f"seed_words({6 * B} task seeds)"] = {"ms": round(t, 4), "seedings_per_s": round(6 * B / (t / 1e3))}
    # ---- overlay images: contour trace + record, then the OpenCV rasterisers restated
    rec = ops.analyze_records(mask, info, max_pts=8192, max_hull=1000)
    if int(rec["counts"].min()) >= 0 and int(rec["rec_i"][:, 12].min()) >= 0:
        add("analyze_records (trace + record)", lambda: ops.analyze_records(mask, info, max_pts=8192, max_hull=1000), N)
        edges0 = torch.zeros_like(mask)
        add("analyze_overlay", lambda: ops.analyze_overlay(x, rec, edges0, mask), 8 * N)
        add("draw_rectangles", lambda: ops.draw_rectangles(x, info), 6 * N)

    # ---- augmentations (parameters drawn like the reference)
    dmode = torch.zeros(B, dtype=torch.int32, device=dev)
    fout = torch.empty_like(x)
    lib = ops._lib.load()
    add("flip", lambda: ops._lib.check(lib.lfx_flip(ops._p(x), ops._p(fout), B, S, S, ops._p(dmode), None, B, ops._stream())), 6 * N)
    angles = [rng.uniform(-30, 30) for _ in range(B)]
    params = np.zeros((B, 8), np.int32)
    out_px = 0
    for i, a in enumerate(angles):
        m, nw, nh = augment.rotate_matrix(a, S, S)
        params[i, :6] = augment.fixed_affine(m)
        params[i, 6:] = (nw, nh)
        out_px += nw * nh
    add("rotate_nn (incl. host param upload)", lambda: ops.rotate_nn(x, params), 3 * N + 3 * out_px / B)
    rslab, _ = ops.rotate_nn(x, params)
    rdp = torch.from_numpy(params).to(dev)
    add("rotate_nn", lambda: ops.rotate_nn(x, params, 255, rdp, rslab), 3 * N + 3 * out_px / B)
    coeffs = np.array([[1 + s, 0, -s * S, 0, 1 + s, -s * S, 0, 0] for s in (rng.uniform(0.05, 0.15) for _ in range(B))])
    dco = torch.from_numpy(coeffs).to(dev)
    dpe = torch.ones(B, dtype=torch.int32, device=dev)
    add("warp_bicubic_skew", lambda: ops.warp_bicubic(x, dco, dpe), 6 * N)
    sh = np.array([[1, k, 0, 0, 1, 0, 0, 0] if rng.random() < 0.5 else [1, 0, 0, k, 1, 0, 0, 0] for k in (rng.uniform(-0.2, 0.2) for _ in range(B))], np.float64)
    dsh = torch.from_numpy(sh).to(dev)
    dpa = torch.zeros(B, dtype=torch.int32, device=dev)
    add("warp_bicubic_shear", lambda: ops.warp_bicubic(x, dsh, dpa), 6 * N)
    boxes = np.zeros((B, 4), np.int32)
    crop_px = 0
    for i in range(B):
        r = rng.uniform(0.8, 0.95)
        nw, nh = int(S * r), int(S * r)
        boxes[i] = (rng.randint(0, S - nw), rng.randint(0, S - nh), nw, nh)
        crop_px += nw * nh
    plan = ops.CropPlan(boxes, (S, S), dev)
    cout = torch.empty((B, S, S, 3), dtype=torch.uint8, device=dev)
    add("crop_lanczos", lambda: ops.crop_lanczos(x, plan, out=cout), 3 * crop_px / B + 3 * N)
    plan224 = ops.CropPlan(np.tile(np.array([0, 0, S, S], np.int32), (B, 1)), (224, 224), dev)
    o8 = torch.empty((B, 224, 224, 3), dtype=torch.uint8, device=dev)
    of = torch.empty((B, 224, 224, 3), dtype=torch.float32, device=dev)
    add("resize224_normalize", lambda: ops.crop_lanczos(x, plan224, want_f32=True, out=o8, outf=of),
        3 * N + 224 * 224 * 3 + 224 * 224 * 3 * 4)
    noise = torch.randint(0, 256, x.shape, dtype=torch.uint8, device=dev)
    cuts = [int(N * rng.uniform(0, 2) // 100) for _ in range(B)]
    add("distort", lambda: ops.distort(x, noise, cuts), 9 * N)
    seeds = np.array([rng.randint(1, 1000000) for _ in range(B)], np.int64)
    dseeds = torch.from_numpy((seeds & 0xFFFFFFFF).astype(np.uint32).view(np.int32)).to(dev)
    t = timed(lambda: ops.legacy_normal_noise(seeds, 3 * N, 5.0, dev, dseeds=dseeds), args.reps)
    res["legacy_normal_noise (device MT19937 + polar gauss)"] = {"ms": round(t, 4), "algo_bytes_per_image": 3 * N,
                                                                  "Gsamples/s": round(B * 3 * N / (t / 1e3) / 1e9, 1),
                                                                  "images_per_s": round(B / (t / 1e3))}
    print(json.dumps({"batch": B, "size": S, "peak_GB/s": peak, "ops": res}, indent=1))


if __name__ == "__main__":
    main()
