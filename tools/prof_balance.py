#!/usr/bin/env python3
"""Device time of every stage of the class-balancing pass (config 3, 36,864 tasks), CUDA events, second run of each stage.
Diagnostic only (tools/bench_balance.py is the throughput measurement)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from leaffliction_b200 import augment, balance, ops, synth  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    counts = balance.synthetic_class_counts()
    names = [c for p in counts.values() for c in p]
    plants = {p: list(c) for p, c in counts.items()}
    labels = np.repeat(np.arange(len(names)), [n for p in counts.values() for n in p.values()])
    N, S = len(labels), 256
    x = torch.from_numpy(synth.leaf_batch(128, S, S)).to(dev).repeat(N // 128, 1, 1, 1).contiguous()
    plan, tasks = balance.tasks_for_labels(labels, names, plants, seed=42)
    ta = augment.TaskArrays(tasks)
    ip, dp = augment.draw_params_batch(ta.transform, ta.seed, S, S)
    code = augment.TRANSFORM_CODE
    res = {}

    def ids_of(*n):
        return np.nonzero(np.isin(ta.transform, [code[k] for k in n]))[0]

    def dev_time(name, fn):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = fn()
        b.record()
        torch.cuda.synchronize()
        res[name] = round(a.elapsed_time(b), 3)
        return r
    src = torch.from_numpy(ta.source_index).to(dev)
    g = {}
    for key, nn in (("flip", ("flip",)), ("rotate", ("rotate",)), ("warp", ("skew", "shear")), ("crop", ("crop",)), ("distortion", ("distortion",))):
        ids = ids_of(*nn)
        sel = src[torch.from_numpy(ids).to(dev)]
        g[key] = (ids, dev_time(f"gather_{key}_ms", lambda: x.index_select(0, sel)))
    fm = torch.from_numpy(ip[g["flip"][0], 0]).to(dev)
    dev_time("flip_ms", lambda: ops.flip(g["flip"][1], fm))
    rp = ip[g["rotate"][0]]
    drp = torch.from_numpy(rp).to(dev)
    slab, _ = ops.rotate_nn(g["rotate"][1], rp, 255, dparams=drp)
    dev_time("rotate_ms", lambda: ops.rotate_nn(g["rotate"][1], rp, 255, dparams=drp, out=slab))
    wc, wp = torch.from_numpy(dp[g["warp"][0]]).to(dev), torch.from_numpy(ip[g["warp"][0], 0]).to(dev)
    dev_time("warp_ms", lambda: ops.warp_bicubic(g["warp"][1], wc, wp))
    cplan = ops.CropPlan(ip[g["crop"][0], :4], (S, S), dev)
    dev_time("crop_ms", lambda: ops.crop_lanczos(g["crop"][1], cplan))
    ids = g["distortion"][0]
    noise = dev_time("noise_mt19937_ms", lambda: ops.legacy_normal_noise(ta.seed[ids], S * S * 3, 5, dev).view(len(ids), S, S, 3))
    cuts = torch.from_numpy(ip[ids, 0]).to(dev)
    dev_time("distort_ms", lambda: ops.distort(g["distortion"][1], noise, cuts))
    res["sum_ms"] = round(sum(v for k, v in res.items()), 3)
    dev_time("augment_device_total_ms", lambda: augment.augment_device(x, ta))
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
