#!/usr/bin/env python3
"""Where the class-balancing wall time goes (config 3): host stages with perf_counter, device stages with CUDA
events + synchronize.  Diagnostic only (tools/bench_balance.py is the throughput measurement)."""
import json
import os
import sys
import time

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
    t0 = time.perf_counter()
    plan, tasks = balance.tasks_for_labels(labels, names, plants, seed=42)
    res = {"tasks_for_labels_s": time.perf_counter() - t0}
    augment.augment_device(x, tasks[:64])
    torch.cuda.synchronize()

    def host(name, fn):
        t = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        res[name] = round(time.perf_counter() - t, 5)
        return r
    ta = host("TaskArrays_s", lambda: augment.TaskArrays(tasks))
    ip, dp = host("draw_params_batch_s", lambda: augment.draw_params_batch(ta.transform, ta.seed, S, S))
    src = torch.from_numpy(ta.source_index).to(dev)
    code = augment.TRANSFORM_CODE

    def ids_of(*n):
        return np.nonzero(np.isin(ta.transform, [code[k] for k in n]))[0]
    g = {}
    for key, nn in (("flip", ("flip",)), ("rotate", ("rotate",)), ("warp", ("skew", "shear")), ("crop", ("crop",)), ("distortion", ("distortion",))):
        ids = ids_of(*nn)
        g[key] = (ids, host(f"gather_{key}_s", lambda: x.index_select(0, src[torch.from_numpy(ids).to(dev)])))
    host("flip_s", lambda: ops.flip(g["flip"][1], ip[g["flip"][0], 0]))
    host("rotate_s", lambda: ops.rotate_nn(g["rotate"][1], ip[g["rotate"][0]], 255))
    host("warp_s", lambda: ops.warp_bicubic(g["warp"][1], dp[g["warp"][0]], ip[g["warp"][0], 0]))
    host("crop_s", lambda: ops.crop_lanczos(g["crop"][1], ip[g["crop"][0], :4], (S, S)))
    ids = g["distortion"][0]
    noise = host("noise_mt19937_s", lambda: ops.legacy_normal_noise(ta.seed[ids], S * S * 3, 5, dev).view(len(ids), S, S, 3))
    host("distort_s", lambda: ops.distort(g["distortion"][1], noise, ip[ids, 0]))
    host("augment_device_total_s", lambda: augment.augment_device(x, ta))
    host("augment_device_from_list_s", lambda: augment.augment_device(x, tasks))
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
