#!/usr/bin/env python3
"""BASELINE config 3: class-balancing augmentation of a synthetic PlantVillage-shaped dataset (8 classes,
imbalanced; SURVEY.md section 8d counts scaled to --images), images resident in HBM, augment tasks sharded by
index across ranks, class histogram merged by ONE NCCL allreduce.  Not the contract benchmark (bench.py).

  python tools/bench_balance.py --images 65536                       # 1 GPU
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/bench_balance.py
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from leaffliction_b200 import augment, balance, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=65536)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--chunk", type=int, default=65536, help="tasks per augment_device call on one rank")
    ap.add_argument("--passes", type=int, default=3, help="timed passes over the task list (the mean is reported)")
    ap.add_argument("--host-noise", action="store_true", help="draw the distortion noise with np.random on the host (reference way)")
    args = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    counts = balance.synthetic_class_counts()
    scale = args.images / 65536.0
    names = [c for p in counts.values() for c in p]
    plants = {p: list(c) for p, c in counts.items()}
    per_class = [max(1, int(round(n * scale))) for p in counts.values() for n in p.values()]
    labels = np.repeat(np.arange(len(names)), per_class)
    N, S = len(labels), args.size
    base = torch.from_numpy(synth.leaf_batch(128, S, S)).to(dev)
    x = base.repeat((N + 127) // 128, 1, 1, 1)[:N].contiguous()      # dataset resident in HBM

    # class histogram: per-rank partial counts over this rank's image shard, one allreduce
    t0 = time.perf_counter()
    part = np.bincount(labels[list(balance.shard(N, rank, world))], minlength=len(names)).astype(np.int64)
    merged, _ = balance.allreduce_histograms(part, device=dev)
    assert merged.tolist() == per_class
    plan, all_tasks = balance.task_arrays_for_labels(labels, names, plants, seed=42)     # identical on every rank (native stream)
    mine = all_tasks.shard(rank, world)
    t_plan = time.perf_counter() - t0

    def run():
        n_out = 0
        for c0 in range(0, len(mine), args.chunk):
            res = augment.augment_device(x, mine.slice(c0, c0 + args.chunk), device_noise=not args.host_noise)
            n_out += sum(len(v[0]) for v in res.values())
        return n_out

    # warm-up: a strided sample that contains every transform (module load, Lanczos tables, allocator pools)
    step = max(1, len(mine) // 512)
    warm = mine.slice(0, None, step)
    augment.augment_device(x, warm, device_noise=not args.host_noise)
    if not args.host_noise:
        run()                                                          # one full untimed pass (allocator pools at full size)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    passes = []
    for _ in range(args.passes):                                       # every pass redoes the whole task list
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n_out = run()
        torch.cuda.synchronize()
        passes.append(time.perf_counter() - t0)
    dt = sum(passes) / len(passes)
    tt = torch.tensor([dt], dtype=torch.float64, device=dev)
    nn = torch.tensor([n_out], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(nn)
    if rank == 0:
        print(json.dumps({"workload": f"class balancing: {N} images {S}x{S}, 8 classes, {len(all_tasks)} augment tasks (6 ops), images in HBM",
                          "n_gpus": world, "tasks": int(nn.item()), "seconds": float(tt.item()), "passes": args.passes, "augmented_images_per_s": float(nn.item() / tt.item()),
                          "plan_and_histogram_s": t_plan, "noise": "host np.random" if args.host_noise else "device MT19937",
                          "per_transform": {k: sum(v.values()) if isinstance(v, dict) else v for k, v in
                                            {t: sum(p.get(t, 0) for p in plan.values()) for t in balance.TRANSFORMATIONS}.items()}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
