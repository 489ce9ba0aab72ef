"""Host vs device time of augment_device for the 36,864 balancing tasks (diagnostic)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from leaffliction_b200 import augment, balance, synth
dev = torch.device("cuda:0")
counts = balance.synthetic_class_counts()
names = [c for p in counts.values() for c in p]
plants = {p: list(c) for p, c in counts.items()}
labels = np.repeat(np.arange(len(names)), [n for p in counts.values() for n in p.values()])
x = torch.from_numpy(synth.leaf_batch(128, 256, 256)).to(dev).repeat(len(labels) // 128, 1, 1, 1).contiguous()
plan, tasks = balance.tasks_for_labels(labels, names, plants, seed=42)
ta = augment.TaskArrays(tasks)
for _ in range(2):
    augment.augment_device(x, ta); torch.cuda.synchronize()
import cProfile, pstats
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = augment.augment_device(x, ta)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"host return {1e3*(t1-t0):.2f} ms, device done {1e3*(t2-t0):.2f} ms")
    del r
pr = cProfile.Profile(); pr.enable(); r = augment.augment_device(x, ta); pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
