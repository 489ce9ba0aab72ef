#!/usr/bin/env python3
"""One bench step (k_core + the 6-op augment set) on a small resident batch, for ncu captures:
  ncu --set full --import-source on --clock-control none -k regex:'k_core|k_legacy|k_warp|k_crop|k_rotate' -o gpurun_out/prof python tools/prof_step.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from leaffliction_b200 import augment, ops, synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda:0")
base = synth.leaf_batch(128, 256, 256)
x = torch.from_numpy(np.concatenate([base] * (B // 128))).to(dev)
out = ops.alloc_core_outputs(B, 256, 256, (256, 256), dev)
ds = torch.zeros((9, 256), dtype=torch.int64, device=dev)
aset = augment.AugmentSet(B, 256, 256, dev)
seeds = np.random.default_rng(1).integers(1, 1000001, size=(6, B), dtype=np.int64)
for _ in range(reps):
    ops.pipeline_core(x, ops.mask_cfg("hsv_h"), 1.5, (256, 256), out, ds)
    aset.run(x, seeds)
torch.cuda.synchronize()
print("ok")
