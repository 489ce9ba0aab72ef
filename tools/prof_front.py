#!/usr/bin/env python3
"""The default-strategy profile (k_front inclusive + k_core on its candidate) on a small resident batch, for ncu captures."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from leaffliction_b200 import engine, ops, synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda:0")
base = synth.leaf_batch(128, 256, 256)
x = torch.from_numpy(np.concatenate([base] * (B // 128))).to(dev)
e = engine.TransformEngine(256, 256, ops.mask_cfg("hsv_h"), 1.5, (256, 256), dev, front="inclusive")
out = ops.alloc_core_outputs(B, 256, 256, (256, 256), dev)
for _ in range(2):
    e.run_device(x, out)
torch.cuda.synchronize()
print("ok")
