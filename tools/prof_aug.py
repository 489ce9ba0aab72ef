"""ncu driver for the augment kernels: a few launches of warp_bicubic (skew, shear-x, shear-y), crop_lanczos and the
device noise stream.  Diagnostic only."""
import random
import sys

import numpy as np
import torch

sys.path.insert(0, '/root/repo')
from leaffliction_b200 import ops, synth  # noqa: E402

B, S = 1024, 256
dev = torch.device('cuda:0')
base = synth.leaf_batch(32, S, S)
x = torch.from_numpy(np.concatenate([base] * (B // 32))).to(dev)
rng = random.Random(1)
skew = np.array([[1 + s, 0, -s * S, 0, 1 + s, -s * S, 0, 0] for s in (rng.uniform(0.05, 0.15) for _ in range(B))])
shx = np.array([[1, k, 0, 0, 1, 0, 0, 0] for k in (rng.uniform(-0.2, 0.2) for _ in range(B))], np.float64)
shy = np.array([[1, 0, 0, k, 1, 0, 0, 0] for k in (rng.uniform(-0.2, 0.2) for _ in range(B))], np.float64)
boxes = np.zeros((B, 4), np.int32)
for i in range(B):
    r = rng.uniform(0.8, 0.95)
    nw = nh = int(S * r)
    boxes[i] = (rng.randint(0, S - nw), rng.randint(0, S - nh), nw, nh)
seeds = [rng.randint(1, 1000000) for _ in range(256)]
for _ in range(2):
    ops.warp_bicubic(x, skew, [True] * B)
    ops.warp_bicubic(x, shx, [False] * B)
    ops.warp_bicubic(x, shy, [False] * B)
    ops.crop_lanczos(x, boxes, (S, S))
    ops.legacy_normal_noise(seeds, S * S * 3, 5.0, dev)
torch.cuda.synchronize()
