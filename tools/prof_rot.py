"""ncu driver: rotate_nn / crop_lanczos / distort on 1024 images."""
import random, sys
import numpy as np, torch
sys.path.insert(0, '/root/repo')
from leaffliction_b200 import augment, ops, synth
B, S = 1024, 256
dev = torch.device('cuda:0')
base = synth.leaf_batch(32, S, S)
x = torch.from_numpy(np.concatenate([base] * (B // 32))).to(dev)
rng = random.Random(1)
params = np.zeros((B, 8), np.int32)
for i in range(B):
    m, nw, nh = augment.rotate_matrix(rng.uniform(-30, 30), S, S)
    params[i, :6] = augment.fixed_affine(m); params[i, 6:] = (nw, nh)
for _ in range(2):
    ops.rotate_nn(x, params)
torch.cuda.synchronize()
