import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
from leaffliction_b200 import ops
dev = torch.device('cuda:0')
seeds = list(range(1000, 1000 + 6144))
for _ in range(2):
    ops.legacy_normal_noise(seeds, 196608, 5.0, dev)
torch.cuda.synchronize()
