import sys, numpy as np, torch
sys.path.insert(0,'/root/repo')
from leaffliction_b200 import ops, synth
B,S=592,256
dev=torch.device('cuda:0')
base=synth.leaf_batch(37,S,S)
x=torch.from_numpy(np.concatenate([base]*16)).to(dev)
cfg=ops.mask_cfg("hsv_h")
mask,info=ops.make_mask(x,cfg)
for _ in range(2):
    ops.raw_mask_front_end(x,"inclusive",cfg); ops.saliency_blur(x,mask,cfg,1.5); ops.brown_spots(x,mask,cfg)
torch.cuda.synchronize()
