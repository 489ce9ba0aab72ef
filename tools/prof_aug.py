import sys, random, numpy as np, torch
sys.path.insert(0,'/root/repo')
from leaffliction_b200 import ops, synth
B,S=1024,256
dev=torch.device('cuda:0')
base=synth.leaf_batch(32,S,S)
x=torch.from_numpy(np.concatenate([base]*(B//32))).to(dev)
rng=random.Random(1)
coeffs=np.array([[1+s,0,-s*S,0,1+s,-s*S,0,0] for s in (rng.uniform(0.05,0.15) for _ in range(B))])
boxes=np.zeros((B,4),np.int32)
for i in range(B):
    r=rng.uniform(0.8,0.95); nw=nh=int(S*r); boxes[i]=(rng.randint(0,S-nw),rng.randint(0,S-nh),nw,nh)
for _ in range(3):
    ops.warp_bicubic(x,coeffs,[True]*B); ops.crop_lanczos(x,boxes,(S,S))
torch.cuda.synchronize()
