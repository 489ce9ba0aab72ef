import numpy as np, torch, sys
sys.path.insert(0, '.')
from leaffliction_b200 import engine, ops, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
imgs = synth.leaf_batch(128, 256, 256)
x = torch.from_numpy(np.concatenate([imgs] * (B // 128))).pin_memory()
eng = engine.TransformEngine(256, 256, ops.mask_cfg("hsv_h"), 1.5, (256, 256), torch.device("cuda:0"), chunk=512, augment=True)
seeds = np.random.default_rng(1).integers(1, 1000001, (6, B)).astype(np.int64)
for it in range(3):
    out = eng.run_host(x, seeds=seeds)
    torch.cuda.synchronize()
    print("iter", it, "ok", int(out.aug["flip"].sum()))
