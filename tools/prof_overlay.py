"""One launch of the overlay kernels on 4096 x 256x256 leaves (for `ncu -k regex:"k_analyze_overlay|k_draw_rectangles"`), and their
CUDA-event times.  `python tools/prof_overlay.py [batch]`."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from leaffliction_b200 import ops, synth  # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    dev = torch.device("cuda:0")
    base = synth.leaf_batch(64, 256, 256)
    x = torch.from_numpy(np.concatenate([base] * ((B + 63) // 64))[:B]).to(dev)
    mask, info = ops.make_mask(x, ops.mask_cfg("hsv_h"))
    rec = ops.analyze_records(mask, info, max_pts=8192, max_hull=1000)
    edges = ops.canny(ops.cvt_color(x, "gray"), 80, 160, True)
    print("contour points mean / max", float(rec["counts"].float().mean()), int(rec["counts"].max()),
          "hull vertices mean", float(rec["rec_i"][:, 12].float().mean()),
          "vein pixels per image", float(((edges > 0) & (mask > 0)).sum()) / B)
    print("analyze_overlay ms", timed(lambda: ops.analyze_overlay(x, rec, edges, mask)))
    print("draw_rectangles ms", timed(lambda: ops.draw_rectangles(x, info)))


if __name__ == "__main__":
    main()
