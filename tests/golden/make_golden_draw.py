"""Golden overlays from the reference's OWN functions (build container only: needs /root/reference):
apply_analyze_filter (srcs/transform/filters/analyze.py:20-124) and the `vis` image of apply_roi_filter
(srcs/transform/filters/roi.py:20-46) on seeded synthetic leaves -> tests/golden/golden_draw_v1.npz.

    python tests/golden/make_golden_draw.py

Inputs (image, mask, contour) are stored next to the outputs, so the tests need neither the reference nor OpenCV."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

import ref_harness  # noqa: E402
from leaffliction_b200 import synth  # noqa: E402

CASES = [("s64", 64, 64, 3, 41), ("s96x64", 96, 64, 2, 42), ("s256", 256, 256, 2, 43)]


def main():
    import cv2
    ns = ref_harness.load()
    cfg = ref_harness.ref_config(ns, mask_strategy="hsv_h")
    out = {}
    for tag, h, w, n, seed in CASES:
        for i, img in enumerate(synth.leaf_batch(n, h, w, seed)):
            mask, contour = ns.mask.make_mask(img, cfg)
            if contour is None:
                continue
            masked = ns.mask_utils.apply_mask(img, mask, mask_color="white") if hasattr(ns.mask_utils, "apply_mask") else img
            key = f"{tag}_{i}"
            out[key + "_rgb"] = masked
            out[key + "_mask"] = mask
            out[key + "_contour"] = contour.astype(np.int32)
            out[key + "_analyze"] = ns.analyze.apply_analyze_filter(masked, mask, contour, cfg)
            _, vis, box = ns.roi.apply_roi_filter(masked, contour, cfg)
            out[key + "_roi_vis"] = vis
            out[key + "_roi_box"] = np.array(box, np.int32)
            out[key + "_edges"] = cv2.Canny(cv2.cvtColor(masked, cv2.COLOR_RGB2GRAY), 80, 160, L2gradient=True)
    path = os.path.join(HERE, "golden_draw_v1.npz")
    np.savez_compressed(path, **out)
    print(path, len(out), "arrays", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
