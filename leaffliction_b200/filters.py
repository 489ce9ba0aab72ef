"""Drop-in for srcs/transform/filters/{blur,analyze,brown,hist}.py (same signatures).

  apply_blur_filter(rgb, cfg, make_mask_func) -> rgb           blur.py:18-79
  apply_brown_filter(rgb, mask, cfg) -> (vis, pct, count)      brown.py:21-89
  apply_analyze_filter(rgb, mask, contour, cfg) -> rgb         analyze.py:20-124 (numeric record + vein edges;
                                                               anti-aliased overlay drawing is SURVEY.md 8f #3)
  apply_histogram_filter(rgb, cfg)                             hist.py:181-300 (numeric core; the matplotlib
                                                               figure is out of scope)
The arithmetic runs in libleafx (CUDA) through the C ABI.
"""
from __future__ import annotations

import logging
from typing import Dict, Optional, Tuple

import numpy as np

from .transform import TransformConfig, _dev, _ops, mask_cfg_from

HIST_CATEGORIES = ("Vert Sain", "Vert Jaunâtre", "Jaune", "Brun/Orange", "Rouge", "Zones Sombres", "Zones Claires",
                   "Violet/Pourpre")
HUE_RANGES = ("Vert (35-85°)", "Jaune/Orange (15-35°)", "Rouge (0-15° & 160-180°)", "Violet (120-160°)", "Autres")


def _mask2d(mask: np.ndarray) -> np.ndarray:
    m = mask if mask.ndim == 2 else mask[..., 0]
    return (m > 0).astype(np.uint8) * 255


def apply_blur_filter(rgb: np.ndarray, cfg: TransformConfig, make_mask_func) -> np.ndarray:
    mask, _ = make_mask_func(rgb)
    if mask is None:
        return rgb
    ops = _ops()
    out = ops.saliency_blur(_dev(rgb[None]), _dev(_mask2d(mask)[None]), mask_cfg_from(cfg), cfg.gaussian_sigma)
    return out.cpu().numpy()[0]


def apply_brown_filter(rgb: np.ndarray, mask: Optional[np.ndarray], cfg: TransformConfig) -> Tuple[np.ndarray, float, int]:
    if mask is None:
        return rgb, 0.0, 0
    ops = _ops()
    spots, stats = ops.brown_spots(_dev(rgb[None]), _dev(_mask2d(mask)[None]), mask_cfg_from(cfg))
    spots = spots.cpu().numpy()[0]
    leaf_px, count, spot_px, _ = (int(v) for v in stats.cpu().numpy()[0])
    pct = (spot_px / max(leaf_px, 1)) * 100
    vis = rgb.copy()
    vis[spots > 0] = (255, 100, 0)
    logging.info(f"Brown spots detected: {count} regions, {pct:.1f}% of leaf area ({spot_px} pixels)")
    return vis, pct, count


def record_from_device(rec_i: np.ndarray, rec_f: np.ndarray, hull: np.ndarray) -> Dict:
    """One image's lfx_analyze_record output (include/leafx.h layout) as the dictionary analyze_record returns."""
    nh = int(rec_i[12])
    return dict(centroid=(int(rec_i[2]), int(rec_i[3])), area=float(rec_f[0]), n_points=int(rec_i[1]),
                left=(int(rec_i[4]), int(rec_i[5])), right=(int(rec_i[6]), int(rec_i[7])),
                top=(int(rec_i[8]), int(rec_i[9])), bottom=(int(rec_i[10]), int(rec_i[11])),
                hull=hull[:nh].reshape(-1, 1, 2).astype(np.int32), hull_area=float(rec_f[1]),
                pca_mean=rec_f[2:4].copy(), pca_eigenvectors=rec_f[4:8].reshape(2, 2).copy(), pca_eigenvalues=rec_f[8:10].copy(),
                axes=(((int(rec_i[14]), int(rec_i[15])), (int(rec_i[16]), int(rec_i[17]))),
                      ((int(rec_i[18]), int(rec_i[19])), (int(rec_i[20]), int(rec_i[21])))))


def _device_record(rgb: np.ndarray, mask: np.ndarray, contour: np.ndarray):
    """(device record of lfx_analyze_record for one contour, device rgb, device edges, device mask)."""
    import torch
    ops = _ops()
    pts = np.ascontiguousarray(contour[:, 0, :], np.int32)
    H, W = rgb.shape[:2]
    d_pts = _dev(pts[None])
    d_cnt = torch.tensor([len(pts)], dtype=torch.int32, device=d_pts.device)
    max_hull = 512
    while True:
        rec = ops.analyze_points(d_pts, d_cnt, H, W, max_hull)
        nh = int(rec["rec_i"][0, 12])
        if nh >= 0:
            break
        max_hull = -nh + 8
    d_rgb = _dev(rgb[None])
    edges = ops.canny(ops.cvt_color(d_rgb, "gray"), 80, 160, True)
    return rec, d_rgb, edges, _dev(np.ascontiguousarray(_mask2d(mask))[None])


def analyze_record(rgb: np.ndarray, mask: np.ndarray, contour: np.ndarray) -> Dict:
    """Numeric content of apply_analyze_filter (analyze.py:43-122) for one image: centroid of the contour polygon
    (cv2.moments), extreme points, convex hull, PCA axes (all by lfx_analyze_record on the device), vein-edge mask
    (Canny 80/160 L2 inside the mask).  The batched form is ops.analyze_records / TransformEngine(analyze=True)."""
    rec, _, edges, d_mask = _device_record(rgb, mask, contour)
    out = record_from_device(rec["rec_i"].cpu().numpy()[0], rec["rec_f"].cpu().numpy()[0], rec["hull"].cpu().numpy()[0])
    out["veins"] = ((edges > 0) & (d_mask > 0)).cpu().numpy()[0]
    return out


def apply_analyze_filter(rgb: np.ndarray, mask: Optional[np.ndarray], contour: Optional[np.ndarray],
                         cfg: TransformConfig) -> np.ndarray:
    """The overlay of analyze.py:37-122, drawn on the device by lfx_analyze_overlay with OpenCV's rasterisers restated
    (contour, centroid cross, extreme points with anti-aliased rays, anti-aliased hull, PCA axes, cyan vein edges):
    bit-identical to the reference's image whenever the PCA end points are (no tied projections).  Without a contour the
    reference returns the image with an "Analyze: no object" text banner (putText); here the image is returned unchanged."""
    if contour is None or mask is None:
        return rgb.copy()
    rec, d_rgb, edges, d_mask = _device_record(rgb, mask, contour)
    return _ops().analyze_overlay(d_rgb, rec, edges, d_mask).cpu().numpy()[0]


def histogram_stats(rgb: np.ndarray) -> Dict:
    """hist.py numeric core on an (already masked) image: leaf_mask (:188), the eight category
    percentages of _analyze_color_regions (:38-65), 256-bin H/S/V histograms over leaf_mask (the 60-bin
    density plot of :140-168 derives from them) and the five hue-range counts (:248-256)."""
    ops = _ops()
    full = np.full(rgb.shape[:2], 255, np.uint8)
    _, h3, cn = ops.color_stats(_dev(rgb[None]), _dev(full[None]), hist9=False)
    h3 = h3.cpu().numpy()[0]
    cn = cn.cpu().numpy()[0].astype(np.int64)
    total = int(cn[0])
    analysis = {} if total == 0 else {k: (int(cn[1 + i]) / total) * 100 for i, k in enumerate(HIST_CATEGORIES)}
    return dict(total_pixels=total, color_analysis=analysis, hsv_hist=h3,
                hue_ranges={k: int(cn[9 + i]) for i, k in enumerate(HUE_RANGES)})


def apply_histogram_filter(rgb: np.ndarray, cfg: TransformConfig):
    """The reference returns a rasterised matplotlib figure (hist.py:191-300); figure rendering is host
    plotting and out of scope.  Returns the numeric statistics instead."""
    return histogram_stats(rgb)
