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


def analyze_record(rgb: np.ndarray, mask: np.ndarray, contour: np.ndarray) -> Dict:
    """Numeric content of apply_analyze_filter: centroid of the contour polygon (cv2.moments),
    extreme points, bounding box, polygon area, vein-edge mask (Canny 80/160 L2 inside the mask)."""
    pts = contour[:, 0, :].astype(np.int64)
    n = len(pts)
    prev = np.roll(pts, 1, axis=0)
    d = prev[:, 0] * pts[:, 1] - pts[:, 0] * prev[:, 1]          # Green's formula, exact in int64
    a00 = float(d.sum())
    a10 = float((d * (prev[:, 0] + pts[:, 0])).sum())
    a01 = float((d * (prev[:, 1] + pts[:, 1])).sum())
    if abs(a00) > 1.1920928955078125e-07:
        sg = 1.0 if a00 > 0 else -1.0
        m00, m10, m01 = a00 * (0.5 * sg), a10 * (0.16666666666666666 * sg), a01 * (0.16666666666666666 * sg)
        cx, cy = int(m10 / m00), int(m01 / m00)
    else:
        m00 = 0.0
        cm = contour[:, 0, :].mean(axis=0)
        cx, cy = int(cm[0]), int(cm[1])
    p = contour[:, 0, :]
    ops = _ops()
    gray = ops.cvt_color(_dev(rgb[None]), "gray")
    edges = ops.canny(gray, 80, 160, True).cpu().numpy()[0]
    veins = (edges > 0) & (_mask2d(mask) > 0)
    hull = convex_hull(p)
    mean, evecs, evals, ends = pca_axes(p)
    return dict(centroid=(cx, cy), area=m00, n_points=n,
                left=tuple(p[p[:, 0].argmin()]), right=tuple(p[p[:, 0].argmax()]),
                top=tuple(p[p[:, 1].argmin()]), bottom=tuple(p[p[:, 1].argmax()]), veins=veins,
                hull=hull, pca_mean=mean, pca_eigenvectors=evecs, pca_eigenvalues=evals, axes=ends)


def convex_hull(pts: np.ndarray) -> np.ndarray:
    """Vertices of the convex hull of integer points [K,2] -> int32 [M,1,2] (analyze.py:77 cv2.convexHull):
    Andrew's monotone chain in exact integer arithmetic, collinear points dropped."""
    q = np.unique(np.asarray(pts, np.int64).reshape(-1, 2), axis=0)      # sorted by x, then y
    if len(q) <= 2:
        return q.astype(np.int32).reshape(-1, 1, 2)

    def half(seq):
        out = []
        for x, y in seq:
            while len(out) >= 2 and (out[-1][0] - out[-2][0]) * (y - out[-2][1]) - (out[-1][1] - out[-2][1]) * (x - out[-2][0]) <= 0:
                out.pop()
            out.append((int(x), int(y)))
        return out
    lower, upper = half(q), half(q[::-1])
    return np.array(lower[:-1] + upper[:-1], np.int32).reshape(-1, 1, 2)


def pca_axes(pts: np.ndarray):
    """analyze.py:88-98: PCA of the contour points (cv2.PCACompute2 on float32 data) and the contour points with the
    extreme projections on the major / minor axis -> (mean[2], eigenvectors[2,2] rows, eigenvalues[2],
    ((p0_min, p0_max), (p1_min, p1_max)))."""
    d = np.asarray(pts, np.float32).reshape(-1, 2)
    mean = d.mean(axis=0, dtype=np.float64)
    c = d.astype(np.float64) - mean
    cov = (c.T @ c) / max(len(d), 1)
    w, v = np.linalg.eigh(cov)
    order = np.argsort(w)[::-1]
    evals, evecs = w[order], v[:, order].T
    ends = []
    for k in range(2):
        proj = d.astype(np.float64) @ evecs[k]
        ends.append((tuple(int(t) for t in d[int(proj.argmin())]), tuple(int(t) for t in d[int(proj.argmax())])))
    return mean, evecs, evals, tuple(ends)


def apply_analyze_filter(rgb: np.ndarray, mask: Optional[np.ndarray], contour: Optional[np.ndarray],
                         cfg: TransformConfig) -> np.ndarray:
    """Overlay with the exact (non anti-aliased) elements: cyan vein edges, centroid and extreme-point
    markers.  Lines/hull/PCA axes (LINE_AA drawing) are cosmetic and not rasterised here."""
    if contour is None or mask is None:
        return rgb.copy()
    rec = analyze_record(rgb, mask, contour)
    overlay = rgb.copy()
    H, W = overlay.shape[:2]
    for (x, y) in (rec["left"], rec["right"], rec["top"], rec["bottom"], rec["centroid"]):
        overlay[max(0, int(y) - 2):min(H, int(y) + 3), max(0, int(x) - 2):min(W, int(x) + 3)] = (255, 255, 0)
    overlay[rec["veins"]] = (0, 255, 255)
    return overlay


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
