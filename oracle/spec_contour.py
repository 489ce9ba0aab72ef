"""Spec of cv2.findContours(RETR_EXTERNAL, CHAIN_APPROX_SIMPLE) for ONE 8-connected component
(Suzuki-Abe border following as implemented by OpenCV's icvFetchContour), plus cv2.moments /
boundingRect of the resulting polygon.

Restates largest_contour (srcs/cli/Transformation.py:285-292) and the contour consumers
roi.py:26 (boundingRect) and analyze.py:43-64 (moments, extreme points).
Pure-Python loops: small cases only.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
from __future__ import annotations

import numpy as np

# direction codes 0..7: E, NE, N, NW, W, SW, S, SE  (x right, y down)
DX = (1, 1, 0, -1, -1, -1, 0, 1)
DY = (0, -1, -1, -1, 0, 1, 1, 1)


def trace_external(mask: np.ndarray, start_xy) -> np.ndarray:
    """Outer border of the component whose first raster pixel is start_xy -> int32 [K,1,2]."""
    H, W = mask.shape

    def on(x, y):
        return 0 <= x < W and 0 <= y < H and mask[y, x] != 0
    x0, y0 = start_xy
    pts = []
    s_end = s = 4
    while True:
        s = (s - 1) & 7
        if on(x0 + DX[s], y0 + DY[s]) or s == s_end:
            break
    if not on(x0 + DX[s], y0 + DY[s]):
        return np.array([[[x0, y0]]], np.int32)                 # isolated pixel
    x1, y1 = x0 + DX[s], y0 + DY[s]
    x3, y3 = x0, y0
    prev_s = s ^ 4
    px, py = x0, y0
    while True:
        s_end = s
        while True:
            s += 1
            x4, y4 = x3 + DX[s & 7], y3 + DY[s & 7]
            if on(x4, y4):
                break
        s &= 7
        if s != prev_s:
            pts.append((px, py))
            prev_s = s
        px += DX[s]
        py += DY[s]
        if (x4, y4) == (x0, y0) and (x3, y3) == (x1, y1):
            break
        x3, y3 = x4, y4
        s = (s + 4) & 7
    return np.array(pts, np.int32).reshape(-1, 1, 2)


def moments_polygon(cnt: np.ndarray):
    """cv2.moments(contour) m00, m10, m01 (contourMoments, Green's formula in float64)."""
    p = cnt.reshape(-1, 2).astype(np.float64)
    n = len(p)
    if n == 0:
        return 0.0, 0.0, 0.0
    a00 = a10 = a01 = 0.0
    xi_1, yi_1 = p[n - 1]
    for i in range(n):
        xi, yi = p[i]
        dxy = xi_1 * yi - xi * yi_1
        xii_1 = xi_1 + xi
        yii_1 = yi_1 + yi
        a00 += dxy
        a10 += dxy * xii_1
        a01 += dxy * yii_1
        xi_1, yi_1 = xi, yi
    if abs(a00) <= 1.1920928955078125e-07:
        return 0.0, 0.0, 0.0
    db1_2, db1_6 = 0.5, 0.16666666666666666666666666666667
    if a00 > 0:
        return a00 * db1_2, a10 * db1_6, a01 * db1_6
    return a00 * -db1_2, a10 * -db1_6, a01 * -db1_6


def analyze_record(cnt: np.ndarray):
    """Numeric record of apply_analyze_filter (analyze.py:43-64): centroid + extreme points."""
    m00, m10, m01 = moments_polygon(cnt)
    pts = cnt[:, 0, :]
    if m00 != 0:
        cx, cy = int(m10 / m00), int(m01 / m00)
    else:
        cm = pts.mean(axis=0)
        cx, cy = int(cm[0]), int(cm[1])
    left = tuple(pts[pts[:, 0].argmin()])
    right = tuple(pts[pts[:, 0].argmax()])
    top = tuple(pts[pts[:, 1].argmin()])
    bottom = tuple(pts[pts[:, 1].argmax()])
    return dict(centroid=(cx, cy), left=left, right=right, top=top, bottom=bottom)
