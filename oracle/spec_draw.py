"""Spec of the OpenCV 4.13 rasterisers behind the reference's overlays:

  apply_analyze_filter   srcs/transform/filters/analyze.py:37-122   drawContours(2 px), drawMarker(cross, 2 px), filled
                                                                    circle r=3, anti-aliased 1 px lines / hull polyline,
                                                                    2 px PCA axes, cyan vein edges
  apply_roi_filter       srcs/transform/filters/roi.py:43-44        rectangle(2 px)

OpenCV is a third-party dependency of the reference (requirements.txt, unpinned; 4.13.0 in this image); what is
restated here is its published drawing algorithm (modules/imgproc/src/drawing.cpp): 16.16 fixed-point Line2 /
FillConvexPoly / ThickLine, the midpoint Circle, and the table-driven LineAA.  Pinned live against cv2 in
tests/test_oracle_draw.py (every primitive on random end points, inside and across the image border, and the whole overlay).
Pure-Python loops: small cases only.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
from __future__ import annotations

import math

import numpy as np

XY_SHIFT = 16
XY_ONE = 1 << XY_SHIFT

# LineAA's two constant tables (cross-section filter, 64 entries; slope correction, 32 entries)
FILTER_TABLE = (
    168, 177, 185, 194, 202, 210, 218, 224, 231, 236, 241, 246, 249, 252, 254, 254,
    254, 254, 252, 249, 246, 241, 236, 231, 224, 218, 210, 202, 194, 185, 177, 168,
    158, 149, 140, 131, 122, 114, 105, 97, 89, 82, 75, 68, 62, 56, 50, 45,
    40, 36, 32, 28, 25, 22, 19, 16, 14, 12, 11, 9, 8, 7, 5, 5)
SLOPE_CORR_TABLE = (
    181, 181, 181, 182, 182, 183, 184, 185, 187, 188, 190, 192, 194, 196, 198, 201,
    203, 206, 209, 211, 214, 218, 221, 224, 227, 231, 235, 238, 242, 246, 250, 254)


def _trunc_div(a: int, b: int) -> int:
    """C integer division (truncation towards zero)."""
    q = abs(a) // abs(b)
    return q if (a < 0) == (b < 0) else -q


def _c_double_to_i64(v: float) -> int:
    return int(v)   # truncation, like the (int64) cast


def clip_line(width: int, height: int, p1, p2):
    """cv::clipLine(Size2l, Point2l&, Point2l&): -> (visible, p1, p2)."""
    x1, y1 = p1
    x2, y2 = p2
    right, bottom = width - 1, height - 1
    if width <= 0 or height <= 0:
        return False, p1, p2
    c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8
    c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8
    if (c1 & c2) == 0 and (c1 | c2) != 0:
        if c1 & 12:
            a = 0 if c1 < 8 else bottom
            x1 += _c_double_to_i64(float(a - y1) * float(x2 - x1) / float(y2 - y1))
            y1 = a
            c1 = (x1 < 0) + (x1 > right) * 2
        if c2 & 12:
            a = 0 if c2 < 8 else bottom
            x2 += _c_double_to_i64(float(a - y2) * float(x2 - x1) / float(y2 - y1))
            y2 = a
            c2 = (x2 < 0) + (x2 > right) * 2
        if (c1 & c2) == 0 and (c1 | c2) != 0:
            if c1:
                a = 0 if c1 == 1 else right
                y1 += _c_double_to_i64(float(a - x1) * float(y2 - y1) / float(x2 - x1))
                x1 = a
                c1 = 0
            if c2:
                a = 0 if c2 == 1 else right
                y2 += _c_double_to_i64(float(a - x2) * float(y2 - y1) / float(x2 - x1))
                x2 = a
                c2 = 0
    return (c1 | c2) == 0, (x1, y1), (x2, y2)


def _hline(img, y, xa, xb, color):
    if xa <= xb:
        img[y, xa:xb + 1] = color


def line2(img: np.ndarray, p1, p2, color) -> None:
    """drawing.cpp Line2: 8-connected DDA on 16.16 end points (the outline FillConvexPoly draws when shift != 0)."""
    H, W = img.shape[:2]
    ok, p1, p2 = clip_line(W << XY_SHIFT, H << XY_SHIFT, p1, p2)
    if not ok:
        return
    x1, y1 = p1
    x2, y2 = p2
    dx, dy = x2 - x1, y2 - y1
    ax, ay = abs(dx), abs(dy)

    def put(x, y):
        if 0 <= x < W and 0 <= y < H:
            img[y, x] = color
    if ax > ay:
        if dx < 0:
            dy = -dy
            x1, x2, y1, y2 = x2, x1, y2, y1
        x_step, y_step = XY_ONE, _trunc_div(dy << XY_SHIFT, ax | 1)
        ecount = (x2 - x1) >> XY_SHIFT
    else:
        if dy < 0:
            dx = -dx
            x1, x2, y1, y2 = x2, x1, y2, y1
        x_step, y_step = _trunc_div(dx << XY_SHIFT, ay | 1), XY_ONE
        ecount = (y2 - y1) >> XY_SHIFT
    x1 += XY_ONE >> 1
    y1 += XY_ONE >> 1
    put((x2 + (XY_ONE >> 1)) >> XY_SHIFT, (y2 + (XY_ONE >> 1)) >> XY_SHIFT)
    if ax > ay:
        x = x1 >> XY_SHIFT
        while ecount >= 0:
            put(x, y1 >> XY_SHIFT)
            x += 1
            y1 += y_step
            ecount -= 1
    else:
        y = y1 >> XY_SHIFT
        while ecount >= 0:
            put(x1 >> XY_SHIFT, y)
            x1 += x_step
            y += 1
            ecount -= 1


def fill_convex_poly(img: np.ndarray, v, color) -> None:
    """drawing.cpp FillConvexPoly(line_type = LINE_8, shift = XY_SHIFT) on 16.16 vertices: Line2 outline, then the
    two-edge scan conversion (edge slope rounded once per edge, spans from (x + 0.5) >> 16)."""
    H, W = img.shape[:2]
    npts = len(v)
    delta = XY_ONE >> 1
    delta1 = delta2 = XY_ONE >> 1
    p0 = v[npts - 1]
    xmin = xmax = v[0][0]
    ymin = ymax = v[0][1]
    imin = 0
    for i in range(npts):
        p = v[i]
        if p[1] < ymin:
            ymin = p[1]
            imin = i
        ymax = max(ymax, p[1])
        xmax = max(xmax, p[0])
        xmin = min(xmin, p[0])
        line2(img, p0, p, color)
        p0 = p
    xmin = (xmin + delta) >> XY_SHIFT
    xmax = (xmax + delta) >> XY_SHIFT
    ymin = (ymin + delta) >> XY_SHIFT
    ymax = (ymax + delta) >> XY_SHIFT
    if npts < 3 or xmax < 0 or ymax < 0 or xmin >= W or ymin >= H:
        return
    ymax = min(ymax, H - 1)
    e_idx = [imin, imin]
    e_ye = [ymin, ymin]
    e_di = [1, npts - 1]
    e_x = [-XY_ONE, -XY_ONE]
    e_dx = [0, 0]
    edges = npts
    y = ymin
    while True:
        for i in range(2):
            if y >= e_ye[i]:
                idx0, di = e_idx[i], e_di[i]
                idx = idx0 + di
                if idx >= npts:
                    idx -= npts
                while True:
                    edges -= 1
                    if edges + 1 <= 0:      # `for (; edges-- > 0; )`
                        break
                    ty = (v[idx][1] + delta) >> XY_SHIFT
                    if ty > y:
                        xs, xe = v[idx0][0], v[idx][0]
                        e_ye[i] = ty
                        e_dx[i] = _trunc_div((xe - xs) * 2 + (ty - y), 2 * (ty - y))
                        e_x[i] = xs
                        e_idx[i] = idx
                        break
                    idx0 = idx
                    idx += di
                    if idx >= npts:
                        idx -= npts
        if edges < 0:
            break
        if y >= 0:
            left, right = (1, 0) if e_x[0] > e_x[1] else (0, 1)
            xx1 = (e_x[left] + delta1) >> XY_SHIFT
            xx2 = (e_x[right] + delta2) >> XY_SHIFT
            if xx2 >= 0 and xx1 < W:
                _hline(img, y, max(xx1, 0), min(xx2, W - 1), color)
        e_x[0] += e_dx[0]
        e_x[1] += e_dx[1]
        y += 1
        if y > ymax:
            break


def circle_filled(img: np.ndarray, center, radius: int, color) -> None:
    """drawing.cpp Circle(fill = 1): midpoint circle, four clipped spans per step."""
    H, W = img.shape[:2]
    cx, cy = center
    err, dx, dy, plus, minus = 0, radius, 0, 1, (radius << 1) - 1

    def span(y, xa, xb):
        if 0 <= y < H:
            xa, xb = max(xa, 0), min(xb, W - 1)
            _hline(img, y, xa, xb, color)
    while dx >= dy:
        span(cy - dy, cx - dx, cx + dx)
        span(cy + dy, cx - dx, cx + dx)
        span(cy - dx, cx - dy, cx + dy)
        span(cy + dx, cx - dy, cx + dy)
        dy += 1
        err += plus
        plus += 2
        mask = -1 if err > 0 else 0
        err -= minus & mask
        dx += mask
        minus -= mask & 2


def _cv_round(v: float) -> int:
    """cvRound: round half to even (lrint)."""
    return int(np.rint(v))


def thick_line(img: np.ndarray, p0, p1, color, thickness: int, flags: int = 3) -> None:
    """drawing.cpp ThickLine for LINE_8, shift = 0, thickness >= 2 (integer pixel end points): the 4-vertex polygon
    around the segment + a filled circle of radius thickness/2 at the end points selected by `flags`."""
    assert thickness >= 2
    H, W = img.shape[:2]
    # the segment is first clipped to the image rectangle grown by `thickness` on every side (integer pixel units)
    ok, a, b = clip_line(W + 2 * thickness, H + 2 * thickness, (p0[0] + thickness, p0[1] + thickness),
                         (p1[0] + thickness, p1[1] + thickness))
    if not ok:
        return
    p0, p1 = (a[0] - thickness, a[1] - thickness), (b[0] - thickness, b[1] - thickness)
    INV = 1.0 / XY_ONE
    q0 = (p0[0] << XY_SHIFT, p0[1] << XY_SHIFT)
    q1 = (p1[0] << XY_SHIFT, p1[1] << XY_SHIFT)
    dx, dy = (q0[0] - q1[0]) * INV, (q1[1] - q0[1]) * INV
    r = dx * dx + dy * dy
    odd = thickness & 1
    th = thickness << (XY_SHIFT - 1)
    if abs(r) > 2.220446049250313e-16:
        r = (th + odd * XY_ONE * 0.5) / math.sqrt(r)
        dpx, dpy = _cv_round(dy * r), _cv_round(dx * r)
        pt = [(q0[0] + dpx, q0[1] + dpy), (q0[0] - dpx, q0[1] - dpy), (q1[0] - dpx, q1[1] - dpy), (q1[0] + dpx, q1[1] + dpy)]
        fill_convex_poly(img, pt, color)
    rad = (th + (XY_ONE >> 1)) >> XY_SHIFT
    for i, q in enumerate((q0, q1)):
        if flags & (i + 1):
            circle_filled(img, ((q[0] + (XY_ONE >> 1)) >> XY_SHIFT, (q[1] + (XY_ONE >> 1)) >> XY_SHIFT), rad, color)


def line_aa(img: np.ndarray, p1, p2, color) -> None:
    """drawing.cpp LineAA on an 8-bit 3-channel image; p1, p2 are 16.16 fixed-point.  Three pixels per step across the
    minor axis, weights FILTER_TABLE x slope / end-point correction, every pixel blended twice with
    c += ((colour - c) * a + 127) >> 8."""
    H, W = img.shape[:2]
    ok, p1, p2 = clip_line(W << XY_SHIFT, H << XY_SHIFT, p1, p2)
    if not ok:
        return
    x1, y1 = p1
    x2, y2 = p2
    dx, dy = x2 - x1, y2 - y1
    ax, ay = abs(dx), abs(dy)
    col = [int(c) for c in color]

    def put(x, y, a):
        for c in range(3):
            t = int(img[y, x, c])
            t += ((col[c] - t) * a + 127) >> 8
            t += ((col[c] - t) * a + 127) >> 8
            img[y, x, c] = t & 255
    if ax > ay:
        if dx < 0:
            dy = -dy
            x1, x2, y1, y2 = x2, x1, y2, y1
        y_step = _trunc_div(dy << XY_SHIFT, ax | 1)
        x2 += XY_ONE
        ecount = (x2 >> XY_SHIFT) - (x1 >> XY_SHIFT)
        j = -(x1 & (XY_ONE - 1))
        y1 += ((y_step * j) >> XY_SHIFT) + (XY_ONE >> 1)
        slope = (y_step >> (XY_SHIFT - 5)) & 0x3f
        slope ^= 0x3f if y_step < 0 else 0
        i = (x1 >> (XY_SHIFT - 7)) & 0x78
        j = (x2 >> (XY_SHIFT - 7)) & 0x78
    else:
        if dy < 0:
            dx = -dx
            x1, x2, y1, y2 = x2, x1, y2, y1
        x_step = _trunc_div(dx << XY_SHIFT, ay | 1)
        y2 += XY_ONE
        ecount = (y2 >> XY_SHIFT) - (y1 >> XY_SHIFT)
        j = -(y1 & (XY_ONE - 1))
        x1 += ((x_step * j) >> XY_SHIFT) + (XY_ONE >> 1)
        slope = (x_step >> (XY_SHIFT - 5)) & 0x3f
        slope ^= 0x3f if x_step < 0 else 0
        i = (y1 >> (XY_SHIFT - 7)) & 0x78
        j = (y2 >> (XY_SHIFT - 7)) & 0x78
    slope = 0x100 if (slope & 0x20) else SLOPE_CORR_TABLE[slope]
    t0 = slope << 7
    t1 = ((0x78 - i) | 4) * slope
    t2 = (j | 4) * slope
    ep = [0] * 9
    ep[8] = slope
    ep[1] = ep[3] = (((((j - i) & 0x78) | 4) * slope) >> 8) & 0x1ff
    ep[2] = (t1 >> 8) & 0x1ff
    ep[4] = (((((j - i) + 0x80) | 4) * slope) >> 8) & 0x1ff
    ep[5] = ((t1 + t0) >> 8) & 0x1ff
    ep[6] = (t2 >> 8) & 0x1ff
    ep[7] = ((t2 + t0) >> 8) & 0x1ff
    scount = 0
    if ax > ay:
        x = x1 >> XY_SHIFT
        while ecount >= 0:
            if 0 <= x < W:
                y = (y1 >> XY_SHIFT) - 1
                ep_corr = ep[((((scount >= 2) + 1) & (scount | 2)) * 3) + (((ecount >= 2) + 1) & (ecount | 2))]
                dist = (y1 >> (XY_SHIFT - 5)) & 31
                for k, f in enumerate((FILTER_TABLE[dist + 32], FILTER_TABLE[dist], FILTER_TABLE[63 - dist])):
                    a = ((ep_corr * f) >> 8) & 0xff
                    if 0 <= y + k < H:
                        put(x, y + k, a)
            x += 1
            y1 += y_step
            scount += 1
            ecount -= 1
    else:
        y = y1 >> XY_SHIFT
        while ecount >= 0:
            if 0 <= y < H:
                x = (x1 >> XY_SHIFT) - 1
                ep_corr = ep[((((scount >= 2) + 1) & (scount | 2)) * 3) + (((ecount >= 2) + 1) & (ecount | 2))]
                dist = (x1 >> (XY_SHIFT - 5)) & 31
                for k, f in enumerate((FILTER_TABLE[dist + 32], FILTER_TABLE[dist], FILTER_TABLE[63 - dist])):
                    a = ((ep_corr * f) >> 8) & 0xff
                    if 0 <= x + k < W:
                        put(x + k, y, a)
            x1 += x_step
            y += 1
            scount += 1
            ecount -= 1


def line_aa_px(img, p0, p1, color):
    """cv2.line(..., thickness 1, LINE_AA) on integer pixel end points."""
    line_aa(img, (p0[0] << XY_SHIFT, p0[1] << XY_SHIFT), (p1[0] << XY_SHIFT, p1[1] << XY_SHIFT), color)


def polylines_closed(img, pts, color, thickness: int, aa: bool) -> None:
    """drawing.cpp PolyLine(is_closed = true): segment i joins pts[i-1] -> pts[i], starting with pts[n-1] -> pts[0]; thick
    segments carry a round cap at their second end point (flags = 2).  cv2.polylines / drawContours(thickness >= 0) /
    rectangle(thickness >= 0)."""
    n = len(pts)
    if n == 0:
        return
    p0 = tuple(int(t) for t in pts[n - 1])
    for i in range(n):
        p = tuple(int(t) for t in pts[i])
        if thickness <= 1:
            assert aa
            line_aa_px(img, p0, p, color)
        else:
            thick_line(img, p0, p, color, thickness, flags=2)
        p0 = p


def draw_marker_cross(img, pos, color, size: int, thickness: int) -> None:
    """cv2.drawMarker(MARKER_CROSS): the horizontal then the vertical bar, each cv2.line (round caps at both ends)."""
    h = size // 2
    thick_line(img, (pos[0] - h, pos[1]), (pos[0] + h, pos[1]), color, thickness, 3)
    thick_line(img, (pos[0], pos[1] - h), (pos[0], pos[1] + h), color, thickness, 3)


def rectangle2(img, x, y, w, h, color=(255, 0, 0)) -> None:
    """cv2.rectangle(vis, (x, y), (x + w, y + h), color, 2) (roi.py:44)."""
    polylines_closed(img, [(x, y), (x + w, y), (x + w, y + h), (x, y + h)], color, 2, aa=False)


def hull_in_cv_order(contour_pts: np.ndarray, hull_ccw: np.ndarray) -> np.ndarray:
    """cv2.convexHull's vertex order for a traced contour: counter-clockwise in (x, y) like `hull_ccw`
    (spec_mask.convex_hull_points), rotated so that the indices into the contour form a monotone sequence -- for contours
    of cv2.findContours that is: start at the hull vertex met LAST along the contour."""
    pl = [tuple(int(t) for t in p) for p in np.asarray(contour_pts).reshape(-1, 2)]
    hl = [tuple(int(t) for t in p) for p in np.asarray(hull_ccw).reshape(-1, 2)]
    if len(hl) < 3:
        return np.asarray(hl, np.int32).reshape(-1, 2)
    last = {}
    for i, p in enumerate(pl):
        last[p] = i
    k = max(range(len(hl)), key=lambda t: last[hl[t]])
    return np.asarray(hl[k:] + hl[:k], np.int32)


def analyze_overlay(rgb: np.ndarray, contour: np.ndarray, rec: dict, veins: np.ndarray) -> np.ndarray:
    """The overlay apply_analyze_filter returns (analyze.py:37-122) from the numeric record: rec has centroid, left, right,
    top, bottom, hull (cv2 order, [M,1,2] or [M,2]) and axes ((p0_min, p0_max), (p1_min, p1_max)); veins = Canny & mask."""
    out = rgb.copy()
    pts = np.asarray(contour).reshape(-1, 2)
    polylines_closed(out, pts, (255, 0, 0), 2, aa=False)                    # :40  drawContours
    c = tuple(int(t) for t in rec["centroid"])
    draw_marker_cross(out, c, (255, 255, 0), 14, 2)                         # :50-57
    for key in ("left", "right", "top", "bottom"):                          # :65-75
        p = tuple(int(t) for t in rec[key])
        circle_filled(out, p, 3, (255, 255, 0))
        line_aa_px(out, c, p, (255, 255, 0))
    hull = np.asarray(rec["hull"]).reshape(-1, 2)
    polylines_closed(out, hull, (0, 255, 0), 1, aa=True)                    # :78-85
    if rec.get("axes") is not None:                                         # :99-112
        (a0, a1), (b0, b1) = rec["axes"]
        thick_line(out, tuple(map(int, a0)), tuple(map(int, a1)), (255, 255, 0), 2, 3)
        thick_line(out, tuple(map(int, b0)), tuple(map(int, b1)), (255, 0, 255), 2, 3)
    out[veins] = (0, 255, 255)                                              # :115-122
    return out


def overlay_record(contour: np.ndarray) -> dict:
    """Everything analyze_overlay needs, from the contour alone (analyze.py:43-98): centroid / extreme points
    (spec_contour.analyze_record), hull in cv2.convexHull's order, PCA axis end points (float64 eigen-decomposition of the
    vertex covariance; cv2.PCACompute2 works in float32, the end points agree unless two projections tie)."""
    from . import spec_contour, spec_mask
    pts = np.asarray(contour).reshape(-1, 2)
    rec = dict(spec_contour.analyze_record(np.asarray(contour).reshape(-1, 1, 2)))
    rec["hull"] = hull_in_cv_order(pts, spec_mask.convex_hull_points(pts))
    rec["axes"] = None
    if len(pts) >= 2:
        d = pts.astype(np.float64)
        cov = np.cov(d.T, bias=True)
        w, v = np.linalg.eigh(cov)
        axes = []
        for k in (1, 0):   # major axis first
            proj = d @ v[:, k]
            axes.append((tuple(int(t) for t in pts[int(proj.argmin())]), tuple(int(t) for t in pts[int(proj.argmax())])))
        rec["axes"] = tuple(axes)
    return rec


def draw_primitives(img: np.ndarray, prims) -> np.ndarray:
    """The primitive list of lfx_draw_primitives (include/leafx.h) on ONE image, in place, through the functions above:
    rows (kind, x0, y0, x1, y1, r | g << 8 | b << 16, size, 0)."""
    for kind, x0, y0, x1, y1, c, size, _ in np.asarray(prims).reshape(-1, 8).tolist():
        col = (c & 255, (c >> 8) & 255, (c >> 16) & 255)
        if kind == 1 and size >= 2:
            thick_line(img, (x0, y0), (x1, y1), col, size, 3)
        elif kind == 2:
            line_aa_px(img, (x0, y0), (x1, y1), col)
        elif kind == 3 and size >= 0:
            circle_filled(img, (x0, y0), size, col)
        elif kind == 4 and size >= 2:
            polylines_closed(img, [(x0, y0), (x1, y0), (x1, y1), (x0, y1)], col, size, aa=False)
        elif kind == 5 and size >= 2:
            draw_marker_cross(img, (x0, y0), col, x1, size)
    return img
