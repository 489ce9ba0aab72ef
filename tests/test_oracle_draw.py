"""oracle/spec_draw.py (OpenCV's drawing.cpp restated: Line2, FillConvexPoly, ThickLine, Circle, LineAA, PolyLine) pinned
against cv2 itself and against the reference's own apply_analyze_filter / apply_roi_filter (analyze.py:37-122, roi.py:43-44).
CPU only."""
import numpy as np
import pytest

from leaffliction_b200 import synth
from oracle import spec_draw as sd

cv2 = pytest.importorskip("cv2")


def _rand_img(rng, H, W):
    return rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)


def _pt(rng, lo, hi):
    return int(rng.integers(lo, hi)), int(rng.integers(lo, hi))


def test_primitives_against_cv2():
    """Every primitive the overlays use, random end points inside the image and across its border, random background and
    colour: bit-identical to cv2.line / circle / drawMarker / rectangle."""
    rng = np.random.default_rng(5)
    H, W = 48, 64
    for t in range(300):
        img = _rand_img(rng, H, W)
        lo, hi = (-10, 74) if t % 2 else (0, 48)
        p0, p1 = _pt(rng, lo, hi), _pt(rng, lo, hi)
        if t % 7 == 0:
            p1 = p0
        col = tuple(int(c) for c in rng.integers(0, 256, 3))
        for th in (2, 3):
            a = img.copy(); cv2.line(a, p0, p1, col, th)
            b = img.copy(); sd.thick_line(b, p0, p1, col, th, 3)
            assert np.array_equal(a, b), ("thick", th, p0, p1)
        a = img.copy(); cv2.line(a, p0, p1, col, 1, cv2.LINE_AA)
        b = img.copy(); sd.line_aa_px(b, p0, p1, col)
        assert np.array_equal(a, b), ("aa", p0, p1)
        a = img.copy(); cv2.circle(a, p0, 3, col, -1)
        b = img.copy(); sd.circle_filled(b, p0, 3, col)
        assert np.array_equal(a, b), ("circle", p0)
        a = img.copy(); cv2.drawMarker(a, p0, col, markerType=cv2.MARKER_CROSS, markerSize=14, thickness=2)
        b = img.copy(); sd.draw_marker_cross(b, p0, col, 14, 2)
        assert np.array_equal(a, b), ("marker", p0)
        x, y = int(rng.integers(0, W - 2)), int(rng.integers(0, H - 2))
        w, h = int(rng.integers(1, W - x + 1)), int(rng.integers(1, H - y + 1))
        a = img.copy(); cv2.rectangle(a, (x, y), (x + w, y + h), (255, 0, 0), 2)
        b = img.copy(); sd.rectangle2(b, x, y, w, h)
        assert np.array_equal(a, b), ("rect", x, y, w, h)


def test_fixed_point_entry_points_against_cv2():
    """FillConvexPoly and LineAA on 16.16 vertices (cv2's shift = 16), including vertices outside the image."""
    rng = np.random.default_rng(2)
    H, W = 48, 64
    for t in range(200):
        lo, hi = (-10 << 16, 74 << 16) if t % 2 else (0, 48 << 16)
        v = [_pt(rng, lo, hi) for _ in range(3)]
        a = np.zeros((H, W, 3), np.uint8); cv2.fillConvexPoly(a, np.array(v, np.int32), (255, 255, 255), cv2.LINE_8, shift=16)
        b = np.zeros((H, W, 3), np.uint8); sd.fill_convex_poly(b, v, (255, 255, 255))
        assert np.array_equal(a, b), v
        img = _rand_img(rng, H, W)
        a = img.copy(); cv2.line(a, v[0], v[1], (10, 200, 90), 1, cv2.LINE_AA, shift=16)
        b = img.copy(); sd.line_aa(b, v[0], v[1], (10, 200, 90))
        assert np.array_equal(a, b), v[:2]


def _leaf_contours(n, size, seed):
    out = []
    for img in synth.leaf_batch(n, size, size, seed):
        hsv = cv2.cvtColor(img, cv2.COLOR_RGB2HSV)
        m = cv2.inRange(hsv, (25, 40, 20), (95, 255, 255))
        cs, _ = cv2.findContours(m, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        if cs:
            c = max(cs, key=cv2.contourArea)
            full = np.zeros_like(m)
            cv2.drawContours(full, [c], -1, 255, -1)
            out.append((img, full, c))
    return out


def test_polylines_and_hull_order_against_cv2():
    """drawContours(thickness 2) and the anti-aliased hull polyline on real contours; hull_in_cv_order reproduces
    cv2.convexHull's vertex order (the blend of overlapping anti-aliased segments depends on it)."""
    from oracle import spec_mask
    seen = 0
    for img, _, c in _leaf_contours(6, 96, 11) + _leaf_contours(2, 256, 3):
        pts = c[:, 0, :]
        a = img.copy(); cv2.drawContours(a, [c], -1, (255, 0, 0), 2)
        b = img.copy(); sd.polylines_closed(b, pts, (255, 0, 0), 2, aa=False)
        assert np.array_equal(a, b)
        hv = cv2.convexHull(c)
        mine = sd.hull_in_cv_order(pts, spec_mask.convex_hull_points(pts))
        assert np.array_equal(hv[:, 0, :], mine)
        a = img.copy(); cv2.polylines(a, [hv], isClosed=True, color=(0, 255, 0), thickness=1, lineType=cv2.LINE_AA)
        b = img.copy(); sd.polylines_closed(b, mine, (0, 255, 0), 1, aa=True)
        assert np.array_equal(a, b)
        seen += 1
    assert seen >= 6


def _cv_overlay(rgb, mask, contour):
    """The drawing calls of apply_analyze_filter (analyze.py:37-122), restated for boxes without /root/reference."""
    overlay = rgb.copy()
    cv2.drawContours(overlay, [contour], -1, (255, 0, 0), 2)
    M = cv2.moments(contour)
    cx, cy = int(M["m10"] / M["m00"]), int(M["m01"] / M["m00"])
    cv2.drawMarker(overlay, (cx, cy), (255, 255, 0), markerType=cv2.MARKER_CROSS, markerSize=14, thickness=2)
    pts = contour[:, 0, :]
    for p in (pts[pts[:, 0].argmin()], pts[pts[:, 0].argmax()], pts[pts[:, 1].argmin()], pts[pts[:, 1].argmax()]):
        cv2.circle(overlay, (int(p[0]), int(p[1])), 3, (255, 255, 0), -1)
        cv2.line(overlay, (cx, cy), (int(p[0]), int(p[1])), (255, 255, 0), 1, lineType=cv2.LINE_AA)
    cv2.polylines(overlay, [cv2.convexHull(contour)], isClosed=True, color=(0, 255, 0), thickness=1, lineType=cv2.LINE_AA)
    data = pts.astype(np.float32)
    _, ev, _ = cv2.PCACompute2(data, mean=None)
    for k, colr in ((0, (255, 255, 0)), (1, (255, 0, 255))):
        pr = data @ ev[k]
        a, b = data[int(pr.argmin())], data[int(pr.argmax())]
        cv2.line(overlay, (int(a[0]), int(a[1])), (int(b[0]), int(b[1])), colr, 2)
    gray = cv2.cvtColor(rgb, cv2.COLOR_RGB2GRAY)
    edges = cv2.Canny(gray, threshold1=80, threshold2=160, L2gradient=True)
    overlay[(edges > 0) & (mask > 0)] = (0, 255, 255)
    return overlay


def test_analyze_overlay_against_cv2_calls():
    from oracle import spec_color, spec_filters
    for img, mask, c in _leaf_contours(4, 96, 21) + _leaf_contours(1, 256, 8):
        veins = (spec_filters.canny(spec_color.rgb_to_gray(img), 80, 160, True) > 0) & (mask > 0)
        got = sd.analyze_overlay(img, c, sd.overlay_record(c), veins)
        assert np.array_equal(got, _cv_overlay(img, mask, c))


@pytest.mark.needs_reference
def test_overlays_against_the_reference_functions():
    """The reference's own apply_analyze_filter / apply_roi_filter (imported from /root/reference) on seeded leaves."""
    import ref_harness
    from oracle import spec_color, spec_filters
    ns = ref_harness.load()
    cfg = ref_harness.ref_config(ns)
    for img, mask, c in _leaf_contours(4, 96, 33) + _leaf_contours(1, 256, 9):
        ref = ns.analyze.apply_analyze_filter(img, mask, c, cfg)
        veins = (spec_filters.canny(spec_color.rgb_to_gray(img), 80, 160, True) > 0) & (mask > 0)
        assert np.array_equal(sd.analyze_overlay(img, c, sd.overlay_record(c), veins), ref)
        _, vis, box = ns.roi.apply_roi_filter(img, c, cfg)
        mine = img.copy()
        sd.rectangle2(mine, *box)
        assert np.array_equal(mine, vis)


def _golden_cases():
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_draw_v1.npz"))
    keys = sorted({k.rsplit("_rgb", 1)[0] for k in g.files if k.endswith("_rgb")})
    return g, keys


def test_oracle_against_golden_overlays():
    """tests/golden/golden_draw_v1.npz holds the outputs of the reference's own apply_analyze_filter / apply_roi_filter
    (make_golden_draw.py); the oracle reproduces them from the stored inputs, without OpenCV."""
    g, keys = _golden_cases()
    assert len(keys) >= 5
    for k in keys:
        rgb, mask, c = g[k + "_rgb"], g[k + "_mask"], g[k + "_contour"]
        veins = (g[k + "_edges"] > 0) & (mask > 0)
        assert np.array_equal(sd.analyze_overlay(rgb, c, sd.overlay_record(c), veins), g[k + "_analyze"]), k
        vis = rgb.copy()
        sd.rectangle2(vis, *[int(v) for v in g[k + "_roi_box"]])
        assert np.array_equal(vis, g[k + "_roi_vis"]), k


def test_primitive_list_against_cv2():
    """oracle.spec_draw.draw_primitives (the list format of lfx_draw_primitives) == the cv2 calls it stands for, in order."""
    from test_draw_hostsim_cpu import random_primitives
    rng = np.random.default_rng(4)
    B, P, H, W = 20, 8, 40, 56
    imgs = rng.integers(0, 256, size=(B, H, W, 3), dtype=np.uint8)
    prims = random_primitives(rng, B, P, H, W)
    for b in range(B):
        exp = imgs[b].copy()
        for kind, x0, y0, x1, y1, c, size, _ in prims[b].tolist():
            col = (c & 255, (c >> 8) & 255, (c >> 16) & 255)
            if kind == 1:
                cv2.line(exp, (x0, y0), (x1, y1), col, size)
            elif kind == 2:
                cv2.line(exp, (x0, y0), (x1, y1), col, 1, cv2.LINE_AA)
            elif kind == 3:
                cv2.circle(exp, (x0, y0), size, col, -1)
            elif kind == 4:
                cv2.rectangle(exp, (x0, y0), (x1, y1), col, size)
            elif kind == 5:
                cv2.drawMarker(exp, (x0, y0), col, markerType=cv2.MARKER_CROSS, markerSize=x1, thickness=size)
        assert np.array_equal(sd.draw_primitives(imgs[b].copy(), prims[b]), exp), b
