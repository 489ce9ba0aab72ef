"""NumPy spec of the leaf-mask path: threshold strategies, post-processing, brown extension,
make_mask orchestration, apply_mask, ROI letterbox, colour statistics.

Restates srcs/transform/filters/mask.py (make_mask :548-582, _postprocess_mask :53-69,
_create_hsv_masks :72-98, _create_lab_mask :101-106, _create_fallback_mask :395-411,
_extend_mask_with_brown_regions :335-392, _create_inclusive_mask :727-831,
_create_enhanced_mask :610-724), srcs/cli/Transformation.py:285-299 (largest_contour,
contour_to_mask), srcs/utils/mask_utils.py:10-83 (apply_mask), srcs/transform/filters/roi.py:20-46,
hist.py:22-67,188,248-256, brown.py:21-89, blur.py:18-79 -- all under /root/reference --
without contour tracing: findContours(RETR_EXTERNAL) -> max(contourArea) -> drawContours(filled)
is restated with connected-component counts.
PlantCV (>=3.14, not installable here) semantics restated: fill = skimage remove_small_objects
(connectivity 1, strict '<'), rgb2gray_hsv = HSV channel, threshold.otsu = cv2 Otsu.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Tuple

import numpy as np

from . import spec_color as sc
from . import spec_filters as sf


@dataclass(frozen=True)
class Cfg:
    """The TransformConfig fields the numeric path reads (Transformation.py:63-93), with the
    values of srcs/transform/config.yaml and parity profile P0 (grabcut off, no upscale)."""
    gaussian_sigma: float = 1.5
    hsv_channel_for_mask: str = "s"
    fill_size: int = 1000
    morph_kernel: int = 3
    roi_size: Tuple[int, int] = (256, 256)
    mask_strategy: str = "inclusive"
    bg_bias: Optional[str] = "light_bg"
    green_hue_range: Tuple[int, int] = (25, 100)
    min_object_area_ratio: float = 0.10
    max_object_area_ratio: float = 0.98
    brown_hue_range: Tuple[int, int] = (0, 30)
    brown_s_min: int = 20
    brown_v_max: int = 200
    brown_min_area_px: int = 25
    brown_morph_kernel: int = 3
    use_lab_brown: bool = False
    lab_b_min: int = 125
    lab_a_min: int = 125


# ------------------------------------------------------------ components
def remove_small_objects4(mask: np.ndarray, size: int) -> np.ndarray:
    """pcv.fill(bin_img, size): drop 4-connected foreground components with < size pixels."""
    b = mask > 0
    lab = sf.label4(b)
    sizes = np.bincount(lab.ravel())
    small = sizes < size
    small[0] = False
    return (b & ~small[lab]).astype(np.uint8) * 255


def fill_holes(mask: np.ndarray) -> np.ndarray:
    """Foreground plus every background region not 4-connected to the image border."""
    b = mask > 0
    H, W = b.shape
    p = np.zeros((H + 2, W + 2), bool)
    p[1:-1, 1:-1] = b
    lab = sf.label4(~p)
    outside = lab == lab[0, 0]
    return (~outside[1:-1, 1:-1]).astype(np.uint8) * 255


def component_stats8(mask: np.ndarray):
    """8-connected components with the counts that determine cv2.contourArea of the external
    contour of a hole-free component: N pixels, P crack edges, Q1 2x2 windows with one pixel.

    Returns (labels, dict of arrays indexed by label 1..n): n, area2 (= 2*contourArea), bbox,
    first (raster index of first pixel)."""
    b = mask > 0
    H, W = b.shape
    lab = sf.label8(b)
    n = int(lab.max())
    p = np.zeros((H + 2, W + 2), bool)
    p[1:-1, 1:-1] = b
    c = p[1:-1, 1:-1]
    up, dn, lf, rt = p[:-2, 1:-1], p[2:, 1:-1], p[1:-1, :-2], p[1:-1, 2:]
    ul, ur, dl, dr = p[:-2, :-2], p[:-2, 2:], p[2:, :-2], p[2:, 2:]
    edges = (~up).astype(np.int64) + (~dn) + (~lf) + (~rt)
    q1 = ((~rt & ~dn & ~dr).astype(np.int64) + (~lf & ~dn & ~dl) + (~rt & ~up & ~ur) + (~lf & ~up & ~ul))
    flat = lab.ravel()
    N = np.bincount(flat, minlength=n + 1)
    P = np.bincount(flat, weights=(edges * c).ravel(), minlength=n + 1).astype(np.int64)
    Q = np.bincount(flat, weights=(q1 * c).ravel(), minlength=n + 1).astype(np.int64)
    area2 = 2 * N - (P - Q) - 2          # 2 * contourArea
    ys, xs = np.nonzero(b)
    l = lab[ys, xs]
    x0 = np.full(n + 1, W); x1 = np.full(n + 1, -1); y0 = np.full(n + 1, H); y1 = np.full(n + 1, -1)
    np.minimum.at(x0, l, xs); np.maximum.at(x1, l, xs); np.minimum.at(y0, l, ys); np.maximum.at(y1, l, ys)
    first = np.full(n + 1, H * W)
    np.minimum.at(first, l, ys * W + xs)
    return lab, dict(n=N, area2=area2, x0=x0, y0=y0, x1=x1, y1=y1, first=first, count=n)


def largest_external(mask: np.ndarray):
    """largest_contour + contour_to_mask (Transformation.py:285-299) without tracing.

    Returns (filled_mask u8, info) where info = dict(bbox=(x,y,w,h), area2, label_mask) or
    (mask, None) when there is no foreground.  Ties on contourArea go to the component whose
    first raster pixel comes LAST (findContours returns contours in reverse discovery order and
    Python's max() keeps the first maximum)."""
    filled = fill_holes(mask)
    lab, st = component_stats8(filled)
    if st["count"] == 0:
        return np.zeros_like(mask), None
    ids = np.arange(1, st["count"] + 1)
    a2 = st["area2"][1:]
    best = ids[a2 == a2.max()]
    pick = best[np.argmax(st["first"][best])]
    out = (lab == pick).astype(np.uint8) * 255
    bbox = (int(st["x0"][pick]), int(st["y0"][pick]),
            int(st["x1"][pick] - st["x0"][pick] + 1), int(st["y1"][pick] - st["y0"][pick] + 1))
    return out, dict(bbox=bbox, area2=int(a2.max()), npix=int(st["n"][pick]))


def filter_components8(mask: np.ndarray, min_area: int):
    """connectedComponentsWithStats(8) + keep area >= min_area (mask.py:373-380, brown.py:62-74).
    Returns (mask, kept_count, kept_pixels)."""
    b = mask > 0
    lab = sf.label8(b)
    sizes = np.bincount(lab.ravel())
    keep = sizes >= min_area
    keep[0] = False
    return (keep[lab]).astype(np.uint8) * 255, int(keep.sum()), int(sizes[keep].sum())


def keep_largest8(mask: np.ndarray) -> np.ndarray:
    """mask.py:711-718, :818-825: keep the 8-connected component with most pixels
    (np.argmax -> first label in raster order on ties); unchanged when there is no foreground."""
    b = mask > 0
    lab = sf.label8(b)
    if lab.max() < 1:
        return b.astype(np.uint8) * 255
    sizes = np.bincount(lab.ravel())[1:]
    return (lab == 1 + int(np.argmax(sizes))).astype(np.uint8) * 255


# ------------------------------------------------------------ strategies
def mask_hsv_green(rgb, cfg: Cfg):
    hsv = sc.rgb_to_hsv(rgb)
    h, s = hsv[..., 0], hsv[..., 1]
    lo, hi = cfg.green_hue_range
    return ((h >= lo) & (h <= hi) & (s >= 40)).astype(np.uint8) * 255          # mask.py:86-91


def mask_lab(rgb):
    lab = sc.rgb_to_lab(rgb)
    a, b = lab[..., 1], lab[..., 2]
    return ((a <= 135) & (b >= 115) & (b <= 170)).astype(np.uint8) * 255      # mask.py:101-106


def mask_hsv_otsu(rgb, channel: str, object_type: str):
    """pcv.rgb2gray_hsv + pcv.threshold.otsu (mask.py:76-84); S and V are R/B-symmetric."""
    hsv = sc.rgb_to_hsv(rgb)
    g = hsv[..., "hsv".index(channel)]
    if channel == "h":  # PlantCV treats the array as BGR: hue differs, S and V do not
        g = sc.rgb_to_hsv(rgb[..., ::-1])[..., 0]
    return sf.otsu_binary(g, object_type)


def brown_predicate(rgb, cfg: Cfg):
    """mask.py:353-364 / brown.py:34-49 / blur.py:44-53 without the spatial constraint."""
    if cfg.use_lab_brown:
        lab = sc.rgb_to_lab(rgb)
        return (lab[..., 1] >= cfg.lab_a_min) & (lab[..., 2] >= cfg.lab_b_min)
    hsv = sc.rgb_to_hsv(rgb)
    lo, hi = cfg.brown_hue_range
    return ((hsv[..., 0] >= lo) & (hsv[..., 0] <= hi) & (hsv[..., 1] >= cfg.brown_s_min)
            & (hsv[..., 2] <= cfg.brown_v_max))


def mask_inclusive(rgb, cfg: Cfg):
    """_create_inclusive_mask (mask.py:727-831), including the uint8 wrap of `r + 15`."""
    hsv = sc.rgb_to_hsv(rgb)
    lab = sc.rgb_to_lab(rgb)
    h, s, v = (hsv[..., i].astype(np.int32) for i in range(3))
    L, a, b = (lab[..., i].astype(np.int32) for i in range(3))
    r, g, bl = (rgb[..., i].astype(np.int32) for i in range(3))
    lo, hi = cfg.green_hue_range
    elo, ehi = max(0, lo - 10), min(179, hi + 15)
    strong_green = (h >= elo) & (h <= ehi) & (s >= 30) & (v >= 30)
    w8 = lambda x, k: (x + k) & 255                                   # uint8 wrap (:753-757)
    green_dom = (g > w8(r, 15)) | (g > w8(bl, 15)) | ((g > w8(r, 5)) & (g > w8(bl, 5)) & (s >= 20))
    lab_green = (a <= 125) & (b >= 120) & (L >= 20) & (L <= 240)
    gray = sc.rgb_to_gray(rgb)
    blur = sf.gaussian_blur_u8(gray, 15, 0)
    tdiff = np.abs(gray.astype(np.int32) - blur.astype(np.int32))
    bg = (((s <= 25) & (v >= 50) & (v <= 220))
          | ((h >= 120) & (h <= 160) & (s >= 20) & (r > g) & (bl > g))
          | ((s <= 15) & (tdiff < 10)))
    edges = sf.canny(gray, 30, 100, False)
    dil = sf.dilate(edges, sf.ellipse_footprint(3))
    cand = strong_green | green_dom | lab_green | (dil > 0)
    m = (cand & ~bg).astype(np.uint8) * 255
    m = sf.morph_open(m, sf.ellipse_footprint(3))
    m = sf.morph_close(m, sf.ellipse_footprint(9))
    m = sf.morph_close(m, sf.ellipse_footprint(7))
    m = keep_largest8(m)
    m = sf.morph_close(m, sf.ellipse_footprint(5))
    return m


def mask_enhanced(rgb, cfg: Cfg):
    """_create_enhanced_mask (mask.py:610-724)."""
    hsv = sc.rgb_to_hsv(rgb)
    lab = sc.rgb_to_lab(rgb)
    h, s, v = (hsv[..., i].astype(np.int32) for i in range(3))
    L, a, b = (lab[..., i].astype(np.int32) for i in range(3))
    lo, hi = cfg.green_hue_range
    veg_hsv = (h >= lo) & (h <= hi) & (s >= 25) & (v >= 20) & (v <= 240)
    veg_lab = (a <= 135) & (b >= 105) & (L >= 30) & (L <= 220)
    if cfg.use_lab_brown:
        brown = (a >= cfg.lab_a_min - 10) & (b >= cfg.lab_b_min - 10) & (L >= 20)
    else:
        blo, bhi = cfg.brown_hue_range
        brown = ((((h >= blo) & (h <= bhi + 20)) | ((h >= 160) & (h <= 180)))
                 & (s >= cfg.brown_s_min - 10) & (v <= cfg.brown_v_max + 30))
    gray = sc.rgb_to_gray(rgb)
    edges = sf.canny(gray, 30, 100) | sf.canny(gray, 50, 150)
    edge_regions = sf.dilate(edges, sf.ellipse_footprint(5), 2)
    # vegetation(0/1) + 0.3*edge(0/1) > 0.3  <=>  vegetation  (0.3f > 0.3 is False in float32)
    veg = (veg_hsv | veg_lab | brown).astype(np.float32)
    enh = veg + (edge_regions.astype(np.float32) / np.float32(255.0)) * np.float32(0.3)
    m = (enh > 0.3).astype(np.uint8) * 255
    m = sf.morph_close(m, sf.ellipse_footprint(7))
    m = sf.morph_open(m, sf.ellipse_footprint(3))
    m = sf.morph_close(m, sf.ellipse_footprint(9))
    m = keep_largest8(m)
    m = sf.morph_close(m, sf.ellipse_footprint(3))
    return m


def raw_candidate(rgb, cfg: Cfg):
    """_build_mask_candidates for the single-strategy settings (mask.py:414-434)."""
    bias = (cfg.bg_bias or "auto").lower()
    st = cfg.mask_strategy
    if st == "hsv_s":
        return mask_hsv_otsu(rgb, "s", "light" if bias != "dark_bg" else "dark")
    if st == "hsv_v_dark":
        return mask_hsv_otsu(rgb, "v", "dark")
    if st == "hsv_h":
        return mask_hsv_green(rgb, cfg)
    if st == "lab":
        return mask_lab(rgb)
    if st == "enhanced":
        return mask_enhanced(rgb, cfg)
    if st == "inclusive":
        return mask_inclusive(rgb, cfg)
    if st == "kmeans":
        from . import spec_kmeans
        return spec_kmeans.kmeans_mask(rgb, cfg)
    raise ValueError(f"strategy {st!r} is outside the bit-exact contract (SURVEY.md section 8a tier C)")


# ------------------------------------------------------------ post-process / make_mask
def postprocess(raw: np.ndarray, cfg: Cfg):
    """_postprocess_mask (mask.py:53-69). Returns (mask, info|None); info None <=> cnt is None."""
    b = (raw > 0).astype(np.uint8) * 255
    filled = remove_small_objects4(b, cfg.fill_size)
    fp = sf.ellipse_footprint(cfg.morph_kernel)
    opened = sf.morph_open(sf.morph_close(filled, fp), fp)
    m, info = largest_external(opened)
    if info is None:
        return opened, None
    return m, info


def extend_with_brown(best: np.ndarray, rgb: np.ndarray, cfg: Cfg):
    """_extend_mask_with_brown_regions (mask.py:335-392): returns the UNFILLED union and the
    stats of its largest external contour."""
    search = sf.dilate(best, sf.ellipse_footprint(20), 2) > 0
    brown = brown_predicate(rgb, cfg) & search
    fp = sf.ellipse_footprint(cfg.brown_morph_kernel)
    clean = sf.morph_close(sf.morph_open(brown.astype(np.uint8) * 255, fp), fp)
    filt, _, _ = filter_components8(clean, cfg.brown_min_area_px)
    ext = ((best > 0) | (filt > 0)).astype(np.uint8) * 255
    _, info = largest_external(ext)
    if info is None:
        return best, None
    return ext, info


def make_mask(rgb: np.ndarray, cfg: Cfg):
    """make_mask (mask.py:548-582) under parity profile P0/P1 (no upscale, no grabCut, no shadow).

    Returns (mask u8 [H,W], info) with info = dict(bbox=(x,y,w,h) of the returned contour,
    area2 = 2*contourArea) or None when the reference returns contour None."""
    raw = raw_candidate(rgb, cfg)
    m, info = postprocess(raw, cfg)
    # _find_best_mask/_score_mask (:446-461,:143-152): a single candidate is rejected only when
    # cnt is None or contourArea <= 1 (score -1.0 is not > -1.0).
    if info is None or info["area2"] <= 2:
        # _create_fallback_mask (:395-411): Otsu 'light' on hsv_channel_for_mask
        fb = mask_hsv_otsu(rgb, cfg.hsv_channel_for_mask, "light")
        m, info = postprocess(fb, cfg)
    return extend_with_brown(m, rgb, cfg)


# ------------------------------------------------------------ _score_mask / "auto" (mask.py:143-188, :435-461)
def first_pixel(mask: np.ndarray):
    """(x, y) of the first raster pixel of a non-empty mask: where findContours starts the outer border."""
    idx = int(np.flatnonzero(mask.ravel() > 0)[0])
    return idx % mask.shape[1], idx // mask.shape[1]


def convex_hull_points(pts: np.ndarray) -> np.ndarray:
    """Vertices of the convex hull of integer points (cv2.convexHull, mask.py:158, analyze.py:77): Andrew's monotone
    chain, exact integer arithmetic, collinear points dropped.  int64 [M,2], counter-clockwise in image coordinates."""
    q = np.unique(np.asarray(pts, np.int64).reshape(-1, 2), axis=0)
    if len(q) <= 2:
        return q

    def half(seq):
        out = []
        for x, y in seq:
            while len(out) >= 2 and (out[-1][0] - out[-2][0]) * (y - out[-2][1]) - (out[-1][1] - out[-2][1]) * (x - out[-2][0]) <= 0:
                out.pop()
            out.append((int(x), int(y)))
        return out
    lower, upper = half(q), half(q[::-1])
    return np.array(lower[:-1] + upper[:-1], np.int64)


def polygon_area2(pts: np.ndarray) -> int:
    """2 * cv2.contourArea of an integer polygon (exact)."""
    p = np.asarray(pts, np.int64).reshape(-1, 2)
    if len(p) < 3:
        return 0
    q = np.roll(p, -1, axis=0)
    return abs(int((p[:, 0] * q[:, 1] - q[:, 0] * p[:, 1]).sum()))


def score_features(mask: np.ndarray, rgb: np.ndarray, cfg: Cfg):
    """The image-dependent terms of _score_mask (mask.py:162-177): boundary strength and green fraction."""
    gray = sc.rgb_to_gray(rgb)
    mag = sf.normalize_minmax_f32(sf.sobel_magnitude_f32(gray), 0.0, 1.0)          # :162-165
    fp3 = sf.ellipse_footprint(3)
    boundary = (sf.dilate(mask, fp3) > 0) ^ (sf.erode(mask, fp3) > 0)               # :166-169
    b_strength = float(mag[boundary].mean()) if boundary.sum() > 0 else 0.0         # :170
    hsv = sc.rgb_to_hsv(rgb)
    lo, hi = cfg.green_hue_range
    green = (hsv[..., 0] >= lo) & (hsv[..., 0] <= hi) & (hsv[..., 1] >= 40)         # :172-175
    denom = max(1, int(mask.astype(np.int64).sum() // 255))
    green_frac = float((green & (mask > 0)).sum()) / float(denom)                   # :176-177
    return b_strength, green_frac


def score_mask(mask: np.ndarray, info, rgb: np.ndarray, cfg: Cfg) -> float:
    """_score_mask (mask.py:143-188) for a post-processed candidate (`info` None <=> cnt is None)."""
    from . import spec_contour
    if info is None:
        return -1.0
    h, w = mask.shape[:2]
    area = info["area2"] / 2.0
    if area <= 1:
        return -1.0
    area_ratio = area / float(h * w)
    if area_ratio < cfg.min_object_area_ratio or area_ratio > cfg.max_object_area_ratio:
        return 0.01
    cnt = spec_contour.trace_external(mask, first_pixel(mask))
    hull_area = polygon_area2(convex_hull_points(cnt[:, 0, :])) / 2.0
    solidity = (area / hull_area) if hull_area > 1 else 0.0
    b_strength, green_frac = score_features(mask, rgb, cfg)
    x, y, ww, hh = info["bbox"]
    touches = (x <= 0) or (y <= 0) or (x + ww >= w - 1) or (y + hh >= h - 1)
    target = 0.35
    area_term = max(0.0, 1.0 - abs(area_ratio - target) / target)
    score = 0.35 * area_term + 0.25 * solidity + 0.25 * b_strength + 0.15 * green_frac
    if touches:
        score *= 0.75
    return float(score)


AUTO_CANDIDATES = ("hsv_s", "hsv_v_dark", "hsv_h", "lab", "kmeans", "enhanced", "inclusive")


def make_mask_auto(rgb: np.ndarray, cfg: Cfg, return_choice: bool = False, with_kmeans: bool = True):
    """make_mask with mask_strategy "auto" (mask.py:435-441): the seven candidates in the reference's order (the k-means
    one -- spec_kmeans, cv2.kmeans restated -- can be left out: `with_kmeans=False`), _find_best_mask's strict-greater
    selection (:446-461), then the fallback and the brown extension exactly as for a single strategy."""
    import dataclasses
    best, best_info, best_score, choice = None, None, -1.0, None
    for st in AUTO_CANDIDATES:
        if st == "kmeans" and not with_kmeans:
            continue
        raw = raw_candidate(rgb, dataclasses.replace(cfg, mask_strategy=st))
        m, info = postprocess(raw, cfg)
        sc_ = score_mask(m, info, rgb, cfg)
        if sc_ > best_score:
            best, best_info, best_score, choice = m, info, sc_, st
    if best is None:
        fb = mask_hsv_otsu(rgb, cfg.hsv_channel_for_mask, "light")
        best, best_info = postprocess(fb, cfg)
    out = extend_with_brown(best, rgb, cfg)
    return (out + (choice, best_score)) if return_choice else out


def apply_mask(img: np.ndarray, mask: np.ndarray, color: str = "white") -> np.ndarray:
    """mask_utils.py:10-83: binarise at >127, paint the rest 255 or 0."""
    val = 255 if color.lower() == "white" else 0
    out = img.copy()
    out[~(mask > 127)] = val
    return out


# ------------------------------------------------------------ ROI letterbox
INTER_RESIZE_COEF_BITS = 11


def _area_up_taps(src: int, dst: int):
    """cv::resize INTER_AREA when dst > src: 2-tap linear with area-mode offsets (resize.cpp)."""
    inv = np.float64(dst) / np.float64(src)  # inv_scale_x = (double)dsize.width / ssize.width
    scale = np.float64(1.0) / inv            # scale_x = 1. / inv_scale_x  (NOT src/dst: differs by an ulp)
    ofs = np.zeros(dst, np.int64)
    w = np.zeros((dst, 2), np.int64)
    for d in range(dst):
        sx = int(np.floor(d * scale))
        fx = np.float32((d + 1) - (sx + 1) * inv)
        fx = np.float32(0.0) if fx <= 0 else fx - np.floor(fx)
        if sx < 0:
            fx, sx = np.float32(0), 0
        if sx >= src - 1:
            fx, sx = np.float32(0), src - 1
        ofs[d] = sx
        c0 = np.float32(1.0) - np.float32(fx)
        w[d, 0] = int(np.rint(np.float32(c0 * np.float32(1 << INTER_RESIZE_COEF_BITS))))
        w[d, 1] = int(np.rint(np.float32(np.float32(fx) * np.float32(1 << INTER_RESIZE_COEF_BITS))))
    return ofs, w


def resize_area_up(img: np.ndarray, nw: int, nh: int) -> np.ndarray:
    """cv2.resize(img,(nw,nh),INTER_AREA) for nw >= w and nh >= h (8-bit fixed-point bilinear)."""
    h, w = img.shape[:2]
    if (nw, nh) == (w, h):
        return img.copy()
    xo, xa = _area_up_taps(w, nw)
    yo, ya = _area_up_taps(h, nh)
    s = img.astype(np.int64)
    x1 = np.minimum(xo + 1, w - 1)
    hrow = s[:, xo] * xa[:, 0][None, :, None] + s[:, x1] * xa[:, 1][None, :, None]
    y1 = np.minimum(yo + 1, h - 1)
    b0 = ya[:, 0][:, None, None]
    b1 = ya[:, 1][:, None, None]
    out = (((b0 * (hrow[yo] >> 4)) >> 16) + ((b1 * (hrow[y1] >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def _area_down_tab(ssize: int, dsize: int):
    """cv::computeResizeAreaTab (OpenCV resize.cpp): (dst index, src index, float32 weight) triples of the area resampling
    of `ssize` samples down to `dsize`, computed in float64 and rounded to float32 like the library."""
    import math
    scale = ssize / dsize
    tab = []
    for dx in range(dsize):
        fsx1 = dx * scale
        fsx2 = fsx1 + scale
        cell = min(scale, ssize - fsx1)
        sx1, sx2 = math.ceil(fsx1), math.floor(fsx2)
        sx2 = min(sx2, ssize - 1)
        sx1 = min(sx1, sx2)
        if sx1 - fsx1 > 1e-3:
            tab.append((dx, sx1 - 1, np.float32((sx1 - fsx1) / cell)))
        for sx in range(sx1, sx2):
            tab.append((dx, sx, np.float32(1.0 / cell)))
        if fsx2 - sx2 > 1e-3:
            tab.append((dx, sx2, np.float32(min(min(fsx2 - sx2, 1.0), cell) / cell)))
    return tab


def resize_area_down(img: np.ndarray, nw: int, nh: int) -> np.ndarray:
    """cv2.resize(img,(nw,nh),INTER_AREA) for nw <= w and nh <= h (8-bit): integer factors in both directions take
    ResizeAreaFast (integer block sums; 2x2 rounds as (s + 2) >> 2, other factors as cvRound(s * float32(1/area))), every
    other ratio takes ResizeArea_ (float32 accumulation in table order, cvRound at the end).  Bit-exact against OpenCV
    4.13 on random images (tests/test_oracle_golden.py::test_resize_area_down_against_opencv)."""
    h, w, C = img.shape
    fx, fy = w / nw, h / nh
    ix, iy = int(fx + 0.5), int(fy + 0.5)
    eps = 2.220446049250313e-16
    if abs(fx - ix) < eps and abs(fy - iy) < eps:
        blk = img[:nh * iy, :nw * ix].reshape(nh, iy, nw, ix, C).astype(np.int64).sum(axis=(1, 3))
        if ix == 2 and iy == 2:
            return ((blk + 2) >> 2).astype(np.uint8)
        v = blk.astype(np.float32) * (np.float32(1.0) / np.float32(ix * iy))
        return np.clip(np.rint(v), 0, 255).astype(np.uint8)
    xt, yt = _area_down_tab(w, nw), _area_down_tab(h, nh)
    xd = np.array([t[0] for t in xt])
    xs = np.array([t[1] for t in xt])
    xa = np.array([t[2] for t in xt], np.float32)
    S = img.astype(np.float32)
    out = np.zeros((nh, nw, C), np.uint8)
    acc = np.zeros((nw, C), np.float32)
    prev = yt[0][0]
    for dy, sy, beta in yt:
        buf = np.zeros((nw, C), np.float32)
        for k in range(len(xt)):                       # sequential: entries of one destination column add in table order
            buf[xd[k]] = buf[xd[k]] + S[sy, xs[k]] * xa[k]
        if dy != prev:
            out[prev] = np.clip(np.rint(acc), 0, 255).astype(np.uint8)
            acc = beta * buf
            prev = dy
        else:
            acc = acc + beta * buf
    out[prev] = np.clip(np.rint(acc), 0, 255).astype(np.uint8)
    return out


def roi_letterbox(rgb: np.ndarray, bbox, roi_size=(256, 256)):
    """apply_roi_filter canvas (roi.py:26-40)."""
    x, y, w, h = bbox
    H, W = roi_size
    roi = rgb[y:y + h, x:x + w]
    scale = min(W / max(w, 1), H / max(h, 1))
    nw, nh = max(int(w * scale), 1), max(int(h * scale), 1)
    res = resize_area_up(roi, nw, nh) if (nw >= w and nh >= h) else resize_area_down(roi, nw, nh)
    canvas = np.zeros((H, W, 3), np.uint8)
    oy, ox = (H - nh) // 2, (W - nw) // 2
    canvas[oy:oy + nh, ox:ox + nw] = res
    return canvas


# ------------------------------------------------------------ colour statistics
HIST_CATEGORIES = ("Vert Sain", "Vert Jaunatre", "Jaune", "Brun/Orange", "Rouge", "Zones Sombres",
                   "Zones Claires", "Violet/Pourpre")


def hist9(rgb: np.ndarray, mask: Optional[np.ndarray] = None) -> np.ndarray:
    """Masked 256-bin histograms of R,G,B,H,S,V,L,a,b -> int64 [9,256]."""
    sel = np.ones(rgb.shape[:2], bool) if mask is None else mask > 0
    planes = np.concatenate([rgb, sc.rgb_to_hsv(rgb), sc.rgb_to_lab(rgb)], axis=-1)
    return np.stack([np.bincount(planes[..., c][sel], minlength=256) for c in range(9)]).astype(np.int64)


def hist_counters(rgb: np.ndarray) -> np.ndarray:
    """hist.py numeric core on an (already masked) image: [leaf_px, 8 category counts (:38-65),
    5 hue-range counts (:248-256)] -> int64 [14]."""
    hsv = sc.rgb_to_hsv(rgb)
    h, s, v = (hsv[..., i].astype(np.int32) for i in range(3))
    m = (s > 10) & (v > 15) & (v < 245)                                        # hist.py:188
    cats = [
        m & (h >= 35) & (h <= 85) & (s >= 40) & (v >= 30),
        m & (h >= 20) & (h <= 40) & (s >= 25) & (v >= 30),
        m & (h >= 15) & (h <= 35) & (s >= 50) & (v >= 50),
        m & ((h <= 25) | (h >= 160)) & (s >= 30) & (v >= 20),
        m & (((h >= 160) & (h <= 180)) | (h <= 10)) & (s >= 40) & (v >= 30),
        m & (v <= 50) & (s >= 20),
        m & (v >= 200) & (s <= 30),
        m & (h >= 120) & (h <= 160) & (s >= 20),
    ]
    hues = [
        m & (h >= 35) & (h <= 85),
        m & (h >= 15) & (h <= 35),
        m & ((h <= 15) | (h >= 160)),
        m & (h >= 120) & (h <= 160),
        m & (h > 85) & (h < 120),          # second clause of hist.py:255 is always false
    ]
    return np.array([m.sum()] + [c.sum() for c in cats] + [c.sum() for c in hues], np.int64)


def hsv_hist_leaf(rgb: np.ndarray) -> np.ndarray:
    """256-bin H,S,V histograms over hist.py's leaf_mask (:188,:140-168) -> int64 [3,256]."""
    hsv = sc.rgb_to_hsv(rgb)
    s, v = hsv[..., 1], hsv[..., 2]
    m = (s > 10) & (v > 15) & (v < 245)
    return np.stack([np.bincount(hsv[..., c][m], minlength=256) for c in range(3)]).astype(np.int64)


def brown_spots(rgb: np.ndarray, mask: np.ndarray, cfg: Cfg):
    """apply_brown_filter numeric core (brown.py:21-89): (filtered mask, pct, count)."""
    leaf = mask > 0
    brown = brown_predicate(rgb, cfg) & leaf
    fp = sf.ellipse_footprint(cfg.brown_morph_kernel)
    clean = sf.morph_close(sf.morph_open(brown.astype(np.uint8) * 255, fp), fp)
    filt, cnt, px = filter_components8(clean, cfg.brown_min_area_px)
    pct = (px / max(int(leaf.sum()), 1)) * 100
    return filt, pct, cnt


# ------------------------------------------------------------ saliency "Blur" (blur.py:18-79)
def saliency_blur(rgb: np.ndarray, mask: np.ndarray, cfg: Cfg) -> np.ndarray:
    """apply_blur_filter given the mask its make_mask_func callback returned.

    Float stages (three min-max normalisations, 0.4/0.3/0.6/0.2 weights) follow OpenCV/NumPy
    float32 semantics; the uint8 stages are exact."""
    leaf = mask > 0
    gray = sc.rgb_to_gray(rgb)
    sal = np.zeros(gray.shape, np.float32)
    fp3 = sf.ellipse_footprint(3)
    edges = sf.dilate(sf.canny(gray, 50, 150, True), fp3)
    sal += edges.astype(np.float32) * np.float32(0.4)
    mag = sf.sobel_magnitude_f32(gray)
    gnorm = sf.normalize_minmax_f32(mag, 0, 255).astype(np.uint8)
    sal += gnorm.astype(np.float32) * np.float32(0.3)
    brown = brown_predicate(rgb, Cfg(brown_hue_range=cfg.brown_hue_range, brown_s_min=cfg.brown_s_min,
                                       brown_v_max=cfg.brown_v_max)) & leaf      # blur.py:47-53 is HSV-only
    bclean = sf.morph_close(brown.astype(np.uint8) * 255, fp3)
    bdil = sf.dilate(bclean, fp3, 2)
    sal += bdil.astype(np.float32) * np.float32(0.6)
    blurred = sf.gaussian_blur_u8(rgb, 15, 0)
    diff = np.abs(rgb.astype(np.float32) - blurred.astype(np.float32))
    cdiff = np.mean(diff, axis=2)                      # float32 mean over 3 channels
    sal += sf.normalize_minmax_f32(cdiff, 0, 255) * np.float32(0.2)
    snorm = sf.normalize_minmax_f32(sal, 0, 255).astype(np.uint8)
    sblur = sf.gaussian_blur_u8(snorm, 5, cfg.gaussian_sigma)
    res = np.where(leaf, sblur, 0).astype(np.uint8)
    return np.repeat(res[..., None], 3, axis=2)
