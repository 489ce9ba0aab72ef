"""NumPy spec of OpenCV's 8-bit RGB->GRAY / HSV / Lab conversions.

Restates ``cv2.cvtColor(rgb, COLOR_RGB2{GRAY,HSV,LAB})`` as called at
srcs/transform/filters/mask.py:87,103,623-624,736-737, blur.py:27,44,
brown.py:35,40, hist.py:184, analyze.py:119 (all under /root/reference).
Arithmetic follows OpenCV 4.13 imgproc color_{rgb,hsv,lab}.simd.hpp 8-bit paths
(third-party dependency, unpinned by the reference: requirements.txt:10).
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
from __future__ import annotations

import numpy as np

# ---------------------------------------------------------------- GRAY
R2Y, G2Y, B2Y, GRAY_SHIFT = 9798, 19235, 3735, 15


def rgb_to_gray(rgb: np.ndarray) -> np.ndarray:
    """(9798 R + 19235 G + 3735 B + 2^14) >> 15  (blur.py:27, mask.py:766)."""
    r = rgb[..., 0].astype(np.int32)
    g = rgb[..., 1].astype(np.int32)
    b = rgb[..., 2].astype(np.int32)
    return ((r * R2Y + g * G2Y + b * B2Y + (1 << (GRAY_SHIFT - 1))) >> GRAY_SHIFT).astype(np.uint8)


# ---------------------------------------------------------------- HSV
HSV_SHIFT = 12


def hsv_div_tables():
    """sdiv[v] = rint((255<<12)/v), hdiv180[d] = rint((180<<12)/(6 d)); index 0 -> 0."""
    i = np.arange(1, 256, dtype=np.float64)
    sdiv = np.zeros(256, np.int32)
    hdiv = np.zeros(256, np.int32)
    sdiv[1:] = np.rint((255 << HSV_SHIFT) / i).astype(np.int32)
    hdiv[1:] = np.rint((180 << HSV_SHIFT) / (6.0 * i)).astype(np.int32)
    return sdiv, hdiv


_SDIV, _HDIV = hsv_div_tables()


def rgb_to_hsv(rgb: np.ndarray) -> np.ndarray:
    """8-bit HSV, H in [0,179] (mask.py:87, hist.py:184, brown.py:40)."""
    r = rgb[..., 0].astype(np.int32)
    g = rgb[..., 1].astype(np.int32)
    b = rgb[..., 2].astype(np.int32)
    v = np.maximum(np.maximum(r, g), b)
    vmin = np.minimum(np.minimum(r, g), b)
    diff = v - vmin
    vr = v == r
    vg = v == g
    s = (diff * _SDIV[v] + (1 << (HSV_SHIFT - 1))) >> HSV_SHIFT
    h = np.where(vr, g - b, np.where(vg, b - r + 2 * diff, r - g + 4 * diff))
    h = (h * _HDIV[diff] + (1 << (HSV_SHIFT - 1))) >> HSV_SHIFT
    h = np.where(h < 0, h + 180, h)
    return np.stack([h, s, v], axis=-1).astype(np.uint8)


# ---------------------------------------------------------------- Lab
LAB_GAMMA_SHIFT = 3
LAB_CBRT_TAB_SIZE_B = 256 * 3 // 2 * (1 << LAB_GAMMA_SHIFT)  # 3072
LAB_SHIFT = 12
LAB_SHIFT2 = 15  # lab_shift + gamma_shift
# sRGB->XYZ (D65), rows pre-divided by the white point, scaled by 2^12.
LAB_COEFFS = np.array([[1777, 1541, 778], [871, 2929, 296], [73, 448, 3575]], np.int64)
LAB_L_SCALE = (116 * 255 + 50) // 100            # 296
LAB_L_SHIFT = -((16 * 255 * (1 << LAB_SHIFT2) + 50) // 100)  # -1336934


def lab_tables():
    """gtab[256] (float64-built sRGB gamma, x2040) and ctab[3072] (float32 cube root, x32768)."""
    x = np.arange(256, dtype=np.float64) / 255.0
    g = np.where(x <= 0.04045, x / 12.92, ((x + 0.055) / 1.055) ** 2.4)
    gtab = np.rint(g * (255.0 * (1 << LAB_GAMMA_SHIFT))).astype(np.int64)
    j = np.arange(LAB_CBRT_TAB_SIZE_B, dtype=np.float32)
    xx = j / np.float32(255.0 * (1 << LAB_GAMMA_SHIFT))
    f = np.where(xx < np.float32(0.008856),
                 xx * np.float32(7.787) + np.float32(0.13793103448275862),
                 np.cbrt(xx).astype(np.float32)).astype(np.float32)
    ctab = np.rint(np.float32(1 << LAB_SHIFT2) * f).astype(np.int64)
    return gtab, ctab


_GTAB, _CTAB = lab_tables()


def _descale(x, n):
    return (x + (1 << (n - 1))) >> n


def rgb_to_lab(rgb: np.ndarray) -> np.ndarray:
    """8-bit CIE Lab (mask.py:103,624,737; brown.py:35)."""
    r = _GTAB[rgb[..., 0]]
    g = _GTAB[rgb[..., 1]]
    b = _GTAB[rgb[..., 2]]
    c = LAB_COEFFS
    fx = _CTAB[_descale(r * c[0, 0] + g * c[0, 1] + b * c[0, 2], LAB_SHIFT)]
    fy = _CTAB[_descale(r * c[1, 0] + g * c[1, 1] + b * c[1, 2], LAB_SHIFT)]
    fz = _CTAB[_descale(r * c[2, 0] + g * c[2, 1] + b * c[2, 2], LAB_SHIFT)]
    L = _descale(LAB_L_SCALE * fy + LAB_L_SHIFT, LAB_SHIFT2)
    a = _descale(500 * (fx - fy) + 128 * (1 << LAB_SHIFT2), LAB_SHIFT2)
    bb = _descale(200 * (fy - fz) + 128 * (1 << LAB_SHIFT2), LAB_SHIFT2)
    out = np.stack([L, a, bb], axis=-1)
    return np.clip(out, 0, 255).astype(np.uint8)


def all_colours() -> np.ndarray:
    """All 2^24 RGB triples as a [4096,4096,3] uint8 image (exhaustive-domain tests)."""
    i = np.arange(1 << 24, dtype=np.uint32)
    rgb = np.stack([(i >> 16) & 255, (i >> 8) & 255, i & 255], axis=-1).astype(np.uint8)
    return rgb.reshape(4096, 4096, 3)
