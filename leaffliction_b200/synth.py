"""Seeded synthetic "leaf-like" images: the dataset is not shipped with
the reference (README section 4), so benches and tests use this generator.

Background uniform colour (grey / purple-grey / near-white), a lobed elliptical green leaf with
jittered centre, axes and rotation, 0-12 brown elliptical spots, +-12 uniform noise and an
optional soft shadow.  uint8 HWC.  Pure NumPy, host only.
"""
from __future__ import annotations

import numpy as np

_BACKGROUNDS = ((150, 140, 160), (132, 120, 150), (236, 236, 232))


def leaf_image(index: int, h: int = 256, w: int = 256, seed: int = 1234) -> np.ndarray:
    rng = np.random.default_rng(seed + index)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.empty((h, w, 3), np.float32)
    img[...] = _BACKGROUNDS[int(rng.integers(0, len(_BACKGROUNDS)))]
    cx = w * (0.5 + rng.uniform(-0.08, 0.08))
    cy = h * (0.5 + rng.uniform(-0.08, 0.08))
    ax = w * rng.uniform(0.26, 0.38)
    ay = h * rng.uniform(0.16, 0.27)
    th = rng.uniform(0, np.pi)
    c, s = np.cos(th), np.sin(th)
    u = ((xx - cx) * c + (yy - cy) * s) / ax
    v = (-(xx - cx) * s + (yy - cy) * c) / ay
    ang = np.arctan2(v, u)
    lobes = 1.0 + 0.06 * np.cos(rng.integers(3, 8) * ang + rng.uniform(0, 6.28))
    leaf = (u * u + v * v) < lobes * lobes
    if rng.random() < 0.5:  # soft shadow next to the leaf
        sh = (((xx - cx - 0.06 * w) / (ax * 1.05)) ** 2 + ((yy - cy - 0.05 * h) / (ay * 1.1)) ** 2) < 1.0
        img[sh & ~leaf] *= 0.82
    green = np.array((60, 140, 50), np.float32) + rng.uniform(-18, 18, 3).astype(np.float32)
    img[leaf] = green
    vein = np.abs(v) < 0.02
    img[leaf & vein] = green * 0.8
    for _ in range(int(rng.integers(0, 13))):
        r = rng.uniform(2, 10) * (w / 256.0)
        a = rng.uniform(0, 6.28)
        d = rng.uniform(0, 0.8)
        sx = cx + (d * ax * np.cos(a)) * c - (d * ay * np.sin(a)) * s
        sy = cy + (d * ax * np.cos(a)) * s + (d * ay * np.sin(a)) * c
        spot = ((xx - sx) / r) ** 2 + ((yy - sy) / (r * rng.uniform(0.6, 1.0))) ** 2 < 1.0
        img[spot & leaf] = np.array((120, 70, 30), np.float32) + rng.uniform(-12, 12, 3).astype(np.float32)
    img += rng.integers(-12, 13, (h, w, 3)).astype(np.float32)
    return np.clip(img, 0, 255).astype(np.uint8)


def leaf_batch(n: int, h: int = 256, w: int = 256, seed: int = 1234, start: int = 0) -> np.ndarray:
    out = np.empty((n, h, w, 3), np.uint8)
    for i in range(n):
        out[i] = leaf_image(start + i, h, w, seed)
    return out


def adversarial_images(h: int = 64, w: int = 64):
    """Edge cases the parity tests sweep: flat images, 1-px lines, diagonal pinches,
    checkerboards, salt noise, frames touching the border."""
    rng = np.random.default_rng(99)
    out = {}
    out["black"] = np.zeros((h, w, 3), np.uint8)
    out["white"] = np.full((h, w, 3), 255, np.uint8)
    out["green"] = np.full((h, w, 3), (60, 140, 50), np.uint8)
    g = np.full((h, w, 3), (150, 140, 160), np.uint8)
    a = g.copy(); a[h // 2, :] = (60, 140, 50); out["hline"] = a
    a = g.copy(); a[:, w // 3] = (60, 140, 50); out["vline"] = a
    a = g.copy()
    for i in range(min(h, w)):
        a[i, i] = (60, 140, 50)
    out["diag"] = a
    a = g.copy(); yy, xx = np.mgrid[0:h, 0:w]; a[(yy + xx) % 2 == 0] = (60, 140, 50); out["checker"] = a
    a = g.copy(); a[rng.random((h, w)) < 0.2] = (60, 140, 50); out["salt"] = a
    a = g.copy(); a[2:-2, 2:-2] = (60, 140, 50); a[h // 4:3 * h // 4, w // 4:3 * w // 4] = (150, 140, 160); out["frame"] = a
    a = np.full((h, w, 3), (60, 140, 50), np.uint8); a[h // 3:h // 2, w // 3:w // 2] = (150, 140, 160); out["hole_border"] = a
    a = g.copy(); a[4:h // 2, 4:w // 2] = (60, 140, 50); a[h // 2:h - 4, w // 2:w - 4] = (60, 140, 50); out["pinch"] = a
    a = g.copy(); a[4:20, 4:20] = (60, 140, 50); a[4:20, 30:46] = (60, 140, 50); a[30:46, 4:20] = (60, 140, 50); out["ties"] = a
    out["noise"] = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    return out
