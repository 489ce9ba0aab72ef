"""NumPy spec of the six class-balancing augmentations.

Restates the Pillow 12.2.0 / NumPy 2.3.5 arithmetic behind
srcs/preprocessing/image_augmenter.py:20-133 (flip :24,:26; rotate :37;
skew :50-66; shear :79-89; crop :99-109; distortion :119-127) and the train-input
tail srcs/utils/image_utils.py:109-130 (Lanczos resize + /255).
Third-party arithmetic (Pillow libImaging Geometry.c / Resample.c, ImageOps.py,
NumPy legacy MT19937) is not under /root/reference; versions unpinned there.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
from __future__ import annotations

import math
import random as _random

import numpy as np

# ---------------------------------------------------------------- parameters
# The reference draws every parameter from Python's `random` (and NumPy's legacy
# global stream for the noise) in a fixed order; these helpers repeat the draws so
# that spec, refcalls and the CUDA host shim see identical parameters.


def draw_flip(rng=_random):
    """image_augmenter.py:23 -> True = FLIP_LEFT_RIGHT, False = FLIP_TOP_BOTTOM."""
    return rng.choice([True, False])


def draw_rotate(rng=_random):
    return rng.uniform(-30, 30)                       # :36


def draw_skew(rng=_random):
    return rng.uniform(0.05, 0.15)                    # :48


def draw_shear(rng=_random):
    k = rng.uniform(-0.2, 0.2)                        # :77
    horiz = rng.choice([True, False])                 # :79
    return k, horiz


def draw_crop(width, height, rng=_random):
    r = rng.uniform(0.8, 0.95)                        # :101
    nw, nh = int(width * r), int(height * r)          # :102-103
    left = rng.randint(0, width - nw)                 # :105
    top = rng.randint(0, height - nh)                 # :106
    return left, top, nw, nh


def skew_coeffs(s, width, height):
    """PERSPECTIVE coefficients of :50-59 (g = h = 0)."""
    return [1 + s, 0, -s * width, 0, 1 + s, -s * height, 0, 0]


def shear_coeffs(k, horiz):
    return [1, k, 0, 0, 1, 0, 0, 0] if horiz else [1, 0, 0, k, 1, 0, 0, 0]  # :80-82


# ---------------------------------------------------------------- flip
def flip(img: np.ndarray, left_right: bool) -> np.ndarray:
    return img[:, ::-1].copy() if left_right else img[::-1].copy()


# ---------------------------------------------------------------- rotate (NEAREST, expand, white fill)
def rotate_params(angle: float, w: int, h: int):
    """PIL Image.rotate(angle, expand=True): returns (matrix[6], nw, nh) or a transpose tag.

    Follows PIL/Image.py rotate(): angle %= 360; multiples of 90 use transpose fast paths.
    """
    angle = angle % 360.0
    if angle == 0:
        return ("copy", w, h)
    if angle == 180:
        return ("rot180", w, h)
    if angle in (90, 270):
        return ("rot90" if angle == 90 else "rot270", h, w)
    cx, cy = w / 2.0, h / 2.0
    a = -math.radians(angle)
    m = [round(math.cos(a), 15), round(math.sin(a), 15), 0.0,
         round(-math.sin(a), 15), round(math.cos(a), 15), 0.0]

    def tf(x, y, mm):
        (a_, b_, c_, d_, e_, f_) = mm
        return a_ * x + b_ * y + c_, d_ * x + e_ * y + f_
    m[2], m[5] = tf(-cx, -cy, m)
    m[2] += cx
    m[5] += cy
    xx, yy = [], []
    for x, y in ((0, 0), (w, 0), (w, h), (0, h)):
        tx, ty = tf(x, y, m)
        xx.append(tx)
        yy.append(ty)
    nw = math.ceil(max(xx)) - math.floor(min(xx))
    nh = math.ceil(max(yy)) - math.floor(min(yy))
    m[2], m[5] = tf(-(nw - w) / 2.0, -(nh - h) / 2.0, m)
    return (m, nw, nh)


def affine_fixed_coeffs(m):
    """libImaging Geometry.c affine_fixed(): 16.16 coefficients incl. the half-pixel offset."""
    def fix(v):
        return int(math.floor(v * 65536.0 + 0.5))
    a0, a1, a3, a4 = fix(m[0]), fix(m[1]), fix(m[3]), fix(m[4])
    a2 = fix(m[2] + m[0] * 0.5 + m[1] * 0.5)
    a5 = fix(m[5] + m[3] * 0.5 + m[4] * 0.5)
    return a0, a1, a2, a3, a4, a5


def rotate_nn(img: np.ndarray, angle: float, fill=255) -> np.ndarray:
    """img.rotate(angle, expand=True, fillcolor='white') with the default NEAREST filter."""
    h, w = img.shape[:2]
    p = rotate_params(angle, w, h)
    if p[0] == "copy":
        return img.copy()
    if p[0] == "rot180":
        return img[::-1, ::-1].copy()
    if p[0] == "rot90":
        return np.transpose(img, (1, 0, 2))[::-1].copy()
    if p[0] == "rot270":
        return np.transpose(img, (1, 0, 2))[:, ::-1].copy()
    m, nw, nh = p
    a0, a1, a2, a3, a4, a5 = affine_fixed_coeffs(m)
    ys, xs = np.mgrid[0:nh, 0:nw].astype(np.int64)
    xin = (a2 + ys * a1 + xs * a0) >> 16
    yin = (a5 + ys * a4 + xs * a3) >> 16
    inside = (xin >= 0) & (xin < w) & (yin >= 0) & (yin < h)
    out = np.full((nh, nw, img.shape[2]), fill, np.uint8)
    out[inside] = img[yin[inside], xin[inside]]
    return out


# ---------------------------------------------------------------- bicubic warp (AFFINE / PERSPECTIVE)
def warp_bicubic(img: np.ndarray, coeffs, perspective: bool) -> np.ndarray:
    """img.transform(size, AFFINE|PERSPECTIVE, coeffs, BICUBIC): fp64, a=-1 cubic, truncation.

    libImaging Geometry.c affine_transform / perspective_transform + bicubic_filter32RGB.
    `coeffs` has 8 entries (a6=a7=0 for AFFINE).
    """
    h, w = img.shape[:2]
    a = [float(c) for c in coeffs]
    ys, xs = np.mgrid[0:h, 0:w]
    xc = xs + 0.5
    yc = ys + 0.5
    if perspective:
        den = a[6] * xc + a[7] * yc + 1.0
        xin = (a[0] * xc + a[1] * yc + a[2]) / den
        yin = (a[3] * xc + a[4] * yc + a[5]) / den
    else:
        xin = a[0] * xc + a[1] * yc + a[2]
        yin = a[3] * xc + a[4] * yc + a[5]
    outside = (xin < 0.0) | (xin >= w) | (yin < 0.0) | (yin >= h)
    xin = xin - 0.5
    yin = yin - 0.5

    def floor_c(v):  # C: (v < 0 ? (int)floor(v) : (int)v)
        return np.floor(v).astype(np.int64)
    x0 = floor_c(xin)
    y0 = floor_c(yin)
    dx = xin - x0
    dy = yin - y0
    x0 -= 1
    y0 -= 1
    src = img.astype(np.float64)
    xi = [np.clip(x0 + k, 0, w - 1) for k in range(4)]

    def cubic(v1, v2, v3, v4, d):
        p1 = v2
        p2 = -v1 + v3
        p3 = 2 * (v1 - v2) + v3 - v4
        p4 = -v1 + v2 - v3 + v4
        return p1 + d * (p2 + d * (p3 + d * p4))
    out = np.zeros(img.shape, np.uint8)
    dxe = dx[..., None]
    dye = dy[..., None]
    rows = []
    for k in range(4):
        yk = y0 + k
        if k == 0:
            yk_c = np.clip(yk, 0, h - 1)
            r = cubic(src[yk_c, xi[0]], src[yk_c, xi[1]], src[yk_c, xi[2]], src[yk_c, xi[3]], dxe)
        else:
            ok = (yk >= 0) & (yk < h)
            yk_c = np.clip(yk, 0, h - 1)
            r = cubic(src[yk_c, xi[0]], src[yk_c, xi[1]], src[yk_c, xi[2]], src[yk_c, xi[3]], dxe)
            r = np.where(ok[..., None], r, rows[k - 1])
        rows.append(r)
    v = cubic(rows[0], rows[1], rows[2], rows[3], dye)
    res = np.where(v <= 0.0, 0, np.where(v >= 255.0, 255, np.trunc(v))).astype(np.uint8)
    res[outside] = 0
    out[...] = res
    return out


# ---------------------------------------------------------------- Lanczos resize (8 bpc)
PRECISION_BITS = 32 - 8 - 2


def _lanczos(x):
    x = abs(x)
    if x == 0.0:
        return 1.0
    if x < 3.0:
        px = math.pi * x
        return (math.sin(px) / px) * (math.sin(px / 3.0) / (px / 3.0))
    return 0.0


def lanczos_coeffs(in_size: int, out_size: int, in0: float = 0.0, in1: float | None = None):
    """libImaging Resample.c precompute_coeffs + normalize_coeffs_8bpc.

    Returns (ksize, bounds[out,2] = (xmin, count), kk[out,ksize] int32 fixed-point 2^22)."""
    in1 = float(in_size) if in1 is None else in1
    support = 3.0
    scale = filterscale = (in1 - in0) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = support * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = in0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = [0.0] * ksize
        ww = 0.0
        for x in range(xmax):
            wgt = _lanczos((x + xmin - center + 0.5) * ss)
            k[x] = wgt
            ww += wgt
        for x in range(xmax):
            if ww != 0.0:
                k[x] /= ww
        for x in range(ksize):
            v = k[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _resample_axis1(a: np.ndarray, out_size: int) -> np.ndarray:
    """One 8-bit pass along axis 1 of a [H,W,C] array."""
    in_size = a.shape[1]
    ksize, bounds, kk = lanczos_coeffs(in_size, out_size)
    src = a.astype(np.int64)
    out = np.zeros((a.shape[0], out_size, a.shape[2]), np.int64)
    for xx in range(out_size):
        xmin, cnt = bounds[xx]
        acc = np.full((a.shape[0], a.shape[2]), 1 << (PRECISION_BITS - 1), np.int64)
        acc += (src[:, xmin:xmin + cnt, :] * kk[xx, :cnt][None, :, None]).sum(axis=1)
        out[:, xx, :] = acc >> PRECISION_BITS
    return np.clip(out, 0, 255).astype(np.uint8)


def resize_lanczos(img: np.ndarray, out_w: int, out_h: int) -> np.ndarray:
    """Image.resize((out_w,out_h), LANCZOS) for uint8 HWC: horizontal pass -> uint8 -> vertical pass."""
    h, w = img.shape[:2]
    cur = img
    if out_w != w:
        cur = _resample_axis1(cur, out_w)
    if out_h != h:
        cur = np.transpose(_resample_axis1(np.transpose(cur, (1, 0, 2)), out_h), (1, 0, 2))
    return np.ascontiguousarray(cur)


def crop_resize(img: np.ndarray, left: int, top: int, nw: int, nh: int) -> np.ndarray:
    """img.crop(box).resize((W,H), LANCZOS) (image_augmenter.py:108-109)."""
    h, w = img.shape[:2]
    return resize_lanczos(np.ascontiguousarray(img[top:top + nh, left:left + nw]), w, h)


def resize_normalize(img: np.ndarray, size: int = 224):
    """sequence.py:84-88: Lanczos to size x size, then float32 /255.0. Returns (u8, f32)."""
    u8 = resize_lanczos(img, size, size)
    return u8, u8.astype(np.float32) / np.float32(255.0)


# ---------------------------------------------------------------- distortion
class MT19937:
    """NumPy legacy global-stream generator: init_genrand seeding, genrand_res53 doubles,
    polar-method gauss with a one-value cache (numpy/random/_legacy: legacy_gauss)."""

    def __init__(self, seed: int):
        mt = np.zeros(624, np.uint32)
        s = seed & 0xFFFFFFFF
        mt_l = [0] * 624
        mt_l[0] = s
        for i in range(1, 624):
            mt_l[i] = (1812433253 * (mt_l[i - 1] ^ (mt_l[i - 1] >> 30)) + i) & 0xFFFFFFFF
        self.mt = np.array(mt_l, np.uint32)
        self.pos = 624

    def _gen(self):
        mt = self.mt.astype(np.uint64)
        U, L, A = 0x80000000, 0x7FFFFFFF, 0x9908B0DF
        out = mt.copy()
        for lo, hi in ((0, 227), (227, 454), (454, 623)):
            i = np.arange(lo, hi)
            y = (out[i] & U) | (out[i + 1] & L)
            out[i] = out[(i + 397) % 624] ^ (y >> np.uint64(1)) ^ np.where(y & np.uint64(1), A, 0).astype(np.uint64)
        y = (out[623] & U) | (out[0] & L)
        out[623] = out[396] ^ (y >> np.uint64(1)) ^ (A if int(y) & 1 else 0)
        self.mt = out.astype(np.uint32)
        self.pos = 0

    def words(self, n: int) -> np.ndarray:
        res = np.empty(n, np.uint32)
        got = 0
        while got < n:
            if self.pos >= 624:
                self._gen()
            take = min(624 - self.pos, n - got)
            y = self.mt[self.pos:self.pos + take].astype(np.uint32)
            y = y ^ (y >> 11)
            y = y ^ ((y << 7) & np.uint32(0x9D2C5680))
            y = y ^ ((y << 15) & np.uint32(0xEFC60000))
            y = y ^ (y >> 18)
            res[got:got + take] = y
            self.pos += take
            got += take
        return res

    def normals(self, n: int, loc: float = 0.0, scale: float = 1.0) -> np.ndarray:
        """First n samples of np.random.normal(loc, scale) after np.random.seed(seed).

        Parallel form of legacy_gauss: attempt t consumes doubles d[2t], d[2t+1];
        x1 = 2d-1, x2 = 2d'-1, accepted iff 0 < r2 < 1; the k-th accepted attempt yields
        normals 2k -> f*x2 and 2k+1 -> f*x1 with f = sqrt(-2 ln(r2)/r2).
        (The stream is consumed; use a fresh instance per image, as the reference does
        through ImageAugmenter(seed) -- dataset_balancer.py:203.)
        """
        need_pairs = (n + 1) // 2
        att = int(need_pairs * 1.35) + 64
        w = self.words(4 * att).astype(np.uint64)
        while True:
            d = ((w[0::2] >> np.uint64(5)) * 67108864.0 + (w[1::2] >> np.uint64(6))) / 9007199254740992.0
            x1 = 2.0 * d[0::2] - 1.0
            x2 = 2.0 * d[1::2] - 1.0
            r2 = x1 * x1 + x2 * x2
            idx = np.nonzero((r2 < 1.0) & (r2 != 0.0))[0]
            if len(idx) >= need_pairs:
                break
            w = np.concatenate([w, self.words(4 * att).astype(np.uint64)])
        idx = idx[:need_pairs]
        f = np.sqrt(-2.0 * np.log(r2[idx]) / r2[idx])
        pair = np.empty(2 * len(idx), np.float64)
        pair[0::2] = f * x2[idx]
        pair[1::2] = f * x1[idx]
        return loc + scale * pair[:n]


def noise_u8(noise_f64: np.ndarray) -> np.ndarray:
    """`.astype(np.uint8)` of a float64 array on x86-64 NumPy 2.3: truncate toward zero, mod 256."""
    return (np.trunc(noise_f64).astype(np.int64) & 255).astype(np.uint8)


def autocontrast_lut(hist: np.ndarray, cutoff: float) -> np.ndarray:
    """PIL ImageOps.autocontrast(cutoff) LUT for one channel's 256-bin histogram."""
    h = [int(v) for v in hist]
    if cutoff:
        n = sum(h)
        cut = int(n * cutoff // 100)
        for lo in range(256):
            if cut > h[lo]:
                cut -= h[lo]
                h[lo] = 0
            else:
                h[lo] -= cut
                cut = 0
            if cut <= 0:
                break
        cut = int(n * cutoff // 100)
        for hi in range(255, -1, -1):
            if cut > h[hi]:
                cut -= h[hi]
                h[hi] = 0
            else:
                h[hi] -= cut
                cut = 0
            if cut <= 0:
                break
    lo = 0
    for lo in range(256):
        if h[lo]:
            break
    hi = 255
    for hi in range(255, -1, -1):
        if h[hi]:
            break
    if hi <= lo:
        return np.arange(256, dtype=np.uint8)
    scale = 255.0 / (hi - lo)
    offset = -lo * scale
    lut = np.empty(256, np.uint8)
    for ix in range(256):
        v = int(ix * scale + offset)
        lut[ix] = 0 if v < 0 else 255 if v > 255 else v
    return lut


def distortion(img: np.ndarray, noise_u8_arr: np.ndarray, cutoff: float) -> np.ndarray:
    """image_augmenter.py:119-127 given the uint8-cast noise: wrap-around add, per-channel autocontrast."""
    x = (img.astype(np.uint16) + noise_u8_arr.astype(np.uint16)).astype(np.uint8)  # uint8 + uint8 wraps
    out = np.empty_like(x)
    for c in range(x.shape[2]):
        hist = np.bincount(x[..., c].ravel(), minlength=256)
        out[..., c] = autocontrast_lut(hist, cutoff)[x[..., c]]
    return out
