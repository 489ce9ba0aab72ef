"""NumPy spec of the OpenCV stencil filters on the hot path.

GaussianBlur (blur.py:61,72; mask.py:223,770), Canny (mask.py:679-680,789;
blur.py:30; analyze.py:120), Sobel float magnitude (blur.py:35-37; mask.py:160-162),
binary morphology with MORPH_ELLIPSE footprints (mask.py:63-64,341,367-370,806-829;
blur.py:31-32,55-58; brown.py:52-59), Otsu threshold (pcv.threshold.otsu via
mask.py:78,83,400) and min-max normalisation (blur.py:38,65,68).
Third-party arithmetic: OpenCV 4.13.0 (unpinned by the reference).
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
from __future__ import annotations

import math

import numpy as np

# ------------------------------------------------------------ Gaussian


def gaussian_kernel_q8(ksize: int, sigma: float) -> np.ndarray:
    """8-fractional-bit Gaussian taps with edge->centre error diffusion, sum == 256.

    cv::getGaussianKernelBitExact + getGaussianKernelFixedPoint_ED (smooth.dispatch.cpp).
    sigma <= 0 => 0.3*((k-1)/2 - 1) + 0.8; k in {1,3,5,7} with sigma<=0 use OpenCV's
    small fixed table.
    """
    n = int(ksize)
    small = {1: [1.0], 3: [0.25, 0.5, 0.25], 5: [0.0625, 0.25, 0.375, 0.25, 0.0625],
             7: [0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125]}
    if sigma <= 0 and n in small:
        k = np.array(small[n], np.float64)
    else:
        s = sigma if sigma > 0 else ((n - 1) * 0.5 - 1) * 0.3 + 0.8
        scale2x = -0.5 / (s * s)
        x = np.arange(n, dtype=np.float64) - (n - 1) * 0.5
        k = np.exp(scale2x * x * x)
        k = k * (1.0 / k.sum())
    out = np.zeros(n, np.int64)
    err = 0.0
    tot = 0
    for i in range(n // 2):
        adj = k[i] * 256.0 + err
        v = int(np.rint(adj))
        err = adj - v
        out[i] = out[n - 1 - i] = v
        tot += v
    out[n // 2] = 256 - 2 * tot
    return out


def reflect101(i: np.ndarray, n: int) -> np.ndarray:
    """BORDER_REFLECT_101 index map (gfedcb|abcdefgh|gfedcba)."""
    if n == 1:
        return np.zeros_like(i)
    p = 2 * (n - 1)
    i = np.mod(i, p)
    return np.where(i >= n, p - i, i)


def gaussian_blur_u8(img: np.ndarray, ksize: int, sigma: float = 0.0) -> np.ndarray:
    """cv2.GaussianBlur(img,(k,k),sigma) for uint8: 8.8 H pass, 16.16 V pass, round."""
    kq = gaussian_kernel_q8(ksize, sigma)
    r = ksize // 2
    a = img.astype(np.int64)
    if a.ndim == 2:
        a = a[..., None]
    H, W = a.shape[:2]
    xs = reflect101(np.arange(-r, W + r), W)
    ys = reflect101(np.arange(-r, H + r), H)
    ap = a[:, xs]
    h = np.zeros_like(a)
    for t in range(ksize):
        h += ap[:, t:t + W] * kq[t]
    hp = h[ys]
    v = np.zeros_like(a)
    for t in range(ksize):
        v += hp[t:t + H] * kq[t]
    out = ((v + 32768) >> 16).astype(np.uint8)
    return out.reshape(img.shape)


# ------------------------------------------------------------ Sobel / Canny
def _pad_replicate(a, r=1):
    return np.pad(a, r, mode="edge")


def _pad_reflect101(a, r=1):
    return np.pad(a, r, mode="reflect")


def sobel3(gray: np.ndarray, border: str = "reflect101"):
    """3x3 Sobel dx, dy as int32. cv2.Sobel default border is REFLECT_101; Canny uses REPLICATE."""
    g = gray.astype(np.int32)
    p = _pad_reflect101(g) if border == "reflect101" else _pad_replicate(g)
    H, W = g.shape

    def s(dy, dx):
        return p[1 + dy:1 + dy + H, 1 + dx:1 + dx + W]
    gx = (s(-1, 1) + 2 * s(0, 1) + s(1, 1)) - (s(-1, -1) + 2 * s(0, -1) + s(1, -1))
    gy = (s(1, -1) + 2 * s(1, 0) + s(1, 1)) - (s(-1, -1) + 2 * s(-1, 0) + s(-1, 1))
    return gx, gy


def sobel_magnitude_f32(gray: np.ndarray) -> np.ndarray:
    """cv2.magnitude(Sobel(CV_32F,1,0), Sobel(CV_32F,0,1)) (blur.py:35-37): float32 sqrt."""
    gx, gy = sobel3(gray, "reflect101")
    fx = gx.astype(np.float32)
    fy = gy.astype(np.float32)
    return np.sqrt(fx * fx + fy * fy, dtype=np.float32)


CANNY_TG22 = 13573  # int(0.4142135623730950488016887242097 * (1 << 15) + 0.5)


def label8(mask: np.ndarray) -> np.ndarray:
    """8-connected labels (0 = background), raster-order numbering. Pure NumPy union-find on runs
    is overkill for an oracle: scipy.ndimage is used when present, else a slow BFS."""
    try:
        from scipy import ndimage as ndi
        lab, _ = ndi.label(mask, structure=np.ones((3, 3), np.int32))
        return lab
    except Exception:  # pragma: no cover
        return _label_bfs(mask, 8)


def label4(mask: np.ndarray) -> np.ndarray:
    try:
        from scipy import ndimage as ndi
        lab, _ = ndi.label(mask)
        return lab
    except Exception:  # pragma: no cover
        return _label_bfs(mask, 4)


def _label_bfs(mask, conn):  # pragma: no cover - fallback only
    H, W = mask.shape
    lab = np.zeros((H, W), np.int32)
    nb = [(-1, 0), (1, 0), (0, -1), (0, 1)]
    if conn == 8:
        nb += [(-1, -1), (-1, 1), (1, -1), (1, 1)]
    cur = 0
    for y in range(H):
        for x in range(W):
            if mask[y, x] and not lab[y, x]:
                cur += 1
                st = [(y, x)]
                lab[y, x] = cur
                while st:
                    cy, cx = st.pop()
                    for dy, dx in nb:
                        ny, nx = cy + dy, cx + dx
                        if 0 <= ny < H and 0 <= nx < W and mask[ny, nx] and not lab[ny, nx]:
                            lab[ny, nx] = cur
                            st.append((ny, nx))
    return lab


def canny(gray: np.ndarray, low: float, high: float, l2: bool = False) -> np.ndarray:
    """cv2.Canny(gray, low, high, apertureSize=3, L2gradient=l2) -> uint8 {0,255}.

    Sobel on BORDER_REPLICATE; magnitude |dx|+|dy| or dx^2+dy^2 (thresholds squared);
    NMS with TG22 fixed point; hysteresis = 8-connected candidate components that contain
    a pixel with magnitude > high.
    """
    if low > high:
        low, high = high, low
    gx, gy = sobel3(gray, "replicate")
    gx = gx.astype(np.int64)
    gy = gy.astype(np.int64)
    if l2:
        low = min(32767.0, low)
        high = min(32767.0, high)
        if low > 0:
            low *= low
        if high > 0:
            high *= high
        mag = gx * gx + gy * gy
    else:
        mag = np.abs(gx) + np.abs(gy)
    lo = int(math.floor(low))
    hi = int(math.floor(high))
    H, W = gray.shape
    mp = np.pad(mag, 1, mode="constant")

    def m(dy, dx):
        return mp[1 + dy:1 + dy + H, 1 + dx:1 + dx + W]
    ax = np.abs(gx)
    ay = np.abs(gy) << 15
    tg22x = ax * CANNY_TG22
    tg67x = tg22x + (ax << 16)
    horiz = ay < tg22x
    vert = ay > tg67x
    s = np.where((gx ^ gy) < 0, -1, 1)
    # diagonal neighbours: (y-1, x-s) and (y+1, x+s)
    d1 = np.where(s < 0, m(-1, 1), m(-1, -1))
    d2 = np.where(s < 0, m(1, -1), m(1, 1))
    keep_h = (mag > m(0, -1)) & (mag >= m(0, 1))
    keep_v = (mag > m(-1, 0)) & (mag >= m(1, 0))
    keep_d = (mag > d1) & (mag > d2)
    keep = np.where(horiz, keep_h, np.where(vert, keep_v, keep_d))
    cand = (mag > lo) & keep
    strong = cand & (mag > hi)
    lab = label8(cand)
    good = np.zeros(lab.max() + 1, bool)
    good[np.unique(lab[strong])] = True
    good[0] = False
    return (good[lab].astype(np.uint8)) * 255


# ------------------------------------------------------------ morphology
def ellipse_footprint(kw: int, kh: int | None = None) -> np.ndarray:
    """cv2.getStructuringElement(MORPH_ELLIPSE,(kw,kh)) (OpenCV 4.13 morph.dispatch.cpp)."""
    kh = kw if kh is None else kh
    r = kh // 2
    c = kw // 2
    inv_r2 = 1.0 / (r * r) if r else 0.0
    el = np.zeros((kh, kw), np.uint8)
    for i in range(kh):
        dy = i - r
        if abs(dy) <= r:
            dx = int(np.rint(c * math.sqrt((r * r - dy * dy) * inv_r2)))
            j1 = max(c - dx, 0)
            j2 = min(c + dx + 1, kw)
            el[i, j1:j2] = 1
    return el


def _morph(mask: np.ndarray, fp: np.ndarray, dilate: bool, iterations: int = 1) -> np.ndarray:
    """Binary erode/dilate, anchor = centre (k//2), out-of-image ignored (default border)."""
    kh, kw = fp.shape
    ay, ax = kh // 2, kw // 2
    b = mask > 0
    H, W = b.shape
    for _ in range(iterations):
        pad_val = False if dilate else True
        p = np.full((H + kh, W + kw), pad_val, bool)
        p[ay:ay + H, ax:ax + W] = b
        out = np.zeros((H, W), bool) if dilate else np.ones((H, W), bool)
        for i in range(kh):
            for j in range(kw):
                if fp[i, j]:
                    sl = p[i:i + H, j:j + W]
                    out = (out | sl) if dilate else (out & sl)
        b = out
    return b.astype(np.uint8) * 255


def dilate(mask, fp, iterations=1):
    return _morph(mask, fp, True, iterations)


def erode(mask, fp, iterations=1):
    return _morph(mask, fp, False, iterations)


def morph_open(mask, fp):
    return dilate(erode(mask, fp), fp)


def morph_close(mask, fp):
    return erode(dilate(mask, fp), fp)


# ------------------------------------------------------------ Otsu / normalise
def otsu_threshold(gray: np.ndarray) -> int:
    """cv::getThreshVal_Otsu_8u: argmax of between-class variance over a 256-bin histogram
    (float64 accumulation, first maximum wins)."""
    h = np.bincount(gray.ravel(), minlength=256).astype(np.float64)
    n = gray.size
    scale = 1.0 / n
    mu = float((np.arange(256) * h).sum() * scale)
    mu1 = 0.0
    q1 = 0.0
    max_sigma = 0.0
    max_val = 0
    for i in range(256):
        p_i = h[i] * scale
        mu1 *= q1
        q1 += p_i
        q2 = 1.0 - q1
        if min(q1, q2) < np.finfo(np.float32).eps or max(q1, q2) > 1.0 - np.finfo(np.float32).eps:
            continue
        mu1 = (mu1 + i * p_i) / q1
        mu2 = (mu - q1 * mu1) / q2
        sigma = q1 * q2 * (mu1 - mu2) * (mu1 - mu2)
        if sigma > max_sigma:
            max_sigma = sigma
            max_val = i
    return max_val


def otsu_binary(gray: np.ndarray, object_type: str = "light") -> np.ndarray:
    """pcv.threshold.otsu(gray, object_type): THRESH_BINARY(_INV)+OTSU, maxval 255."""
    t = otsu_threshold(gray)
    m = gray > t if object_type == "light" else gray <= t
    return m.astype(np.uint8) * 255


def normalize_minmax_f32(a: np.ndarray, lo: float, hi: float) -> np.ndarray:
    """cv2.normalize(a, None, lo, hi, NORM_MINMAX) on float32: scale/shift computed in float64,
    applied by convertTo in float32 precision of a double alpha/beta."""
    a = a.astype(np.float32)
    smin = float(a.min())
    smax = float(a.max())
    dmin, dmax = min(lo, hi), max(lo, hi)
    scale = (dmax - dmin) * (1.0 / (smax - smin) if smax - smin > 2.220446049250313e-16 else 0.0)
    shift = dmin - smin * scale
    return (a.astype(np.float64) * scale + shift).astype(np.float32)


# ---------------------------------------------------------------- cv2.resize INTER_CUBIC / INTER_NEAREST (8-bit)
# _prepare_working_image (mask.py:29-50) upscales the image with INTER_CUBIC before masking and
# _resize_results_to_original (:526-545) brings the mask back with INTER_NEAREST.  OpenCV resize.cpp, 8-bit path:
# a = -0.75 cubic weights in float32, quantised to 11 bits (cvRound), int32 horizontal pass, vertical pass
# (sum + 2^21) >> 22 with saturation; taps outside the image replicate the border.  OpenCV's own SIMD and scalar code
# paths differ from each other by 1 LSB on a few per-mille of the values, so this spec is the +-1 LSB
# class, not bit-exact; INTER_NEAREST (sx = min(floor(dx * sw / dw), sw - 1)) is exact.
def cubic_taps(in_size: int, out_size: int):
    """-> (first source index [out] (tap k reads clip(s + k)), int32 weights [out, 4] x2048)."""
    scale = np.float64(in_size) / np.float64(out_size)
    d = np.arange(out_size, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int32)
    x = (f - s.astype(np.float32)).astype(np.float32)
    A = np.float32(-0.75)
    one = np.float32(1.0)
    c0 = ((A * (x + one) - np.float32(5) * A) * (x + one) + np.float32(8) * A) * (x + one) - np.float32(4) * A
    c1 = ((A + np.float32(2)) * x - (A + np.float32(3))) * x * x + one
    xm = one - x
    c2 = ((A + np.float32(2)) * xm - (A + np.float32(3))) * xm * xm + one
    c3 = one - c0 - c1 - c2
    c = np.stack([c0, c1, c2, c3], axis=1).astype(np.float32)
    w = np.rint(c * np.float32(2048)).astype(np.int32)          # saturate_cast<short>(float): round half to even
    return s - 1, w


def resize_cubic_u8(img: np.ndarray, out_wh) -> np.ndarray:
    """cv2.resize(img, (ow, oh), interpolation=cv2.INTER_CUBIC) for uint8 HxW[xC]."""
    ow, oh = int(out_wh[0]), int(out_wh[1])
    a = img if img.ndim == 3 else img[..., None]
    H, W = a.shape[:2]
    xs, xw = cubic_taps(W, ow)
    ys, yw = cubic_taps(H, oh)
    src = a.astype(np.int64)
    hp = np.zeros((H, ow, a.shape[2]), np.int64)
    for k in range(4):
        hp += src[:, np.clip(xs + k, 0, W - 1), :] * xw[None, :, k, None]
    out = np.zeros((oh, ow, a.shape[2]), np.int64)
    for k in range(4):
        out += hp[np.clip(ys + k, 0, H - 1)] * yw[:, k, None, None]
    out = np.clip((out + (1 << 21)) >> 22, 0, 255).astype(np.uint8)
    return out if img.ndim == 3 else out[..., 0]


def resize_nearest_u8(img: np.ndarray, out_wh) -> np.ndarray:
    """cv2.resize(img, (ow, oh), interpolation=cv2.INTER_NEAREST)."""
    ow, oh = int(out_wh[0]), int(out_wh[1])
    H, W = img.shape[:2]
    sx = np.minimum(np.floor(np.arange(ow) * (np.float64(W) / ow)).astype(np.int64), W - 1)
    sy = np.minimum(np.floor(np.arange(oh) * (np.float64(H) / oh)).astype(np.int64), H - 1)
    return img[sy][:, sx]
