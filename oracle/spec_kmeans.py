"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): NumPy restatement of the reference's k-means mask candidate.

Reference: `_create_kmeans_mask` srcs/transform/filters/mask.py:109-140 -- cv2.setRNGSeed(12345); INTER_AREA resize so that
the longer side is 256; cv2.kmeans(Z float32 [N,3], K=3, criteria (EPS+MAX_ITER, 20, 0.5), attempts=1, KMEANS_PP_CENTERS);
the cluster whose centre is green (H in green_hue_range and S >= 40), else by bg_bias, else the most saturated, becomes
the raw mask; INTER_NEAREST back to the image size.

The arithmetic is OpenCV's (third-party, not under /root/reference; opencv-python-headless 4.13.0, modules/core/src/
kmeans.cpp and rng: published algorithm restated here):
  * cv::RNG: state = (uint32)state * 4164903690 + (state >> 32) (multiply-with-carry), 32-bit outputs; (double)rng =
    ((next << 32) | next) * 2^-64.
  * generateCentersPP (k-means++ with 3 trials per centre): first centre = next % N; a candidate = first index whose running
    sum of squared distances reaches p = (double)rng * sum0; the trial with the smallest new total wins (first on ties).
    All distances here are integers < 2^18 (8-bit colours), every sum is exact, so the evaluation order is irrelevant.
  * Lloyd iterations: centre = (float32 sum of the members) * (1.f / count) -- the sums are integers < 2^24, exact in
    float32; distance = ((x0-c0)^2 + (x1-c1)^2) + (x2-c2)^2 in float32, each product and each sum rounded (hal::normL2Sqr_
    scalar tail, no FMA in the baseline build: checked against cv2 by tests/test_reference_differential.py); the first
    strictly smaller distance wins; the loop stops after the centre update of iteration 20 or when the largest squared
    centre shift (float64) is <= 0.25, and the labels of the LAST assignment are returned with the NEW centres.
  * an empty cluster takes the farthest member (last on ties) of the biggest cluster.
Pinned against cv2.kmeans itself in tests/test_reference_differential.py::test_kmeans_candidate (labels, centres, masks
bit-exact on 64 seeded leaves + degenerate images).
"""
from __future__ import annotations

import numpy as np

from . import spec_color as sc
from . import spec_mask as sm

CV_RNG_COEFF = 4164903690
K = 3
MAX_ITER = 20
EPS2 = 0.25          # criteria.epsilon ** 2
PP_TRIALS = 3


class CvRng:
    """cv::RNG (multiply-with-carry)."""

    def __init__(self, seed: int):
        self.state = seed & 0xFFFFFFFFFFFFFFFF

    def next(self) -> int:
        self.state = ((self.state & 0xFFFFFFFF) * CV_RNG_COEFF + (self.state >> 32)) & 0xFFFFFFFFFFFFFFFF
        return self.state & 0xFFFFFFFF

    def double(self) -> np.float64:
        t = self.next()
        u = (t << 32) | self.next()
        return np.float64(u) * np.float64(5.4210108624275221700372640043497e-20)


def _dist2_int(P: np.ndarray, q: np.ndarray) -> np.ndarray:
    d = P.astype(np.int64) - q.astype(np.int64)
    return (d * d).sum(axis=1)


def centers_pp(P: np.ndarray, rng: CvRng):
    """generateCentersPP on 8-bit points P [N,3]: indices of the K chosen points."""
    N = len(P)
    chosen = [rng.next() % N]
    dist = _dist2_int(P, P[chosen[0]])
    sum0 = int(dist.sum())
    for _k in range(1, K):
        best_sum, best_c, best_d = None, -1, None
        for _j in range(PP_TRIALS):
            p = rng.double() * np.float64(sum0)
            # sequential `p -= dist[ci]; if (p <= 0) break` over ci < N - 1: exact (integers against a double)
            pre = np.cumsum(dist[:N - 1].astype(np.float64))
            hit = np.nonzero(p - pre <= 0)[0]
            ci = int(hit[0]) if len(hit) else N - 1
            td = np.minimum(_dist2_int(P, P[ci]), dist)
            s = int(td.sum())
            if best_sum is None or s < best_sum:
                best_sum, best_c, best_d = s, ci, td
        chosen.append(best_c)
        sum0, dist = best_sum, best_d
    return chosen


def _dist2_f32(Z: np.ndarray, c: np.ndarray) -> np.ndarray:
    f = np.float32
    t0 = (Z[:, 0] - c[0]).astype(f)
    t1 = (Z[:, 1] - c[1]).astype(f)
    t2 = (Z[:, 2] - c[2]).astype(f)
    d = (t0 * t0).astype(f)
    d = (d + (t1 * t1).astype(f)).astype(f)
    d = (d + (t2 * t2).astype(f)).astype(f)
    return d


def _assign(Z: np.ndarray, centers: np.ndarray) -> np.ndarray:
    best = _dist2_f32(Z, centers[0]).astype(np.float64)
    lab = np.zeros(len(Z), np.int32)
    for k in range(1, K):
        d = _dist2_f32(Z, centers[k]).astype(np.float64)
        upd = best > d
        lab[upd] = k
        best = np.where(upd, d, best)
    return lab


def kmeans3(P: np.ndarray, seed: int = 12345):
    """cv2.kmeans(P.astype(float32), 3, None, (EPS+MAX_ITER, 20, 0.5), 1, KMEANS_PP_CENTERS) after cv2.setRNGSeed(seed).
    P: uint8 [N,3].  Returns (labels int32 [N], centers float32 [3,3], iterations)."""
    f = np.float32
    N = len(P)
    Z = P.astype(f)
    rng = CvRng(seed)
    centers = Z[centers_pp(P, rng)].copy()
    labels = _assign(Z, centers)
    it = 1
    while True:
        old = centers
        sums = np.zeros((K, 3), np.int64)
        counts = np.bincount(labels, minlength=K).astype(np.int64)
        for k in range(K):
            sums[k] = P[labels == k].astype(np.int64).sum(axis=0)
        fsum = sums.astype(f)                      # exact: < 2^24
        for k in range(K):
            if counts[k] != 0:
                continue
            max_k = 0
            for k1 in range(1, K):
                if counts[max_k] < counts[k1]:
                    max_k = k1
            base = (fsum[max_k] * (f(1.0) / f(counts[max_k]))).astype(f)
            idx = np.nonzero(labels == max_k)[0]
            d = _dist2_f32(Z[idx], base).astype(np.float64)
            far = int(idx[len(d) - 1 - int(np.argmax(d[::-1]))])     # `max_dist <= dist`: the last of the farthest
            counts[max_k] -= 1
            counts[k] += 1
            labels[far] = k
            fsum[max_k] = (fsum[max_k] - Z[far]).astype(f)
            fsum[k] = (fsum[k] + Z[far]).astype(f)
        centers = np.zeros((K, 3), f)
        shift = 0.0
        for k in range(K):
            centers[k] = (fsum[k] * (f(1.0) / f(counts[k]))).astype(f)
            t = centers[k].astype(np.float64) - old[k].astype(np.float64)
            shift = max(shift, float((t * t).sum()))
        it += 1
        if it == MAX_ITER or shift <= EPS2:
            return labels, centers, it
        labels = _assign(Z, centers)


def pick_cluster(centers: np.ndarray, cfg) -> int:
    """mask.py:123-136."""
    c8 = centers.astype(np.uint8)
    hsv_c = sc.rgb_to_hsv(c8.reshape(1, 3, 3))[0]
    lo, hi = cfg.green_hue_range
    green = np.array([1 if (lo <= int(hv[0]) <= hi and int(hv[1]) >= 40) else 0 for hv in hsv_c])
    bias = (cfg.bg_bias or "auto").lower()
    if bias == "dark_bg":
        return int(np.argmax(c8.mean(axis=1)))
    if bias == "light_bg":
        return int(np.argmin(c8.mean(axis=1)))
    if green.any():
        return int(np.argmax(green))
    return int(np.argmax(hsv_c[:, 1]))


def resize_nearest(img: np.ndarray, nw: int, nh: int) -> np.ndarray:
    """cv2.resize(..., INTER_NEAREST): sx = min(floor(dx * (w / nw)), w - 1) with a double scale."""
    h, w = img.shape[:2]
    xs = np.minimum(np.floor(np.arange(nw) * (np.float64(w) / nw)).astype(np.int64), w - 1)
    ys = np.minimum(np.floor(np.arange(nh) * (np.float64(h) / nh)).astype(np.int64), h - 1)
    return img[ys][:, xs]


def kmeans_mask(rgb: np.ndarray, cfg, return_details: bool = False):
    """_create_kmeans_mask (mask.py:109-140)."""
    h, w = rgb.shape[:2]
    scale = 256 / max(h, w)
    sw, sh = max(1, int(w * scale)), max(1, int(h * scale))
    if (sw, sh) == (w, h):
        small = rgb
    elif sw >= w and sh >= h:
        small = sm.resize_area_up(rgb, sw, sh)
    else:
        small = sm.resize_area_down(rgb, sw, sh)
    labels, centers, it = kmeans3(small.reshape(-1, 3))
    pick = pick_cluster(centers, cfg)
    ms = ((labels.reshape(sh, sw) == pick).astype(np.uint8)) * 255
    m = ms if (sw, sh) == (w, h) else resize_nearest(ms, w, h)
    return (m, labels.reshape(sh, sw), centers, pick, it) if return_details else m
