"""Batched device ops: thin torch-tensor wrappers over the C ABI (include/leafx.h).

torch is plumbing only (device memory, streams); every op below launches the hand-written
sm_100a kernels in libleafx.so on the current CUDA stream.  Inputs are CUDA uint8 tensors
[B,H,W,3]; per-image parameters are host arrays uploaded here.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import MaskCfg

STRATEGY_IDS = {"hsv_h": 0, "lab": 1, "hsv_s": 2, "hsv_v_dark": 3, "external": 4}


def _ready(x: torch.Tensor):
    if not x.is_cuda:
        raise _lib.LeafxError(_lib.ERR_CUDA, "leaffliction_b200 ops need CUDA tensors (no CPU fallback)")
    return _lib.init(x.device.index if x.device.index is not None else torch.cuda.current_device())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t: Optional[torch.Tensor]):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _chk_img(x: torch.Tensor, ch: Optional[int] = 3):
    if x.dtype != torch.uint8 or not x.is_contiguous():
        raise ValueError("expected a contiguous uint8 tensor")
    if ch is not None and (x.dim() != 4 or x.shape[-1] != ch):
        raise ValueError(f"expected shape [B,H,W,{ch}], got {tuple(x.shape)}")


def _dev(a, dtype, device):
    """Host array / list -> device tensor; a tensor already on the device passes through (hot loops
    pre-upload their per-image parameters once)."""
    if isinstance(a, torch.Tensor) and a.device == device:
        return a
    return torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).to(device)


def cvt_color(x: torch.Tensor, code: str) -> torch.Tensor:
    """cv2.cvtColor(rgb, COLOR_RGB2{GRAY,HSV,LAB}); code in {'gray','hsv','lab'}."""
    _chk_img(x)
    lib = _ready(x)
    B, H, W, _ = x.shape
    ci = {"gray": 0, "hsv": 1, "lab": 2}[code]
    out = torch.empty((B, H, W) if ci == 0 else (B, H, W, 3), dtype=torch.uint8, device=x.device)
    _lib.check(lib.lfx_cvt_color(_p(x), _p(out), B, H, W, ci, _stream()))
    return out


def mask_cfg(strategy="hsv_h", green_hue_range=(25, 100), fill_size=1000, morph_kernel=3,
             brown_hue_range=(0, 30), brown_s_min=20, brown_v_max=200, brown_min_area_px=25,
             brown_morph_kernel=3, use_lab_brown=False, lab_a_min=125, lab_b_min=125,
             hsv_channel_for_mask="s", bg_bias="light_bg", extend_brown=True) -> MaskCfg:
    c = MaskCfg()
    c.strategy = STRATEGY_IDS[strategy] if isinstance(strategy, str) else int(strategy)
    c.green_lo, c.green_hi = int(green_hue_range[0]), int(green_hue_range[1])
    c.fill_size, c.morph_kernel = int(fill_size), int(morph_kernel)
    c.brown_lo, c.brown_hi = int(brown_hue_range[0]), int(brown_hue_range[1])
    c.brown_s_min, c.brown_v_max = int(brown_s_min), int(brown_v_max)
    c.brown_min_area_px, c.brown_morph_kernel = int(brown_min_area_px), int(brown_morph_kernel)
    c.use_lab_brown, c.lab_a_min, c.lab_b_min = int(bool(use_lab_brown)), int(lab_a_min), int(lab_b_min)
    c.fallback_channel = "hsv".index(str(hsv_channel_for_mask))
    c.bg_dark = int((bg_bias or "auto").lower() == "dark_bg")
    c.extend_brown = int(bool(extend_brown))
    return c


def threshold_mask(x: torch.Tensor, cfg: MaskCfg) -> torch.Tensor:
    _chk_img(x)
    lib = _ready(x)
    B, H, W, _ = x.shape
    out = torch.empty((B, H, W), dtype=torch.uint8, device=x.device)
    _lib.check(lib.lfx_threshold_mask(_p(x), _p(out), B, H, W, C.byref(cfg), _stream()))
    return out


_ws_cache = {}


def _workspace(nbytes: int, device) -> torch.Tensor:
    key = (device.index, torch.cuda.current_stream().cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


def make_mask(x: torch.Tensor, cfg: MaskCfg, raw: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Batched make_mask (mask.py:548-582, profile P0/P1). Returns (mask [B,H,W] u8, info [B,8] i32)."""
    _chk_img(x)
    lib = _ready(x)
    B, H, W, _ = x.shape
    mask = torch.empty((B, H, W), dtype=torch.uint8, device=x.device)
    info = torch.empty((B, 8), dtype=torch.int32, device=x.device)
    nws = lib.lfx_make_mask_workspace(B, H, W)
    ws = _workspace(nws, x.device)
    _lib.check(lib.lfx_make_mask(_p(x), _p(raw), _p(mask), _p(info), B, H, W, C.byref(cfg), _p(ws), ws.numel(), _stream()))
    return mask, info


def postprocess_mask(raw: torch.Tensor, fill_size=1000, morph_kernel=3):
    _chk_img(raw, None)
    lib = _ready(raw)
    B, H, W = raw.shape
    mask = torch.empty_like(raw)
    info = torch.empty((B, 8), dtype=torch.int32, device=raw.device)
    ws = _workspace(lib.lfx_make_mask_workspace(B, H, W), raw.device)
    _lib.check(lib.lfx_postprocess_mask(_p(raw), _p(mask), _p(info), B, H, W, int(fill_size), int(morph_kernel),
                                        _p(ws), ws.numel(), _stream()))
    return mask, info


def canny(gray: torch.Tensor, low: float, high: float, l2: bool = False) -> torch.Tensor:
    """cv2.Canny(gray, low, high, L2gradient=l2) on [B,H,W] uint8."""
    _chk_img(gray, None)
    lib = _ready(gray)
    B, H, W = gray.shape
    out = torch.empty_like(gray)
    ws = _workspace(lib.lfx_front_workspace(B, H, W), gray.device)
    _lib.check(lib.lfx_canny(_p(gray), _p(out), B, H, W, float(low), float(high), int(bool(l2)), _p(ws), ws.numel(), _stream()))
    return out


def raw_mask_front_end(x: torch.Tensor, which: str, cfg: MaskCfg) -> torch.Tensor:
    """Raw candidate of the 'inclusive' / 'enhanced' strategies (mask.py:727-831, :610-724)."""
    _chk_img(x)
    lib = _ready(x)
    B, H, W, _ = x.shape
    out = torch.empty((B, H, W), dtype=torch.uint8, device=x.device)
    ws = _workspace(lib.lfx_front_workspace(B, H, W), x.device)
    _lib.check(lib.lfx_raw_mask(_p(x), _p(out), B, H, W, {"inclusive": 0, "enhanced": 1}[which], C.byref(cfg),
                                _p(ws), ws.numel(), _stream()))
    return out


def brown_spots(x: torch.Tensor, mask: torch.Tensor, cfg: MaskCfg):
    """apply_brown_filter core: (spots [B,H,W] u8, stats [B,4] = leaf_px, count, spot_px, status)."""
    _chk_img(x)
    lib = _ready(x)
    B, H, W, _ = x.shape
    out = torch.empty((B, H, W), dtype=torch.uint8, device=x.device)
    stats = torch.empty((B, 4), dtype=torch.int32, device=x.device)
    ws = _workspace(lib.lfx_front_workspace(B, H, W), x.device)
    _lib.check(lib.lfx_brown_spots(_p(x), _p(mask), _p(out), _p(stats), B, H, W, C.byref(cfg), _p(ws), ws.numel(), _stream()))
    return out, stats


def saliency_blur(x: torch.Tensor, mask: torch.Tensor, cfg: MaskCfg, gaussian_sigma: float = 1.5) -> torch.Tensor:
    """apply_blur_filter given the leaf mask of make_mask(x): [B,H,W,3] u8."""
    _chk_img(x)
    lib = _ready(x)
    B, H, W, _ = x.shape
    out = torch.empty_like(x)
    ws = _workspace(lib.lfx_front_workspace(B, H, W), x.device)
    _lib.check(lib.lfx_saliency_blur(_p(x), _p(mask), _p(out), B, H, W, float(gaussian_sigma), C.byref(cfg),
                                     _p(ws), ws.numel(), _stream()))
    return out


def trace_contour(mask: torch.Tensor, info: torch.Tensor, max_pts: int = 4096):
    """External contour of the selected component (cv2.findContours RETR_EXTERNAL/CHAIN_APPROX_SIMPLE).
    Returns (points [B,max_pts,2] i32, counts [B] i32, sums [B,3] i64)."""
    _chk_img(mask, None)
    lib = _ready(mask)
    B, H, W = mask.shape
    pts = torch.empty((B, max_pts, 2), dtype=torch.int32, device=mask.device)
    cnt = torch.empty((B,), dtype=torch.int32, device=mask.device)
    sums = torch.empty((B, 3), dtype=torch.int64, device=mask.device)
    _lib.check(lib.lfx_trace_contour(_p(mask), _p(info), _p(pts), _p(cnt), _p(sums), B, H, W, int(max_pts), _stream()))
    return pts, cnt, sums


def analyze_records(mask: torch.Tensor, info: torch.Tensor, max_pts: int = 4096, max_hull: int = 512):
    """Batched numeric record of apply_analyze_filter (analyze.py:43-98) for the contours selected by make_mask:
    trace (lfx_trace_contour) + record (lfx_analyze_record), nothing leaves the device.
    Returns dict(points [B,max_pts,2], counts [B], rec_i [B,24] i32, rec_f [B,12] f64, hull [B,max_hull,2] i32);
    see include/leafx.h for the field layout."""
    pts, cnt, sums = trace_contour(mask, info, max_pts)
    lib = _ready(mask)
    B, H, W = mask.shape
    rec_i = torch.empty((B, 24), dtype=torch.int32, device=mask.device)
    rec_f = torch.empty((B, 12), dtype=torch.float64, device=mask.device)
    hull = torch.zeros((B, max_hull, 2), dtype=torch.int32, device=mask.device)
    ws = _workspace(lib.lfx_analyze_workspace(B, H), mask.device)
    _lib.check(lib.lfx_analyze_record(_p(pts), _p(cnt), _p(sums), _p(rec_i), _p(rec_f), _p(hull), B, H, W, int(max_pts), int(max_hull),
                                      _p(ws), ws.numel(), _stream()))
    return dict(points=pts, counts=cnt, rec_i=rec_i, rec_f=rec_f, hull=hull)


def analyze_points(points: torch.Tensor, counts: torch.Tensor, H: int, W: int, max_hull: int = 512):
    """lfx_analyze_record on contour points that are already known (points int32 [B,max_pts,2], counts int32 [B])."""
    lib = _ready(points)
    B, max_pts = int(points.shape[0]), int(points.shape[1])
    rec_i = torch.empty((B, 24), dtype=torch.int32, device=points.device)
    rec_f = torch.empty((B, 12), dtype=torch.float64, device=points.device)
    hull = torch.zeros((B, max_hull, 2), dtype=torch.int32, device=points.device)
    ws = _workspace(lib.lfx_analyze_workspace(B, H), points.device)
    _lib.check(lib.lfx_analyze_record(_p(points), _p(counts), None, _p(rec_i), _p(rec_f), _p(hull), B, int(H), int(W), max_pts, int(max_hull),
                                      _p(ws), ws.numel(), _stream()))
    return dict(points=points, counts=counts, rec_i=rec_i, rec_f=rec_f, hull=hull)


def analyze_overlay(rgb: torch.Tensor, rec: dict, edges: torch.Tensor = None, mask: torch.Tensor = None) -> torch.Tensor:
    """The overlay image apply_analyze_filter returns (analyze.py:37-122) for a batch: rgb [B,H,W,3] u8 + the record of
    analyze_records / analyze_points (points, counts, rec_i, hull) -> overlay [B,H,W,3], bit-identical to the OpenCV drawing
    calls (contour, centroid cross, extreme-point circles and anti-aliased rays, anti-aliased hull, PCA axes); with `edges`
    (canny(gray, 80, 160, True)) and `mask` the vein pixels edges & mask are painted cyan.  Images without a contour are
    copied.  Raises when a hull did not fit max_hull (rec_i[:, 12] < 0): re-run the record with a larger max_hull."""
    _chk_img(rgb)
    lib = _ready(rgb)
    B, H, W, _ = rgb.shape
    pts, cnt, rec_i, hull = rec["points"], rec["counts"], rec["rec_i"], rec["hull"]
    for t, dt in ((pts, torch.int32), (cnt, torch.int32), (rec_i, torch.int32), (hull, torch.int32)):
        if not t.is_cuda or t.dtype != dt or not t.is_contiguous() or int(t.shape[0]) != B:
            raise ValueError("analyze_overlay: record tensors must be contiguous int32 CUDA tensors of the batch")
    if (edges is None) != (mask is None):
        raise ValueError("analyze_overlay: edges and mask go together")
    if edges is not None:
        _chk_img(edges, None)
        _chk_img(mask, None)
        if tuple(edges.shape) != (B, H, W) or tuple(mask.shape) != (B, H, W):
            raise ValueError("analyze_overlay: edges / mask must be [B,H,W]")
    if bool((rec_i[:, 12] < 0).any()):
        raise ValueError("analyze_overlay: a hull did not fit max_hull; run the record again with a larger max_hull")
    out = torch.empty_like(rgb)
    _lib.check(lib.lfx_analyze_overlay(_p(rgb), _p(pts), _p(cnt), _p(rec_i), _p(hull), _p(edges) if edges is not None else None,
                                       _p(mask) if mask is not None else None, _p(out), B, H, W, int(pts.shape[1]), int(hull.shape[1]),
                                       _stream()))
    return out


def draw_rectangles(rgb: torch.Tensor, info: torch.Tensor, color=(255, 0, 0), thickness: int = 2) -> torch.Tensor:
    """`vis` of apply_roi_filter (roi.py:43-44) for a batch: cv2.rectangle(vis, (x, y), (x + w, y + h), color, thickness) with
    info [B,8] i32 = {found, x, y, w, h, ...} (make_mask's layout); bit-identical to OpenCV."""
    _chk_img(rgb)
    lib = _ready(rgb)
    B, H, W, _ = rgb.shape
    if not info.is_cuda or info.dtype != torch.int32 or not info.is_contiguous() or tuple(info.shape) != (B, 8):
        raise ValueError("draw_rectangles: info must be a contiguous int32 CUDA tensor [B,8]")
    out = torch.empty_like(rgb)
    col = int(color[0]) | (int(color[1]) << 8) | (int(color[2]) << 16)
    _lib.check(lib.lfx_draw_rectangles(_p(rgb), _p(info), _p(out), B, H, W, col, int(thickness), _stream()))
    return out


DRAW_LINE, DRAW_LINE_AA, DRAW_CIRCLE_FILLED, DRAW_RECTANGLE, DRAW_MARKER_CROSS = 1, 2, 3, 4, 5


def draw_primitives(img: torch.Tensor, prims, counts=None) -> torch.Tensor:
    """OpenCV's rasterisers as a general op, IN PLACE on img [B,H,W,3] u8: prims [B,P,8] int32 rows
    (kind, x0, y0, x1, y1, r | g << 8 | b << 16, size, 0) drawn in list order per image, each bit-identical to the cv2 call
    (DRAW_LINE = cv2.line thickness `size` >= 2; DRAW_LINE_AA = cv2.line 1 px LINE_AA; DRAW_CIRCLE_FILLED = cv2.circle radius
    `size` filled; DRAW_RECTANGLE = cv2.rectangle thickness `size` >= 2; DRAW_MARKER_CROSS = cv2.drawMarker cross, markerSize
    x1, thickness `size`).  counts [B] int32 (default: every row of prims).  Returns img."""
    _chk_img(img)
    lib = _ready(img)
    B, H, W, _ = img.shape
    prims = _dev(prims, np.int32, img.device)
    if prims.dtype != torch.int32 or prims.dim() != 3 or int(prims.shape[0]) != B or int(prims.shape[2]) != 8 or not prims.is_contiguous():
        raise ValueError("draw_primitives: prims must be a contiguous int32 tensor [B,P,8]")
    P = int(prims.shape[1])
    counts = torch.full((B,), P, dtype=torch.int32, device=img.device) if counts is None else _dev(counts, np.int32, img.device)
    if counts.dtype != torch.int32 or tuple(counts.shape) != (B,):
        raise ValueError("draw_primitives: counts must be int32 [B]")
    if P:
        _lib.check(lib.lfx_draw_primitives(_p(img), _p(prims), _p(counts), B, H, W, P, _stream()))
    return img


def strategy_raw(x: torch.Tensor, cfg: MaskCfg) -> torch.Tensor:
    """Raw candidate of a threshold strategy (cfg.strategy 0-3), no post-processing: [B,H,W] u8 (mask.py:72-106)."""
    _chk_img(x)
    lib = _ready(x)
    B, H, W, _ = x.shape
    out = torch.empty((B, H, W), dtype=torch.uint8, device=x.device)
    ws = _workspace(lib.lfx_make_mask_workspace(B, H, W), x.device)
    _lib.check(lib.lfx_strategy_raw(_p(x), _p(out), B, H, W, C.byref(cfg), _p(ws), ws.numel(), _stream()))
    return out


def seed_words(dseeds: torch.Tensor, nwords: int = 16, out: torch.Tensor = None) -> torch.Tensor:
    """First `nwords` 32-bit outputs of Python's `random` after random.seed(s) for every task seed s (int32 / uint32 bits,
    device tensor [B]) -> int32 [B, nwords] on the device (lfx_seed_words: CPython's init_by_array seeding, one thread per
    task)."""
    if not dseeds.is_cuda or dseeds.dtype != torch.int32 or not dseeds.is_contiguous():
        raise ValueError("seed_words: contiguous int32 CUDA tensor expected")
    lib = _ready(dseeds)
    B = dseeds.numel()
    if out is None:
        out = torch.empty((B, nwords), dtype=torch.int32, device=dseeds.device)
    _lib.check(lib.lfx_seed_words(_p(dseeds), B, int(nwords), _p(out), _stream()))
    return out


KMEANS_BIAS = {"auto": 0, "dark_bg": 1, "light_bg": 2}


def kmeans_raw(x: torch.Tensor, green_hue_range=(25, 100), bg_bias: str = "auto", seed: int = 12345, details: bool = False):
    """`_create_kmeans_mask` (mask.py:109-140) on a device batch whose longer side is 256: raw candidate [B,H,W] u8
    (0/255); with `details` also (centers float32 [B,3,3], kinfo int32 [B,4] = picked cluster, iterations, empty-cluster
    events, points).  cv2.kmeans / cv::RNG restated exactly (lfx_kmeans.cu)."""
    _chk_img(x)
    lib = _ready(x)
    B, H, W, _ = x.shape
    out = torch.empty((B, H, W), dtype=torch.uint8, device=x.device)
    cen = torch.empty((B, 3, 3), dtype=torch.float32, device=x.device) if details else None
    ki = torch.empty((B, 4), dtype=torch.int32, device=x.device) if details else None
    bias = KMEANS_BIAS.get((bg_bias or "auto").lower(), 0)
    _lib.check(lib.lfx_kmeans_raw(_p(x), _p(out), _p(cen) if details else None, _p(ki) if details else None, B, H, W,
                                  int(green_hue_range[0]), int(green_hue_range[1]), bias, int(seed) & 0xFFFFFFFF, _stream()))
    return (out, cen, ki) if details else out


def score_features(x: torch.Tensor, masks: torch.Tensor, green_hue_range=(25, 100)):
    """Image-dependent terms of _score_mask (mask.py:160-177) for K candidate masks [K,B,H,W] of the images x [B,H,W,3].
    Returns (feat [K,B,4] f64 = boundary |grad| sum, boundary px, mask px, green & mask px; gmax [B] f32; gmin [B] f32)."""
    _chk_img(x)
    lib = _ready(x)
    B, H, W, _ = x.shape
    if masks.dtype != torch.uint8 or masks.dim() != 4 or tuple(masks.shape[1:]) != (B, H, W) or not masks.is_contiguous():
        raise ValueError("score_features: masks must be a contiguous uint8 [K,B,H,W] tensor")
    K = masks.shape[0]
    feat = torch.empty((K, B, 4), dtype=torch.float64, device=x.device)
    mm = torch.empty((B, 2), dtype=torch.int32, device=x.device)
    _lib.check(lib.lfx_score_features(_p(x), _p(masks), _p(feat), _p(mm), B, H, W, K, int(green_hue_range[0]), int(green_hue_range[1]),
                                      _stream()))
    gmax = mm[:, 0].contiguous().view(torch.float32)
    gmin = (~mm[:, 1]).contiguous().view(torch.float32)
    return feat, gmax, gmin


def apply_mask(x: torch.Tensor, mask: torch.Tensor, color_val: int = 255) -> torch.Tensor:
    _chk_img(x)
    lib = _ready(x)
    B, H, W, _ = x.shape
    out = torch.empty_like(x)
    _lib.check(lib.lfx_apply_mask(_p(x), _p(mask), _p(out), B, H, W, int(color_val), _stream()))
    return out


def gauss_taps(ksize: int, sigma: float) -> np.ndarray:
    lib = _lib.load()
    out = np.zeros(ksize, np.int32)
    _lib.check(lib.lfx_gauss_taps(int(ksize), float(sigma), out.ctypes.data_as(C.c_void_p)))
    return out


def gauss_u8(x: torch.Tensor, ksize: int, sigma: float = 0.0) -> torch.Tensor:
    """cv2.GaussianBlur(x,(k,k),sigma) on [B,H,W,3] or [B,H,W] uint8."""
    if x.dtype != torch.uint8 or not x.is_contiguous() or x.dim() not in (3, 4):
        raise ValueError("expected contiguous uint8 [B,H,W] or [B,H,W,3]")
    lib = _ready(x)
    B, H, W = x.shape[:3]
    ch = 1 if x.dim() == 3 else x.shape[3]
    out = torch.empty_like(x)
    _lib.check(lib.lfx_gauss_u8(_p(x), _p(out), B, H, W, ch, int(ksize), float(sigma), _stream()))
    return out


def roi_letterbox(x: torch.Tensor, mask: Optional[torch.Tensor], info: torch.Tensor, roi_size=(256, 256)) -> torch.Tensor:
    _chk_img(x)
    lib = _ready(x)
    B, H, W, _ = x.shape
    RH, RW = int(roi_size[0]), int(roi_size[1])
    out = torch.empty((B, RH, RW, 3), dtype=torch.uint8, device=x.device)
    _lib.check(lib.lfx_roi_letterbox(_p(x), _p(mask), _p(info), _p(out), B, H, W, RH, RW, _stream()))
    return out


def color_stats(x: torch.Tensor, mask: Optional[torch.Tensor], hist9=True, hsv3=True, counters=True):
    """Returns (hist9 [B,9,256] | None, hsv3 [B,3,256] | None, counters [B,16] | None) int32."""
    _chk_img(x)
    lib = _ready(x)
    B, H, W, _ = x.shape
    h9 = torch.zeros((B, 9, 256), dtype=torch.int32, device=x.device) if hist9 else None
    h3 = torch.zeros((B, 3, 256), dtype=torch.int32, device=x.device) if hsv3 else None
    cn = torch.zeros((B, 16), dtype=torch.int32, device=x.device) if counters else None
    _lib.check(lib.lfx_color_stats(_p(x), _p(mask), _p(h9), _p(h3), _p(cn), B, H, W, _stream()))
    return h9, h3, cn


@dataclass
class CoreOutputs:
    blur: torch.Tensor
    mask: torch.Tensor
    info: torch.Tensor
    roi: torch.Tensor
    hist9: torch.Tensor
    hsv3: torch.Tensor
    counters: torch.Tensor


def alloc_core_outputs(B, H, W, roi_size, device) -> CoreOutputs:
    u8 = dict(dtype=torch.uint8, device=device)
    i32 = dict(dtype=torch.int32, device=device)
    return CoreOutputs(torch.empty((B, H, W, 3), **u8), torch.empty((B, H, W), **u8), torch.empty((B, 8), **i32),
                       torch.empty((B, roi_size[0], roi_size[1], 3), **u8), torch.empty((B, 9, 256), **i32),
                       torch.empty((B, 3, 256), **i32), torch.empty((B, 16), **i32))


def pipeline_core(x: torch.Tensor, cfg: MaskCfg, gaussian_sigma: float = 1.5, roi_size=(256, 256),
                  out: Optional[CoreOutputs] = None, dataset_hist: Optional[torch.Tensor] = None,
                  raw: Optional[torch.Tensor] = None) -> CoreOutputs:
    """Core transform profile (BASELINE config 2): 5x5 Gaussian blur + make_mask + masked ROI
    letterbox + RGB/HSV/LAB histograms and hist.py counters, one submission.
    `dataset_hist` (int64 [9,256] on the device): the batch's histograms are added to it by the kernel."""
    if dataset_hist is not None and (dataset_hist.dtype != torch.int64 or dataset_hist.numel() != 9 * 256
                                     or not dataset_hist.is_contiguous() or dataset_hist.device != x.device):
        raise ValueError("dataset_hist must be a contiguous int64 [9,256] tensor on the images' device")
    _chk_img(x)
    lib = _ready(x)
    B, H, W, _ = x.shape
    if out is None:
        out = alloc_core_outputs(B, H, W, roi_size, x.device)
    ws = _workspace(lib.lfx_pipeline_core_workspace(B, H, W), x.device)
    _lib.check(lib.lfx_pipeline_core(_p(x), _p(out.blur), _p(out.mask), _p(out.info), _p(out.roi), _p(out.hist9),
                                     _p(out.hsv3), _p(out.counters), B, H, W, int(roi_size[0]), int(roi_size[1]),
                                     float(gaussian_sigma), C.byref(cfg), _p(ws), ws.numel(), _p(dataset_hist), _p(raw), _stream()))
    return out


def pipeline_front(x: torch.Tensor, which: str, cfg: MaskCfg, gaussian_sigma: float = 1.5, roi_size=(256, 256),
                   out: Optional[CoreOutputs] = None, dataset_hist: Optional[torch.Tensor] = None, bg_bias: str = "light_bg") -> CoreOutputs:
    """The core transform profile with the reference's DEFAULT mask strategies ('inclusive', config.yaml:7, or 'enhanced'), or
    with the k-means candidate ('kmeans', mask.py:109-140; images whose longer side is 256): raw candidate by the front-end
    kernel (lfx_raw_mask / lfx_kmeans_raw), then make_mask on that candidate (strategy 4), 5x5 blur, masked ROI letterbox and
    colour statistics -- the same outputs as pipeline_core."""
    import copy
    _chk_img(x)
    raw = kmeans_raw(x, (cfg.green_lo, cfg.green_hi), bg_bias) if which == "kmeans" else raw_mask_front_end(x, which, cfg)
    cfg4 = copy.copy(cfg)
    cfg4.strategy = STRATEGY_IDS["external"]
    # two launches: the front-end kernel, then the fused core kernel on its candidate (strategy 4)
    return pipeline_core(x, cfg4, gaussian_sigma, roi_size, out, dataset_hist, raw)


# --------------------------------------------------------------------------- augmentations
# Every augment op takes `src_index` (device int32 [B], optional): output i is computed from x[src_index[i]] -- the
# balancer's random.choice(source_images) (dataset_balancer.py:116) read in place, no gather copy -- and `out`
# (preallocated output, optional) so that hot loops neither allocate nor copy.
def _src(x: torch.Tensor, src_index):
    """(B, index pointer, number of source images)."""
    if src_index is None:
        return x.shape[0], None, x.shape[0]
    if src_index.dtype != torch.int32 or src_index.device != x.device or not src_index.is_contiguous():
        raise ValueError("src_index must be a contiguous int32 tensor on the images' device")
    return int(src_index.numel()), src_index, x.shape[0]


def _out_like(x, B, out, shape=None):
    shape = tuple(shape) if shape is not None else (B,) + tuple(x.shape[1:])
    if out is None:
        return torch.empty(shape, dtype=torch.uint8, device=x.device)
    if tuple(out.shape) != shape or out.dtype != torch.uint8 or not out.is_contiguous() or out.device != x.device:
        raise ValueError(f"out must be a contiguous uint8 tensor of shape {shape}")
    return out


def flip(x: torch.Tensor, left_right: Sequence[bool], src_index=None, out=None) -> torch.Tensor:
    _chk_img(x)
    lib = _ready(x)
    _, H, W, _c = x.shape
    B, sidx, nsrc = _src(x, src_index)
    if isinstance(left_right, torch.Tensor):
        mode = left_right                                            # lfx_flip modes already on the device
    elif isinstance(left_right, np.ndarray) and left_right.dtype == np.int32:
        mode = _dev(left_right, np.int32, x.device)                 # already lfx_flip modes (0 = left-right)
    else:
        mode = _dev([0 if lr else 1 for lr in left_right], np.int32, x.device)
    out = _out_like(x, B, out)
    _lib.check(lib.lfx_flip(_p(x), _p(out), B, H, W, _p(mode), _p(sidx), nsrc, _stream()))
    return out


def rotate_nn(x: torch.Tensor, params: np.ndarray, fill: int = 255, dparams: torch.Tensor = None, out: torch.Tensor = None,
              src_index=None):
    """params[B][8] = a0..a5 (16.16 fixed point), nw, nh.  Returns (slab [B, stride] u8, stride):
    image i is slab[i, : nh_i*nw_i*3].view(nh_i, nw_i, 3).  `dparams` (the same table already on the device) and
    `out` (a slab of at least that stride) let a hot loop skip the upload and the allocation."""
    _chk_img(x)
    lib = _ready(x)
    _, H, W, _c = x.shape
    B, sidx, nsrc = _src(x, src_index)
    params = np.ascontiguousarray(params, np.int32).reshape(B, 8)
    max_px = int((params[:, 6].astype(np.int64) * params[:, 7]).max()) if B else 0
    stride = ((max_px * 3 + 15) // 16) * 16
    slab = out if out is not None else torch.empty((B, max(stride, 16)), dtype=torch.uint8, device=x.device)
    if slab.shape[0] != B or slab.shape[1] < max(stride, 16) or slab.dtype != torch.uint8:
        raise ValueError("rotate_nn: output slab too small")
    dp = dparams if dparams is not None else _dev(params, np.int32, x.device)
    _lib.check(lib.lfx_rotate_nn(_p(x), _p(slab), slab.shape[1], B, H, W, _p(dp), int(fill), _p(sidx), nsrc, _stream()))
    return slab, slab.shape[1]


def warp_bicubic(x: torch.Tensor, coeffs: np.ndarray, perspective: Sequence[bool], src_index=None, out=None) -> torch.Tensor:
    _chk_img(x)
    lib = _ready(x)
    _, H, W, _c = x.shape
    B, sidx, nsrc = _src(x, src_index)
    dc = coeffs if isinstance(coeffs, torch.Tensor) else _dev(np.asarray(coeffs, np.float64).reshape(B, 8), np.float64, x.device)
    dpz = perspective if isinstance(perspective, torch.Tensor) else _dev(np.asarray(perspective).astype(bool).astype(np.int32), np.int32, x.device)
    out = _out_like(x, B, out)
    _lib.check(lib.lfx_warp_bicubic(_p(x), _p(out), B, H, W, _p(dc), _p(dpz), _p(sidx), nsrc, _stream()))
    return out


class LanczosTables:
    """Host cache of Pillow's fixed-point Lanczos coefficient tables, concatenated for upload."""

    def __init__(self):
        self.rows = {}      # (in,out) -> (row offset, ksize)
        self.bounds = []
        self.kk = []
        self.kstride = 0
        self.nrows = 0
        self._dev = None
        self._retired = []   # superseded device copies: kernels already queued on other streams may still read them

    def prefill(self, in_lo: int, in_hi: int, out_size: int):
        """Tables for every source size in [in_lo, in_hi] at once (ImageAugmenter.crop draws nw = int(W * r), r in [0.8, 0.95]):
        the device copy then never changes while crops of that image size are in flight."""
        for sz in range(max(1, int(in_lo)), int(in_hi) + 1):
            self.get(sz, out_size)

    def _ensure_stride(self, ks):
        if ks > self.kstride:
            new = ((ks + 7) // 8) * 8
            self.kk = [np.pad(k, ((0, 0), (0, new - k.shape[1]))) for k in self.kk]
            self.kstride = new
            if self._dev is not None:
                self._retired.append(self._dev)
                self._dev = None

    def get(self, in_size: int, out_size: int):
        key = (int(in_size), int(out_size))
        if key not in self.rows:
            lib = _lib.load()
            ks = lib.lfx_lanczos_ksize(*key)
            if ks < 0:
                raise ValueError(f"bad Lanczos sizes {key}")
            self._ensure_stride(ks)
            b = np.zeros((key[1], 2), np.int32)
            k = np.zeros((key[1], self.kstride), np.int32)
            rc = lib.lfx_lanczos_table(key[0], key[1], self.kstride, b.ctypes.data_as(C.c_void_p), k.ctypes.data_as(C.c_void_p))
            if rc < 0:
                _lib.check(rc)
            self.rows[key] = (self.nrows, ks)
            self.bounds.append(b)
            self.kk.append(k)
            self.nrows += key[1]
        return self.rows[key]

    def device(self, device):
        if self._dev is None or self._dev[0].device != device or self._dev[2] != self.nrows:
            if self._dev is not None:
                self._retired.append(self._dev)      # never handed back to the allocator while a queued kernel may use it
                del self._retired[:-64]
            self._dev = (_dev(np.concatenate(self.bounds), np.int32, device), _dev(np.concatenate(self.kk), np.int32, device), self.nrows)
        return self._dev[0], self._dev[1]


_lanczos = LanczosTables()


class CropPlan:
    """Device-side parameters of one crop_lanczos batch (boxes, table offsets), built once and reusable."""

    def __init__(self, boxes: np.ndarray, out_hw: Tuple[int, int], device, upload: bool = True):
        """`upload=False` leaves `.box` / `.off` unset: the caller uploads `.h_box` / `.h_off` itself (packed transfers)."""
        OH, OW = int(out_hw[0]), int(out_hw[1])
        boxes = np.ascontiguousarray(boxes, np.int32).reshape(-1, 4)
        off = np.zeros((len(boxes), 4), np.int32)
        for col, osz, c0 in ((2, OW, 0), (3, OH, 2)):               # one table per distinct source size
            sizes, inv = np.unique(boxes[:, col], return_inverse=True)
            rows = np.array([_lanczos.get(int(sz), osz) for sz in sizes], np.int32).reshape(-1, 2)
            off[:, c0:c0 + 2] = rows[inv.reshape(-1)]
        self.out_hw = (OH, OW)
        self.tb, self.tk = _lanczos.device(device)
        self.kstride = _lanczos.kstride
        self.h_box, self.h_off = boxes, off
        self.box = _dev(boxes, np.int32, device) if upload else None
        self.off = _dev(off, np.int32, device) if upload else None


def crop_lanczos(x: torch.Tensor, boxes, out_hw: Tuple[int, int] = None, want_f32: bool = False, out=None, outf=None,
                 src_index=None):
    """img.crop(box).resize((OW,OH), LANCZOS) per image; boxes[B][4] = left, top, w, h (or a prebuilt CropPlan)."""
    _chk_img(x)
    lib = _ready(x)
    _, H, W, _c = x.shape
    B, sidx, nsrc = _src(x, src_index)
    plan = boxes if isinstance(boxes, CropPlan) else CropPlan(boxes, out_hw, x.device)
    OH, OW = plan.out_hw
    out = _out_like(x, B, out, (B, OH, OW, 3))
    if want_f32 and outf is None:
        outf = torch.empty((B, OH, OW, 3), dtype=torch.float32, device=x.device)
    _lib.check(lib.lfx_crop_lanczos(_p(x), _p(out), _p(outf) if want_f32 else None, B, H, W, _p(plan.box), OH, OW,
                                    _p(plan.tb), _p(plan.tk), plan.kstride, _p(plan.off), _p(sidx), nsrc, _stream()))
    return (out, outf) if want_f32 else out


def distort(x: torch.Tensor, noise_u8: torch.Tensor, cuts: Sequence[int], src_index=None, out=None, hist_ws=None) -> torch.Tensor:
    """(x + noise) mod 256 then per-channel autocontrast; cuts[i] = int(H*W*cutoff_i // 100)."""
    _chk_img(x)
    lib = _ready(x)
    _, H, W, _c = x.shape
    B, sidx, nsrc = _src(x, src_index)
    out = _out_like(x, B, out)
    hist = hist_ws if hist_ws is not None else torch.empty((B, 3, 256), dtype=torch.int32, device=x.device)
    dcut = _dev(cuts, np.int32, x.device)
    _lib.check(lib.lfx_distort(_p(x), _p(noise_u8), _p(out), B, H, W, _p(dcut), _p(hist), _p(sidx), nsrc, _stream()))
    return out


def legacy_normal_noise(seeds: Sequence[int], n: int, scale: float, device, loc: float = 0.0, dseeds: torch.Tensor = None,
                        out: torch.Tensor = None) -> torch.Tensor:
    """uint8 [len(seeds), n]: row i = np.random.normal(loc, scale, n).astype(np.uint8) after np.random.seed(seeds[i]),
    generated on the GPU (NumPy legacy MT19937 + polar gauss, lfx_rng.cu).  Seed 0 means "unseeded" in the reference
    (image_augmenter.py:16: `if seed:`): those rows are drawn from the host's current np.random state."""
    if not torch.cuda.is_available():
        raise RuntimeError("leaffliction_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    lib = _lib.init(device.index if device.index is not None else torch.cuda.current_device())
    seeds = np.asarray(seeds, np.int64).reshape(-1)
    if out is None:
        out = torch.empty((len(seeds), int(n)), dtype=torch.uint8, device=device)
    ds = dseeds if dseeds is not None else _dev((seeds & 0xFFFFFFFF).astype(np.uint32).view(np.int32), np.int32, device)
    _lib.check(lib.lfx_legacy_normal_u8(_p(ds), _p(out), len(seeds), int(n), float(loc), float(scale), _stream()))
    for i in np.nonzero(seeds == 0)[0]:
        out[int(i)] = torch.from_numpy(np.random.normal(loc, scale, int(n)).astype(np.uint8)).to(device)
    return out


# ----------------------------------------------------------------------------- cv2.resize of the mask path
_cubic_tabs = {}


def _cubic_table(in_size: int, out_size: int, device):
    key = (int(in_size), int(out_size), str(device))
    if key not in _cubic_tabs:
        first = np.zeros(out_size, np.int32)
        w = np.zeros((out_size, 4), np.int32)
        _lib.check(_lib.load().lfx_cubic_table(int(in_size), int(out_size), first.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p)))
        _cubic_tabs[key] = (_dev(first, np.int32, device), _dev(w, np.int32, device))
    return _cubic_tabs[key]


def resize_cubic(x: torch.Tensor, out_hw: Tuple[int, int]) -> torch.Tensor:
    """cv2.resize(img, (OW, OH), interpolation=cv2.INTER_CUBIC) per image (8-bit fixed-point path, +-1 LSB class)."""
    _chk_img(x)
    lib = _ready(x)
    B, H, W, _ = x.shape
    OH, OW = int(out_hw[0]), int(out_hw[1])
    xf, xw = _cubic_table(W, OW, x.device)
    yf, yw = _cubic_table(H, OH, x.device)
    out = torch.empty((B, OH, OW, 3), dtype=torch.uint8, device=x.device)
    _lib.check(lib.lfx_resize_cubic(_p(x), _p(out), B, H, W, OH, OW, _p(xf), _p(xw), _p(yf), _p(yw), _stream()))
    return out


def resize_nearest(x: torch.Tensor, out_hw: Tuple[int, int]) -> torch.Tensor:
    """cv2.resize(img, (OW, OH), interpolation=cv2.INTER_NEAREST) per image; x uint8 [B,H,W] or [B,H,W,C]."""
    if x.dtype != torch.uint8 or x.dim() not in (3, 4) or not x.is_contiguous():
        raise ValueError("resize_nearest: contiguous uint8 [B,H,W] or [B,H,W,C] expected")
    lib = _ready(x)
    B, H, W = x.shape[:3]
    Cn = 1 if x.dim() == 3 else int(x.shape[3])
    OH, OW = int(out_hw[0]), int(out_hw[1])
    out = torch.empty((B, OH, OW) if x.dim() == 3 else (B, OH, OW, Cn), dtype=torch.uint8, device=x.device)
    _lib.check(lib.lfx_resize_nearest(_p(x), _p(out), B, H, W, Cn, OH, OW, _stream()))
    return out


# ----------------------------------------------------------------------------- packed asynchronous parameter uploads
class PackedUpload:
    """Several small host arrays -> ONE pinned staging buffer -> one non-blocking H2D copy -> typed device views.
    A pageable `.to(device)` per array blocks the host until the stream has drained; a hot loop that uploads a dozen
    parameter arrays between kernels serialises host and GPU that way.  A small ring of pinned buffers guarded by events
    keeps a buffer from being refilled while its copy is still queued."""
    _ring = {}

    def __init__(self, device, slots: int = 4):
        self.device = device
        key = str(device)
        if key not in PackedUpload._ring:
            PackedUpload._ring[key] = {"bufs": [None] * slots, "events": [None] * slots, "next": 0}
        self.state = PackedUpload._ring[key]

    def upload(self, arrays):
        """arrays: {name: np.ndarray} -> {name: device tensor of the same dtype and shape}."""
        st = self.state
        lay, off = {}, 0
        for k, a in arrays.items():
            a = np.ascontiguousarray(a)
            lay[k] = (off, a)
            off += (a.nbytes + 15) & ~15
        total = max(off, 16)
        i = st["next"]
        st["next"] = (i + 1) % len(st["bufs"])
        if st["bufs"][i] is None or st["bufs"][i].numel() < total:
            # grow EVERY slot at once (pinned allocations cost milliseconds: none may be left for the steady state)
            for ev in st["events"]:
                if ev is not None:
                    ev.synchronize()
            cap = max(total + total // 2, 4 << 20)
            for k in range(len(st["bufs"])):
                st["bufs"][k] = torch.empty(cap, dtype=torch.uint8).pin_memory()
                st["events"][k] = None
        if st["events"][i] is not None:
            st["events"][i].synchronize()
        hb = st["bufs"][i]
        hnp = hb.numpy()
        for k, (o, a) in lay.items():
            hnp[o:o + a.nbytes] = a.reshape(-1).view(np.uint8)
        db = hb[:total].to(self.device, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        st["events"][i] = ev
        out = {}
        for k, (o, a) in lay.items():
            t = db[o:o + a.nbytes].view(_TORCH_DT[a.dtype.type])
            out[k] = t.view(a.shape)
        return out


_TORCH_DT = {np.int32: torch.int32, np.int64: torch.int64, np.float64: torch.float64, np.float32: torch.float32,
             np.uint8: torch.uint8, np.uint32: torch.int32}
