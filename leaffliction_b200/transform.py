"""Drop-in for the reference's transform call surface:

  TransformConfig / load_config           srcs/cli/Transformation.py:63-185
  make_mask(rgb, cfg)                     srcs/transform/filters/mask.py:548-582
  apply_mask(img, mask, mask_color)       srcs/utils/mask_utils.py:10-83
  apply_mask_filter / apply_roi_filter    mask.py:585-607, roi.py:20-46
  TransformPipeline                       Transformation.py:326-390

Same names, argument meaning and error behaviour; arrays in, arrays out (NumPy, uint8 HWC), inputs
never mutated.  The arithmetic runs in libleafx's CUDA kernels through the C ABI -- there is no CPU
fallback and no OpenCV/PlantCV on this path.

Parity contract: the reference's grabCut refinement, k-means strategy and 1.3x
cubic mask upscale are not reproducible even against the reference itself, so this path always
runs parity profile P0 (those three off) and logs once when the config asks for them.
"""
from __future__ import annotations

import logging
import sys
from dataclasses import dataclass
from pathlib import Path
from typing import Dict, Iterable, List, Optional, Tuple

import numpy as np

logger = logging.getLogger(__name__)

IMAGE_EXTS = {".jpg"}
DEFAULT_TYPES = ("Blur", "Mask", "ROI", "Analyze", "Landmarks", "Hist", "Brown")
CANONICAL_TYPES: Dict[str, str] = {
    "blur": "Blur", "mask": "Mask", "roi": "ROI", "analyze": "Analyze", "analyse": "Analyze",
    "landmarks": "Landmarks", "pseudolandmarks": "Landmarks", "pseudo-landmarks": "Landmarks",
    "hist": "Hist", "histogram": "Hist", "brown": "Brown", "disease": "Brown", "spots": "Brown",
}
GPU_STRATEGIES = ("hsv_h", "lab", "hsv_s", "hsv_v_dark", "inclusive", "enhanced", "kmeans", "auto")

_CONFIG_FIELDS = (
    ("gaussian_sigma", float), ("hsv_channel_for_mask", str), ("fill_size", int), ("morph_kernel", int),
    ("landmarks_count", int), ("roi_size", tuple), ("mask_strategy", str), ("bg_bias", None),
    ("grabcut_refine", bool), ("green_hue_range", tuple), ("min_object_area_ratio", float),
    ("max_object_area_ratio", float), ("mask_upscale_factor", float), ("mask_upscale_long_side", int),
    ("shadow_suppression", bool), ("shadow_s_max", int), ("shadow_v_method", str), ("shadow_v_percentile", int),
    ("shadow_morphology_kernel", int), ("brown_hue_range", tuple), ("brown_s_min", int), ("brown_v_max", int),
    ("brown_min_area_px", int), ("brown_morph_kernel", int), ("use_lab_brown", bool), ("lab_b_min", int),
    ("lab_a_min", int), ("debug_shadow_visualization", bool),
)


@dataclass(frozen=True)
class TransformConfig:
    """Same 28 fields as the reference's frozen dataclass (Transformation.py:63-93)."""
    gaussian_sigma: float
    hsv_channel_for_mask: str
    fill_size: int
    morph_kernel: int
    landmarks_count: int
    roi_size: Tuple[int, int]
    mask_strategy: str
    bg_bias: Optional[str]
    grabcut_refine: bool
    green_hue_range: Tuple[int, int]
    min_object_area_ratio: float
    max_object_area_ratio: float
    mask_upscale_factor: float
    mask_upscale_long_side: int
    shadow_suppression: bool
    shadow_s_max: int
    shadow_v_method: str
    shadow_v_percentile: int
    shadow_morphology_kernel: int
    brown_hue_range: Tuple[int, int]
    brown_s_min: int
    brown_v_max: int
    brown_min_area_px: int
    brown_morph_kernel: int
    use_lab_brown: bool
    lab_b_min: int
    lab_a_min: int
    debug_shadow_visualization: bool


DEFAULT_CONFIG_VALUES = dict(
    gaussian_sigma=1.5, hsv_channel_for_mask="s", fill_size=1000, morph_kernel=3, landmarks_count=80,
    roi_size=(256, 256), mask_strategy="inclusive", bg_bias="light_bg", grabcut_refine=True,
    green_hue_range=(25, 100), min_object_area_ratio=0.10, max_object_area_ratio=0.98, mask_upscale_factor=1.3,
    mask_upscale_long_side=1500, shadow_suppression=False, shadow_s_max=40, shadow_v_method="percentile",
    shadow_v_percentile=5, shadow_morphology_kernel=3, brown_hue_range=(0, 30), brown_s_min=20, brown_v_max=200,
    brown_min_area_px=25, brown_morph_kernel=3, use_lab_brown=False, lab_b_min=125, lab_a_min=125,
    debug_shadow_visualization=False)


def default_config(**over) -> TransformConfig:
    """The values of the reference's srcs/transform/config.yaml, with overrides."""
    v = dict(DEFAULT_CONFIG_VALUES)
    v.update(over)
    return TransformConfig(**v)


def load_config(path: Optional[Path]) -> TransformConfig:
    """YAML -> TransformConfig; all 28 keys required, extra keys ignored; on any problem logs and
    exits with status 1 (Transformation.py:105-185)."""
    if not path:
        logging.error("No configuration file path provided")
        sys.exit(1)
    path = Path(path)
    if not path.exists():
        logging.error("Configuration file not found: %s", path)
        sys.exit(1)
    try:
        import yaml
        with path.open("r", encoding="utf-8") as fh:
            data = yaml.safe_load(fh) or {}
        missing = [name for name, _ in _CONFIG_FIELDS if name not in data]
        if missing:
            logging.error("Missing required configuration fields: %s", missing)
            sys.exit(1)
        vals = {}
        for name, conv in _CONFIG_FIELDS:
            vals[name] = data[name] if conv is None else conv(data[name])
        return TransformConfig(**vals)
    except SystemExit:
        raise
    except Exception as exc:
        logging.error("Failed to read configuration file (%s)", exc)
        sys.exit(1)


# --------------------------------------------------------------------------- small host helpers
def is_image(path: Path) -> bool:
    return path.is_file() and path.suffix.lower() in IMAGE_EXTS


def iter_images_in_dir(src: Path) -> Iterable[Path]:
    for p in sorted(src.rglob("*")):
        if is_image(p):
            yield p


def build_types_filter(arg: Optional[str]) -> Tuple[str, ...]:
    if not arg:
        return DEFAULT_TYPES
    out: List[str] = []
    for tok in (t.strip() for t in str(arg).split(",")):
        if not tok:
            continue
        canon = CANONICAL_TYPES.get(tok.lower())
        if canon is None:
            logging.warning("Unknown transform type skipped: %s", tok)
        elif canon not in out:
            out.append(canon)
    return tuple(out) if out else DEFAULT_TYPES


def output_names(stem: str) -> Dict[str, str]:
    return {t: f"{stem}__T_{t}.jpg" for t in DEFAULT_TYPES}


def pil_read_rgb(path: Path) -> np.ndarray:
    from PIL import Image, ImageOps
    with Image.open(path) as im:
        im = ImageOps.exif_transpose(im)
        return np.array(im.convert("RGB"))


def imwrite_rgb(path: Path, rgb_img) -> None:
    """JPEG encode on the host (codec is out of scope).  The reference writes with
    cv2.imwrite (Transformation.py:196-205); Pillow quality=95 is used here (no OpenCV on this path)."""
    if rgb_img is None:
        return
    from PIL import Image
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    Image.fromarray(np.asarray(rgb_img)).save(path, quality=95)


# --------------------------------------------------------------------------- GPU plumbing
_warned = set()


def _warn_once(key, msg):
    if key not in _warned:
        _warned.add(key)
        logger.warning(msg)


def _ops():
    from . import ops
    return ops


def _dev(a: np.ndarray):
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("leaffliction_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _check_rgb(rgb):
    if not isinstance(rgb, np.ndarray) or rgb.ndim != 3 or rgb.shape[2] != 3 or rgb.dtype != np.uint8:
        raise TypeError("expected an RGB uint8 array of shape [H,W,3]")


def mask_cfg_from(cfg: TransformConfig, strategy: Optional[str] = None):
    """lfx_mask_cfg for a TransformConfig (numeric fields only)."""
    st = strategy or cfg.mask_strategy
    dev_strategy = st if st in ("hsv_h", "lab", "hsv_s", "hsv_v_dark") else "external"
    return _ops().mask_cfg(
        strategy=dev_strategy, green_hue_range=cfg.green_hue_range, fill_size=cfg.fill_size, morph_kernel=cfg.morph_kernel,
        brown_hue_range=cfg.brown_hue_range, brown_s_min=cfg.brown_s_min, brown_v_max=cfg.brown_v_max,
        brown_min_area_px=cfg.brown_min_area_px, brown_morph_kernel=cfg.brown_morph_kernel,
        use_lab_brown=cfg.use_lab_brown, lab_a_min=cfg.lab_a_min, lab_b_min=cfg.lab_b_min,
        hsv_channel_for_mask=cfg.hsv_channel_for_mask, bg_bias=cfg.bg_bias, extend_brown=True)


def _profile_checks(cfg: TransformConfig):
    if cfg.grabcut_refine:
        _warn_once("grabcut", "grabcut_refine=true is not reproducible in the reference itself (global OpenCV RNG); "
                              "running parity profile P0 without it")
    if ((cfg.mask_upscale_factor and cfg.mask_upscale_factor > 1.0) or (cfg.mask_upscale_long_side and cfg.mask_upscale_long_side > 0)) \
            and not upscale_enabled():
        _warn_once("upscale", "mask upscaling (INTER_CUBIC) is ISA-dependent in OpenCV itself; running parity profile P0 at native "
                              "size (LEAFX_MASK_UPSCALE=1 or transform.set_upscale(True) enables the +-1 LSB upscale path)")
    if cfg.shadow_suppression:
        _warn_once("shadow", "shadow_suppression uses k-means on OpenCV's global RNG (tier C); ignored")
    if cfg.mask_strategy not in GPU_STRATEGIES:
        raise ValueError(f"mask_strategy {cfg.mask_strategy!r} is outside the bit-exact contract "
                         f"(supported: {', '.join(GPU_STRATEGIES)})")


def make_mask_batch(rgb_batch: np.ndarray, cfg: TransformConfig):
    """Batched make_mask: uint8 [B,H,W,3] -> (masks [B,H,W], info [B,8], contours list)."""
    _profile_checks(cfg)
    ops = _ops()
    x = _dev(rgb_batch)
    H, W = rgb_batch.shape[1:3]
    scale = working_scale(H, W, cfg)
    x0 = x
    if abs(scale - 1.0) >= 1e-6:          # _prepare_working_image (mask.py:29-50): INTER_CUBIC upscale before masking
        x = ops.resize_cubic(x, (int(round(H * scale)), int(round(W * scale))))
    raw = None
    if cfg.mask_strategy == "auto":
        raw, _choice, _scores = auto_candidate(x, cfg)
    elif cfg.mask_strategy == "kmeans":
        raw = kmeans_candidate(x, cfg)
    elif cfg.mask_strategy in ("inclusive", "enhanced"):
        try:
            raw = ops.raw_mask_front_end(x, cfg.mask_strategy, mask_cfg_from(cfg))
        except Exception as e:
            from ._lib import ERR_UNSUPPORTED
            if getattr(e, "code", None) != ERR_UNSUPPORTED or x is x0:
                raise
            # the front ends keep the whole image in shared memory (H*W <= ~65536): mask at native size instead
            _warn_once("upscale-front", f"{cfg.mask_strategy} front end cannot take the upscaled {tuple(x.shape[1:3])} working "
                                        "image; masking at native size")
            x, scale = x0, 1.0
            raw = ops.raw_mask_front_end(x, cfg.mask_strategy, mask_cfg_from(cfg))
    mask, info = ops.make_mask(x, mask_cfg_from(cfg), raw)
    max_pts = 4096
    while True:
        pts, cnt, _ = ops.trace_contour(mask, info, max_pts)
        cnt_h = cnt.cpu().numpy()
        if (cnt_h >= 0).all():
            break
        max_pts = int(-cnt_h.min()) + 16
    pts_h = pts.cpu().numpy()
    contours = [pts_h[i, : cnt_h[i]].reshape(-1, 1, 2).copy() if cnt_h[i] > 0 else None for i in range(len(cnt_h))]
    info_h = info.cpu().numpy()
    if abs(scale - 1.0) >= 1e-6:          # _resize_results_to_original (mask.py:526-545)
        mask = ops.resize_nearest(mask, (H, W))
        contours = [None if c is None else (c.astype(np.float32) / np.float32(scale)).astype(np.int32) for c in contours]
        for i, c in enumerate(contours):
            if c is not None:
                info_h[i, 1:5] = bounding_rect(c)
    return mask.cpu().numpy(), info_h, contours


AUTO_CANDIDATES = ("hsv_s", "hsv_v_dark", "hsv_h", "lab", "kmeans", "enhanced", "inclusive")     # mask.py:435-441


def kmeans_candidate(x, cfg: TransformConfig):
    """`_create_kmeans_mask` (mask.py:109-140) for a device batch of any size: INTER_AREA working copy with a 256-pixel
    longer side (mask.py:113-118; the letterbox kernel with the whole image as its box: same cv2.resize arithmetic, same
    max(int(w * scale), 1) sizes), lfx_kmeans_raw on it, INTER_NEAREST back (mask.py:139)."""
    import torch
    ops = _ops()
    B, H, W = int(x.shape[0]), int(x.shape[1]), int(x.shape[2])
    if max(H, W) == 256:
        return ops.kmeans_raw(x, cfg.green_hue_range, cfg.bg_bias)
    scale = 256 / max(H, W)
    sw, sh = max(1, int(W * scale)), max(1, int(H * scale))
    info = torch.tensor([[1, 0, 0, W, H, 0, 0, 0]] * B, dtype=torch.int32, device=x.device)
    canvas = ops.roi_letterbox(x, None, info, (256, 256))
    oy, ox = (256 - sh) // 2, (256 - sw) // 2
    small = canvas[:, oy:oy + sh, ox:ox + sw, :].contiguous()
    if max(sh, sw) != 256:                       # cannot happen for scale = 256 / max(H, W); guards the kernel's contract
        raise ValueError("kmeans_candidate: working copy without a 256-pixel side")
    return ops.resize_nearest(ops.kmeans_raw(small, cfg.green_hue_range, cfg.bg_bias), (H, W))


def score_mask_terms(area2: int, hull_area: float, bbox, h: int, w: int, b_strength: float, green_frac: float, cfg) -> float:
    """_score_mask (mask.py:143-188) from its terms: contour area (2 * area, exact), convex-hull area, bounding box,
    boundary strength and green fraction.  cnt None is area2 < 0."""
    if area2 < 0:
        return -1.0
    area = area2 / 2.0
    if area <= 1:
        return -1.0
    area_ratio = area / float(h * w)
    if area_ratio < cfg.min_object_area_ratio or area_ratio > cfg.max_object_area_ratio:
        return 0.01
    solidity = (area / hull_area) if hull_area > 1 else 0.0
    x, y, ww, hh = bbox
    touches = (x <= 0) or (y <= 0) or (x + ww >= w - 1) or (y + hh >= h - 1)
    target = 0.35
    area_term = max(0.0, 1.0 - abs(area_ratio - target) / target)
    score = 0.35 * area_term + 0.25 * solidity + 0.25 * b_strength + 0.15 * green_frac
    if touches:
        score *= 0.75
    return float(score)


def auto_candidate(x, cfg: TransformConfig):
    """`mask_strategy: auto` (mask.py:435-461) on a device batch x [B,H,W,3]: the seven candidates in the reference's
    order (hsv_s, hsv_v_dark, hsv_h, lab, kmeans, enhanced, inclusive), each through
    _postprocess_mask, scored by _score_mask, the first strictly greater score wins.  Returns (raw candidate of the winner per image [B,H,W] -- all zero when every candidate is rejected, which
    sends make_mask down the reference's Otsu fallback --, chosen index [B] (-1 = none), scores [K,B]).
    The scores' float terms are accumulated in fp64 on the device (the reference: float32 NumPy mean); two candidates
    whose scores differ by less than ~1e-6 may therefore rank differently."""
    import copy

    import torch
    ops = _ops()
    B, H, W = int(x.shape[0]), int(x.shape[1]), int(x.shape[2])
    raws = []
    for st in AUTO_CANDIDATES:
        if st == "kmeans":
            raws.append(kmeans_candidate(x, cfg))
        elif st in ("enhanced", "inclusive"):
            raws.append(ops.raw_mask_front_end(x, st, mask_cfg_from(cfg, "hsv_h")))
        else:
            raws.append(ops.strategy_raw(x, mask_cfg_from(cfg, st)))
    K = len(raws)
    masks = torch.empty((K, B, H, W), dtype=torch.uint8, device=x.device)
    infos, recs = [], []
    for k in range(K):
        m, info = ops.postprocess_mask(raws[k], cfg.fill_size, cfg.morph_kernel)
        masks[k].copy_(m)
        infos.append(info)
        recs.append(ops.analyze_records(m, info, max_pts=8192))
    feat, gmax, gmin = ops.score_features(x, masks, cfg.green_hue_range)
    feat_h = feat.cpu().numpy()
    gmax_h, gmin_h = gmax.cpu().numpy().astype(np.float64), gmin.cpu().numpy().astype(np.float64)
    scores = np.full((K, B), -1.0)
    for k in range(K):
        info_h = infos[k].cpu().numpy()
        rf = recs[k]["rec_f"].cpu().numpy()
        ri = recs[k]["rec_i"].cpu().numpy()
        if (ri[:, 1] < 0).any() or (ri[:, 12] < 0).any():
            raise RuntimeError("auto strategy: contour or hull buffer too small")
        for i in range(B):
            if not info_h[i, 0]:
                continue
            bsum, bcnt, mpx, gpx = feat_h[k, i]
            rng_ = gmax_h[i] - gmin_h[i]
            scale = (1.0 / rng_) if rng_ > 2.220446049250313e-16 else 0.0      # cv2.normalize NORM_MINMAX to [0, 1]
            b_strength = ((bsum / bcnt) * scale - gmin_h[i] * scale) if bcnt > 0 else 0.0
            green_frac = gpx / max(1.0, mpx)
            scores[k, i] = score_mask_terms(int(info_h[i, 5]), float(rf[i, 1]), tuple(int(v) for v in info_h[i, 1:5]), H, W,
                                            b_strength, green_frac, cfg)
    choice = np.full(B, -1, np.int64)
    best = np.full(B, -1.0)
    for k in range(K):                      # _find_best_mask: strictly greater, candidates in order
        better = scores[k] > best
        choice[better] = k
        best[better] = scores[k][better]
    raw = torch.zeros((B, H, W), dtype=torch.uint8, device=x.device)
    for k in range(K):
        idx = np.nonzero(choice == k)[0]
        if len(idx):
            ti = torch.from_numpy(idx).to(x.device)
            raw[ti] = raws[k][ti]
    return raw, choice, scores


_UPSCALE = None


def set_upscale(on: Optional[bool]) -> None:
    """Honour mask_upscale_factor / mask_upscale_long_side (True), ignore them = parity profile P0 (False), or follow the
    environment variable LEAFX_MASK_UPSCALE (None, the default)."""
    global _UPSCALE
    _UPSCALE = on


def upscale_enabled() -> bool:
    import os
    return bool(_UPSCALE) if _UPSCALE is not None else os.environ.get("LEAFX_MASK_UPSCALE", "0") not in ("", "0", "false", "False")


def working_scale(h: int, w: int, cfg: TransformConfig) -> float:
    """Scale of the working image of make_mask (_prepare_working_image, mask.py:33-39); 1.0 unless the upscale path is
    enabled (its cubic resize is only +-1 LSB from cv2's, so it is outside the bit-exact profile P0)."""
    s = 1.0
    if not upscale_enabled():
        return s
    if cfg.mask_upscale_factor and cfg.mask_upscale_factor > 1.0:
        s = float(cfg.mask_upscale_factor)
    elif cfg.mask_upscale_long_side and cfg.mask_upscale_long_side > 0:
        ls = max(h, w)
        if ls < cfg.mask_upscale_long_side:
            s = float(cfg.mask_upscale_long_side) / float(ls)
    return s


def make_mask(rgb: np.ndarray, cfg: TransformConfig) -> Tuple[Optional[np.ndarray], Optional[np.ndarray]]:
    """(mask uint8 [H,W] 0/255, contour int32 [K,1,2] | None) -- mask.py:548-582."""
    _check_rgb(rgb)
    masks, _, contours = make_mask_batch(rgb[None], cfg)
    return masks[0], contours[0]


def apply_mask(img: np.ndarray, mask: np.ndarray, mask_color: str = "white") -> np.ndarray:
    """PlantCV-style apply_mask (mask_utils.py:10-83): same validation and exceptions."""
    if mask_color.upper() == "WHITE":
        color_val = 255
    elif mask_color.upper() == "BLACK":
        color_val = 0
    else:
        raise ValueError(f'Mask Color {mask_color} is not "white" or "black"!')
    if not isinstance(img, np.ndarray):
        raise TypeError("img must be a numpy array")
    if not isinstance(mask, np.ndarray):
        raise TypeError("mask must be a numpy array")
    if mask.ndim == 3:
        if mask.shape[2] == 3:
            # cv2.cvtColor(mask, COLOR_BGR2GRAY): the same fixed-point weights with B first
            g = _ops().cvt_color(_dev(np.ascontiguousarray(mask[None, ..., ::-1])), "gray").cpu().numpy()[0]
            mask = g
        else:
            mask = mask[:, :, 0]
    elif mask.ndim != 2:
        raise ValueError("mask must be 2D or 3D array")
    if img.ndim not in (2, 3):
        raise ValueError("img must be 2D (grayscale) or 3D (color) array")
    if img.ndim == 2:
        x = np.repeat(img[..., None], 3, axis=2)
        return _ops().apply_mask(_dev(x[None]), _dev(mask[None].astype(np.uint8)), color_val).cpu().numpy()[0, ..., 0].copy()
    return _ops().apply_mask(_dev(img[None]), _dev(mask[None].astype(np.uint8)), color_val).cpu().numpy()[0]


def apply_mask_filter(rgb: np.ndarray, cfg: TransformConfig, make_mask_func) -> np.ndarray:
    mask_img, _ = make_mask_func(rgb)
    if mask_img is not None:
        return apply_mask(rgb, mask_img, mask_color="black")
    return rgb


def bounding_rect(contour: np.ndarray) -> Tuple[int, int, int, int]:
    """cv2.boundingRect of an integer point set (roi.py:26)."""
    p = contour.reshape(-1, 2)
    x0, y0 = int(p[:, 0].min()), int(p[:, 1].min())
    return x0, y0, int(p[:, 0].max()) - x0 + 1, int(p[:, 1].max()) - y0 + 1


def draw_rectangle(img: np.ndarray, x, y, w, h, color=(255, 0, 0)) -> np.ndarray:
    """cv2.rectangle(vis, (x, y), (x + w, y + h), color, 2) on a copy (roi.py:43-44), drawn by lfx_draw_rectangles
    (OpenCV's 2-px PolyLine restated, bit-identical, clipped at the image border like OpenCV)."""
    info = np.array([[1, x, y, w, h, 0, 0, 0]], np.int32)
    return _ops().draw_rectangles(_dev(np.ascontiguousarray(img)[None]), _dev(info), color, 2).cpu().numpy()[0]


def apply_roi_filter(rgb: np.ndarray, contour: Optional[np.ndarray], cfg: TransformConfig):
    """(canvas, vis | None, (x,y,w,h) | None) -- roi.py:20-46."""
    if contour is None:
        return rgb, None, None
    x, y, w, h = bounding_rect(contour)
    if w <= 0 or h <= 0:
        return rgb, None, None
    H, W = cfg.roi_size
    info = np.array([[1, x, y, w, h, 0, 0, 0]], np.int32)
    canvas = _ops().roi_letterbox(_dev(rgb[None]), None, _dev(info), (H, W)).cpu().numpy()[0]
    return canvas, draw_rectangle(rgb, x, y, w, h), (x, y, w, h)


class TransformPipeline:
    """Facade with the reference's method names (Transformation.py:326-390)."""

    def __init__(self, cfg: TransformConfig) -> None:
        self.cfg = cfg

    def make_mask(self, rgb):
        return make_mask(rgb, self.cfg)

    def create_masked_rgb(self, rgb, mask):
        return rgb if mask is None else apply_mask(rgb, mask, mask_color="white")

    def apply_mask(self, rgb, mask, mask_color="white"):
        return apply_mask(rgb, mask, mask_color)

    def roi(self, rgb, contour):
        return apply_roi_filter(rgb, contour, self.cfg)

    def blur(self, rgb):
        from .filters import apply_blur_filter
        return apply_blur_filter(rgb, self.cfg, self.make_mask)

    def analyze(self, rgb, mask, contour):
        from .filters import apply_analyze_filter
        return apply_analyze_filter(rgb, mask, contour, self.cfg)

    def detect_brown_spots(self, rgb, mask):
        from .filters import apply_brown_filter
        return apply_brown_filter(rgb, mask, self.cfg)

    def histogram_hsv(self, rgb):
        from .filters import apply_histogram_filter
        return apply_histogram_filter(rgb, self.cfg)

    def pseudolandmarks(self, rgb, contour):
        raise NotImplementedError("Landmarks are out of scope for the hot path (SURVEY.md section 8f #4)")
