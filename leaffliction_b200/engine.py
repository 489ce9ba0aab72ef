"""Host-facing batch engine for the core transform profile.

`TransformEngine.run_host` is the call a user of the reference's folder mode
(Transformation.py:691-696, one task per image on a process pool) makes instead: one batched
submission per GPU.  Inputs and outputs are HOST arrays (pinned); the engine streams them
through the device in chunks with three CUDA streams (H2D, compute, D2H) so copies overlap the
kernels.  `run_device` is the same computation on images already resident in HBM.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import ops


@dataclass
class HostOutputs:
    blur: torch.Tensor      # [B,H,W,3] u8   (pinned host)
    mask: torch.Tensor      # [B,H,W]   u8
    info: torch.Tensor      # [B,8]     i32
    roi: torch.Tensor       # [B,RH,RW,3] u8
    hist9: torch.Tensor     # [B,9,256] i32
    hsv3: torch.Tensor      # [B,3,256] i32
    counters: torch.Tensor  # [B,16]    i32
    ready: Optional[torch.cuda.Event] = None   # recorded after the last device-to-host copy of run_host
    # the six augment outputs (TransformEngine(augment=True)): flip / skew / shear / crop / distortion [B,H,W,3] u8,
    # rotate = pitched slab [B, stride] with image i = rotate[i, :nh*nw*3].reshape(nh, nw, 3), (nh, nw) = rotate_hw[i]
    aug: Optional[dict] = None
    rotate_hw: Optional[np.ndarray] = None

    def nbytes(self) -> int:
        n = sum(t.numel() * t.element_size() for t in
                (self.blur, self.mask, self.info, self.roi, self.hist9, self.hsv3, self.counters))
        if self.aug:
            n += sum(t.numel() for t in self.aug.values())   # the rotate slab is copied whole (pitched rows)
        return n

    @property
    def rotate_bytes(self) -> int:
        """bytes of the rotate outputs actually copied back (each image's own nh*nw*3)."""
        return int((self.rotate_hw[:, 0].astype(np.int64) * self.rotate_hw[:, 1]).sum()) * 3


def alloc_host_outputs(B, H, W, roi_size=(256, 256), augment: bool = False) -> HostOutputs:
    def pin(shape, dt):
        return torch.empty(shape, dtype=dt).pin_memory()
    out = HostOutputs(pin((B, H, W, 3), torch.uint8), pin((B, H, W), torch.uint8), pin((B, 8), torch.int32),
                      pin((B, roi_size[0], roi_size[1], 3), torch.uint8), pin((B, 9, 256), torch.int32),
                      pin((B, 3, 256), torch.int32), pin((B, 16), torch.int32))
    if augment:
        from .augment import AugmentSet, rotate_matrix
        _, nw30, nh30 = rotate_matrix(30.0, W, H)
        stride = (((nw30 + 1) * (nh30 + 1) * 3 + 15) // 16) * 16
        out.aug = {k: pin((B, H, W, 3), torch.uint8) for k in AugmentSet.OPS if k != "rotate"}
        out.aug["rotate"] = pin((B, stride), torch.uint8)
        out.rotate_hw = np.zeros((B, 2), np.int32)
    return out


class TransformEngine:
    def __init__(self, H: int, W: int, cfg=None, gaussian_sigma: float = 1.5, roi_size=(256, 256),
                 device: Optional[torch.device] = None, chunk: int = 512, front: Optional[str] = None,
                 augment: bool = False, bg_bias: str = "light_bg"):
        """`front`: None = the strategy in `cfg` (hsv_h / lab / hsv_s / hsv_v_dark, fused kernel where the shape allows);
        'inclusive' (the reference's default strategy) or 'enhanced' = raw candidate by the front-end kernel first;
        'kmeans' = the k-means candidate (cv2.kmeans restated, `bg_bias` as TransformConfig.bg_bias; longer side 256).
        `augment`: run_host also produces the six ImageAugmenter outputs of every image (needs `seeds` per call)."""
        if not torch.cuda.is_available():
            raise RuntimeError("TransformEngine needs a CUDA device: leaffliction_b200 has no CPU fallback")
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.H, self.W = H, W
        self.cfg = cfg if cfg is not None else ops.mask_cfg("hsv_h")
        self.sigma = float(gaussian_sigma)
        self.roi_size = (int(roi_size[0]), int(roi_size[1]))
        self.chunk = int(chunk)
        if front not in (None, "inclusive", "enhanced", "kmeans"):
            raise ValueError(f"front must be None, 'inclusive', 'enhanced' or 'kmeans', not {front!r}")
        self.front = front
        self.bg_bias = bg_bias
        self.augment = bool(augment)
        self._aug = None
        self._bufs = None
        self._streams = None

    # ---- device-resident
    def _pipeline(self, x, out, dataset_hist=None):
        if self.front:
            return ops.pipeline_front(x, self.front, self.cfg, self.sigma, self.roi_size, out, dataset_hist, self.bg_bias)
        return ops.pipeline_core(x, self.cfg, self.sigma, self.roi_size, out, dataset_hist)

    def run_device(self, x: torch.Tensor, out: Optional[ops.CoreOutputs] = None,
                   dataset_hist: Optional[torch.Tensor] = None) -> ops.CoreOutputs:
        """`dataset_hist` (int64 [9,256], device): accumulates the batch's colour histograms (the per-rank partial of
        the dataset-level histogram, merged across ranks by one allreduce -- SURVEY.md 8e)."""
        return self._pipeline(x, out, dataset_hist)

    def analyze_device(self, out: ops.CoreOutputs, max_pts: int = 4096, max_hull: int = 512):
        """Batched numeric records of apply_analyze_filter (analyze.py:43-98) for the masks of a run_device result:
        contour points, centroid, extreme points, convex hull, PCA axes -- device tensors, layout in include/leafx.h."""
        return ops.analyze_records(out.mask, out.info, max_pts, max_hull)

    def overlays_device(self, x: torch.Tensor, out: ops.CoreOutputs, masked: Optional[torch.Tensor] = None,
                        max_pts: int = 4096, max_hull: int = 512):
        """The two overlay images of the reference's folder run for a whole batch, nothing leaves the device:
        (analyze overlay, ROI rectangle image) = apply_analyze_filter's image (analyze.py:37-122) and apply_roi_filter's
        `vis` (roi.py:43-44), both drawn over `masked`, the white-background masked image the reference's folder run hands
        to its filters (default: apply_mask(x, mask, white)).  Launches on top of run_device: apply_mask, contour trace,
        record, grey, Canny, overlay, rectangles."""
        if masked is None:
            masked = ops.apply_mask(x, out.mask, 255)
        rec = ops.analyze_records(out.mask, out.info, max_pts, max_hull)
        edges = ops.canny(ops.cvt_color(masked, "gray"), 80, 160, True)
        return ops.analyze_overlay(masked, rec, edges, out.mask), ops.draw_rectangles(masked, out.info)

    # ---- host buffers in, host buffers out
    def _ensure(self):
        if self._bufs is None:
            dev = self.device
            self._bufs = [(torch.empty((self.chunk, self.H, self.W, 3), dtype=torch.uint8, device=dev),
                           ops.alloc_core_outputs(self.chunk, self.H, self.W, self.roi_size, dev)) for _ in range(2)]
            self._streams = (torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev))
            if self.augment:
                from .augment import AugmentSet
                self._aug = [AugmentSet(self.chunk, self.H, self.W, dev) for _ in range(2)]

    def run_host(self, images: torch.Tensor, out: Optional[HostOutputs] = None, sync: bool = True,
                 seeds: Optional[np.ndarray] = None) -> HostOutputs:
        """images: host uint8 [B,H,W,3] (torch tensor, ideally pinned; numpy is wrapped).  Returns when the host
        buffers are filled (`sync=True`); with `sync=False` the copies may still be in flight and `out.ready`
        (a CUDA event) must be synchronised before the buffers are read."""
        if isinstance(images, np.ndarray):
            images = torch.from_numpy(images)
        B = images.shape[0]
        if out is None:
            out = alloc_host_outputs(B, self.H, self.W, self.roi_size, self.augment)
        if self.augment:
            if seeds is None or np.asarray(seeds).shape != (6, B):
                raise ValueError("TransformEngine(augment=True).run_host needs seeds of shape [6, B] (one task seed per op and image)")
            if out.aug is None:
                raise ValueError("run_host: `out` was allocated without augment buffers")
            seeds = np.asarray(seeds, np.int64)
        self._ensure()
        s_in, s_k, s_out = self._streams
        cur = torch.cuda.current_stream(self.device)
        for s in self._streams:
            s.wait_stream(cur)
        n_chunks = (B + self.chunk - 1) // self.chunk
        ev_in = [torch.cuda.Event() for _ in range(n_chunks)]
        ev_k = [torch.cuda.Event() for _ in range(n_chunks)]
        ev_out = [torch.cuda.Event() for _ in range(n_chunks)]
        for i in range(n_chunks):
            a, b = i * self.chunk, min(B, (i + 1) * self.chunk)
            n = b - a
            xin, dev_out = self._bufs[i % 2]
            with torch.cuda.stream(s_in):
                if i >= 2:
                    s_in.wait_event(ev_k[i - 2])       # input buffer free once its kernels finished
                xin[:n].copy_(images[a:b], non_blocking=True)
                ev_in[i].record(s_in)
            with torch.cuda.stream(s_k):
                s_k.wait_event(ev_in[i])
                if i >= 2:
                    s_k.wait_event(ev_out[i - 2])      # output buffers free once copied back
                view = ops.CoreOutputs(dev_out.blur[:n], dev_out.mask[:n], dev_out.info[:n], dev_out.roi[:n],
                                       dev_out.hist9[:n], dev_out.hsv3[:n], dev_out.counters[:n])
                self._pipeline(xin[:n], view)
                if self.augment:
                    aug = self._aug[i % 2]
                    if n == self.chunk:
                        aug.run(xin, seeds[:, a:b])
                    else:   # ragged tail: pad the seeds (the padded rows are computed on stale images and never copied back)
                        sd = np.ones((6, self.chunk), np.int64)
                        sd[:, :n] = seeds[:, a:b]
                        aug.run(xin, sd)
                    out.rotate_hw[a:b] = aug.rotate_hw[:n]
                ev_k[i].record(s_k)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_k[i])
                out.blur[a:b].copy_(dev_out.blur[:n], non_blocking=True)
                out.mask[a:b].copy_(dev_out.mask[:n], non_blocking=True)
                out.info[a:b].copy_(dev_out.info[:n], non_blocking=True)
                out.roi[a:b].copy_(dev_out.roi[:n], non_blocking=True)
                out.hist9[a:b].copy_(dev_out.hist9[:n], non_blocking=True)
                out.hsv3[a:b].copy_(dev_out.hsv3[:n], non_blocking=True)
                out.counters[a:b].copy_(dev_out.counters[:n], non_blocking=True)
                if self.augment:
                    aug = self._aug[i % 2]
                    for k in ("flip", "skew", "shear", "crop", "distortion"):
                        out.aug[k][a:b].copy_(getattr(aug, k)[:n], non_blocking=True)
                    out.aug["rotate"][a:b].copy_(aug.rotate[:n], non_blocking=True)
                ev_out[i].record(s_out)
        for s in self._streams:
            cur.wait_stream(s)
        if sync and n_chunks:
            # the results are HOST buffers: the caller may read them as soon as this returns, so wait (on the host)
            # for the last device-to-host copy.  sync=False returns at once with `out.ready` = the event to wait on.
            ev_out[-1].synchronize()
        out.ready = ev_out[-1] if n_chunks else None
        return out
