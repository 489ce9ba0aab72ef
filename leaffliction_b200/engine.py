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

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in
                   (self.blur, self.mask, self.info, self.roi, self.hist9, self.hsv3, self.counters))


def alloc_host_outputs(B, H, W, roi_size=(256, 256)) -> HostOutputs:
    def pin(shape, dt):
        return torch.empty(shape, dtype=dt).pin_memory()
    return HostOutputs(pin((B, H, W, 3), torch.uint8), pin((B, H, W), torch.uint8), pin((B, 8), torch.int32),
                       pin((B, roi_size[0], roi_size[1], 3), torch.uint8), pin((B, 9, 256), torch.int32),
                       pin((B, 3, 256), torch.int32), pin((B, 16), torch.int32))


class TransformEngine:
    def __init__(self, H: int, W: int, cfg=None, gaussian_sigma: float = 1.5, roi_size=(256, 256),
                 device: Optional[torch.device] = None, chunk: int = 512, front: Optional[str] = None):
        """`front`: None = the strategy in `cfg` (hsv_h / lab / hsv_s / hsv_v_dark, fused kernel where the shape allows);
        'inclusive' (the reference's default strategy) or 'enhanced' = raw candidate by the front-end kernel first."""
        if not torch.cuda.is_available():
            raise RuntimeError("TransformEngine needs a CUDA device: leaffliction_b200 has no CPU fallback")
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.H, self.W = H, W
        self.cfg = cfg if cfg is not None else ops.mask_cfg("hsv_h")
        self.sigma = float(gaussian_sigma)
        self.roi_size = (int(roi_size[0]), int(roi_size[1]))
        self.chunk = int(chunk)
        if front not in (None, "inclusive", "enhanced"):
            raise ValueError(f"front must be None, 'inclusive' or 'enhanced', not {front!r}")
        self.front = front
        self._bufs = None
        self._streams = None

    # ---- device-resident
    def _pipeline(self, x, out):
        if self.front:
            return ops.pipeline_front(x, self.front, self.cfg, self.sigma, self.roi_size, out)
        return ops.pipeline_core(x, self.cfg, self.sigma, self.roi_size, out)

    def run_device(self, x: torch.Tensor, out: Optional[ops.CoreOutputs] = None) -> ops.CoreOutputs:
        return self._pipeline(x, out)

    # ---- host buffers in, host buffers out
    def _ensure(self):
        if self._bufs is None:
            dev = self.device
            self._bufs = [(torch.empty((self.chunk, self.H, self.W, 3), dtype=torch.uint8, device=dev),
                           ops.alloc_core_outputs(self.chunk, self.H, self.W, self.roi_size, dev)) for _ in range(2)]
            self._streams = (torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev))

    def run_host(self, images: torch.Tensor, out: Optional[HostOutputs] = None) -> HostOutputs:
        """images: host uint8 [B,H,W,3] (torch tensor, ideally pinned; numpy is wrapped)."""
        if isinstance(images, np.ndarray):
            images = torch.from_numpy(images)
        B = images.shape[0]
        if out is None:
            out = alloc_host_outputs(B, self.H, self.W, self.roi_size)
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
                ev_out[i].record(s_out)
        for s in self._streams:
            cur.wait_stream(s)
        return out
