"""INTEGRATION.md section B is the binding a reference maintainer would paste: this test extracts its two code blocks, runs them
against libleafx.so as they stand, and compares the results with the package's own wrappers (which the parity tests cover)."""
import os
import re
import types

import numpy as np
import pytest
import torch

from leaffliction_b200 import _lib, engine, ops, synth, transform

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _stub_module():
    txt = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    sec = txt[txt.index("## B. Bind the C ABI directly"):txt.index("## C. Entry points")]
    blocks = re.findall(r"```python\n(.*?)```", sec, flags=re.S)
    assert len(blocks) >= 2
    src = "\n".join(blocks[:2]).replace('C.CDLL("libleafx.so")', f'C.CDLL("{_lib.LIB_PATH}")')
    mod = types.ModuleType("leafx_ffi_stub")
    exec(compile(src, "INTEGRATION.md#B", "exec"), mod.__dict__)
    return mod


def test_integration_stub_runs_and_matches(dev):
    _lib.load()
    ffi = _stub_module()
    imgs = synth.leaf_batch(6, 256, 256, 61)
    cfg = transform.default_config(mask_strategy="hsv_h", grabcut_refine=False, mask_upscale_factor=1.0, mask_upscale_long_side=0)
    blur, mask, boxes, roi, hist = ffi.transform_batch(imgs, cfg)
    x = torch.from_numpy(imgs).to(dev)
    out = engine.TransformEngine(256, 256, cfg=transform.mask_cfg_from(cfg), gaussian_sigma=cfg.gaussian_sigma,
                                 roi_size=tuple(cfg.roi_size), device=dev).run_device(x)
    assert np.array_equal(blur, out.blur.cpu().numpy()) and np.array_equal(mask, out.mask.cpu().numpy())
    assert np.array_equal(boxes, out.info[:, 1:5].cpu().numpy()) and np.array_equal(roi, out.roi.cpu().numpy())
    assert np.array_equal(hist, out.hist9.cpu().numpy())
    analyze, vis = ffi.overlays_batch(x, out.mask, out.info)
    rec = ops.analyze_records(out.mask, out.info, 4096, 512)
    edges = ops.canny(ops.cvt_color(x, "gray"), 80, 160, True)
    assert torch.equal(analyze, ops.analyze_overlay(x, rec, edges, out.mask))
    assert torch.equal(vis, ops.draw_rectangles(x, out.info))
    assert int((analyze != x).any(dim=3).sum()) > 500
