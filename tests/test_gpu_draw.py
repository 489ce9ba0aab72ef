"""GPU tests of the overlay rasterisers (lfx_analyze_overlay, lfx_draw_rectangles) through the C ABI:
against the golden overlays produced by the reference's own apply_analyze_filter / apply_roi_filter
(tests/golden/golden_draw_v1.npz), against the oracle (oracle/spec_draw.py) on batches, and against cv2 itself."""
import os

import numpy as np
import pytest
import torch

from leaffliction_b200 import engine, filters, ops, synth, transform
from oracle import spec_draw as sd

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def up(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def _golden():
    g = np.load(os.path.join(ROOT, "tests", "golden", "golden_draw_v1.npz"))
    return g, sorted({k.rsplit("_rgb", 1)[0] for k in g.files if k.endswith("_rgb")})


def test_overlays_against_reference_golden(dev):
    """Single-image drop-in functions on the stored inputs == the reference's stored outputs, bit for bit."""
    g, keys = _golden()
    cfg = transform.default_config()
    for k in keys:
        rgb, mask, c = g[k + "_rgb"], g[k + "_mask"], g[k + "_contour"]
        got = filters.apply_analyze_filter(rgb, mask, c, cfg)
        assert np.array_equal(got, g[k + "_analyze"]), k
        _, vis, box = transform.apply_roi_filter(rgb, c, cfg)
        assert tuple(box) == tuple(int(v) for v in g[k + "_roi_box"])
        assert np.array_equal(vis, g[k + "_roi_vis"]), k


def test_draw_rectangles_batch_vs_oracle(dev):
    """Boxes inside the image, touching and crossing its border (x + w == W is the usual case: boundingRect is inclusive),
    boxes not found: a batch against the oracle."""
    rng = np.random.default_rng(3)
    for H, W in ((48, 64), (256, 256)):
        B = 24
        imgs = rng.integers(0, 256, size=(B, H, W, 3), dtype=np.uint8)
        info = np.zeros((B, 8), np.int32)
        for i in range(B):
            x, y = int(rng.integers(0, W - 2)), int(rng.integers(0, H - 2))
            w, h = int(rng.integers(1, W - x + 1)), int(rng.integers(1, H - y + 1))
            if i % 5 == 0:
                w, h = W - x, H - y
            if i == 1:
                x, y, w, h = 0, 0, W, H
            info[i, :5] = (0 if i == 7 else 1, x, y, w, h)
        got = ops.draw_rectangles(up(imgs, dev), up(info, dev)).cpu().numpy()
        for i in range(B):
            exp = imgs[i].copy()
            if info[i, 0]:
                sd.rectangle2(exp, *[int(v) for v in info[i, 1:5]])
            assert np.array_equal(got[i], exp), (H, W, i, info[i])


def _batch_overlay(imgs, dev, strategy="hsv_h"):
    x = up(imgs, dev)
    mask, info = ops.make_mask(x, ops.mask_cfg(strategy))
    rec = ops.analyze_records(mask, info, max_pts=4096, max_hull=512)
    edges = ops.canny(ops.cvt_color(x, "gray"), 80, 160, True)
    return x, mask, info, rec, edges, ops.analyze_overlay(x, rec, edges, mask)


def _expected(img, mask, edges, pts, ri, rf, hull):
    r = filters.record_from_device(ri, rf, hull)
    r["hull"] = sd.hull_in_cv_order(pts, r["hull"][:, 0, :])
    return sd.analyze_overlay(img, pts, r, (edges > 0) & (mask > 0))


@pytest.mark.parametrize("hw,n,seed", [((96, 96), 10, 5), ((256, 256), 6, 17), ((64, 100), 4, 2), ((384, 512), 2, 9)])
def test_analyze_overlay_batch_vs_oracle(dev, hw, n, seed):
    """A whole batch in one launch against the oracle drawing from the SAME record (so that the rasterisation, the block-level
    parallel order and the hull re-ordering are what is tested; the record itself is tested in test_gpu_round2)."""
    H, W = hw
    imgs = synth.leaf_batch(n, H, W, seed)
    x, mask, info, rec, edges, over = _batch_overlay(imgs, dev)
    over, mask_h, edges_h = over.cpu().numpy(), mask.cpu().numpy(), edges.cpu().numpy()
    pts, cnt = rec["points"].cpu().numpy(), rec["counts"].cpu().numpy()
    ri, rf, hull = rec["rec_i"].cpu().numpy(), rec["rec_f"].cpu().numpy(), rec["hull"].cpu().numpy()
    seen = 0
    for i in range(n):
        if ri[i, 0] == 0:
            assert np.array_equal(over[i], imgs[i])
            continue
        seen += 1
        exp = _expected(imgs[i], mask_h[i], edges_h[i], pts[i, :cnt[i]], ri[i], rf[i], hull[i])
        assert np.array_equal(over[i], exp), (i, int((over[i] != exp).any(2).sum()))
    assert seen >= n // 2


def test_analyze_overlay_adversarial_and_cv2(dev):
    """Contours on the image border (frame), thin / pinched shapes, salt noise, and images without any contour; where cv2 is
    importable the same overlay is also drawn by the reference's cv2 calls."""
    adv = synth.adversarial_images(64, 64)
    names = sorted(adv)
    imgs = np.stack([adv[k] for k in names] + [np.full((64, 64, 3), 200, np.uint8)])
    x, mask, info, rec, edges, over = _batch_overlay(imgs, dev)
    over, mask_h, edges_h = over.cpu().numpy(), mask.cpu().numpy(), edges.cpu().numpy()
    pts, cnt = rec["points"].cpu().numpy(), rec["counts"].cpu().numpy()
    ri, rf, hull = rec["rec_i"].cpu().numpy(), rec["rec_f"].cpu().numpy(), rec["hull"].cpu().numpy()
    assert ri[-1, 0] == 0 and np.array_equal(over[-1], imgs[-1])
    try:
        import cv2
    except ImportError:
        cv2 = None
    drawn = 0
    for i in range(len(imgs)):
        if ri[i, 0] == 0:
            assert np.array_equal(over[i], imgs[i])
            continue
        drawn += 1
        p = pts[i, :cnt[i]]
        exp = _expected(imgs[i], mask_h[i], edges_h[i], p, ri[i], rf[i], hull[i])
        assert np.array_equal(over[i], exp), names[i] if i < len(names) else i
        if cv2 is not None and len(p) >= 3:
            c = p.reshape(-1, 1, 2).astype(np.int32)
            hv = cv2.convexHull(c)[:, 0, :]
            mine = sd.hull_in_cv_order(p, filters.record_from_device(ri[i], rf[i], hull[i])["hull"][:, 0, :])
            if len(hv) == len(mine):
                assert np.array_equal(hv, mine), ("hull order", i)
    assert drawn >= 3


def test_engine_overlays_device(dev):
    """TransformEngine.overlays_device: both overlay images of a folder run for a batch, on the masked image."""
    imgs = synth.leaf_batch(8, 256, 256, 77)
    eng = engine.TransformEngine(256, 256, device=dev)
    x = up(imgs, dev)
    out = eng.run_device(x)
    over, vis = eng.overlays_device(x, out)
    masked = ops.apply_mask(x, out.mask, 255)
    info = out.info.cpu().numpy()
    masked_h, vis_h = masked.cpu().numpy(), vis.cpu().numpy()
    for i in range(len(imgs)):
        exp = masked_h[i].copy()
        if info[i, 0]:
            sd.rectangle2(exp, *[int(v) for v in info[i, 1:5]])
        assert np.array_equal(vis_h[i], exp), i
    rec = eng.analyze_device(out)
    edges = ops.canny(ops.cvt_color(masked, "gray"), 80, 160, True)
    assert torch.equal(over, ops.analyze_overlay(masked, rec, edges, out.mask))
    assert int((over != masked).any(dim=3).sum()) > 1000


def test_no_contour_is_a_copy(dev):
    """analyze.py:28-29 without a contour: the image comes back unchanged (the reference adds a text banner, which is not drawn)."""
    img = synth.leaf_image(0, 64, 64)
    got = filters.apply_analyze_filter(img, None, None, transform.default_config())
    assert np.array_equal(got, img) and got is not img


def test_draw_primitives_random_lists(dev):
    """lfx_draw_primitives: lists of every kind (2-4 px lines, anti-aliased lines, filled circles, rectangles, cross markers), end
    points inside and outside the image, overlapping, ragged counts: each image against the oracle drawing the same list in order."""
    from test_draw_hostsim_cpu import random_primitives
    rng = np.random.default_rng(8)
    for (B, P, H, W) in ((48, 12, 40, 56), (6, 30, 256, 256)):
        imgs = rng.integers(0, 256, size=(B, H, W, 3), dtype=np.uint8)
        prims = random_primitives(rng, B, P, H, W, margin=12)
        counts = rng.integers(0, P + 1, size=B).astype(np.int32)
        counts[0] = P
        got = ops.draw_primitives(up(imgs, dev), prims, counts).cpu().numpy()
        for b in range(B):
            exp = sd.draw_primitives(imgs[b].copy(), prims[b, :counts[b]])
            assert np.array_equal(got[b], exp), (H, W, b, prims[b, :counts[b]].tolist())


def test_draw_primitives_every_direction(dev):
    """One anti-aliased line, then one 2-px line, from the image centre to EVERY pixel of a 41 x 41 window around it (all slopes,
    all octants, the degenerate point; part of the window lies outside the image): 1681 images, each against the oracle."""
    H = W = 36
    c = 17
    targets = [(c + dx, c + dy) for dy in range(-20, 21) for dx in range(-20, 21)]
    B = len(targets)
    rng = np.random.default_rng(12)
    base = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
    imgs = np.repeat(base[None], B, axis=0)
    prims = np.zeros((B, 2, 8), np.int32)
    for b, (x, y) in enumerate(targets):
        prims[b, 0] = (ops.DRAW_LINE_AA, c, c, x, y, 0x20C0FF, 1, 0)
        prims[b, 1] = (ops.DRAW_LINE, x, y, c - 3, c + 2, 0xFF4010, 2, 0)
    got = ops.draw_primitives(up(imgs, dev), prims).cpu().numpy()
    for b in range(B):
        exp = sd.draw_primitives(base.copy(), prims[b])
        assert np.array_equal(got[b], exp), targets[b]
