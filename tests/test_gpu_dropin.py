"""GPU tests of the drop-in Python surface (same signatures as the reference) and of contour tracing."""
import random

import numpy as np
import pytest

from leaffliction_b200 import augment, synth, transform
from oracle import refcalls as rc
from oracle import spec_augment as sa
from oracle import spec_contour as spc
from oracle import spec_mask as sm

pytestmark = pytest.mark.gpu


def _cases():
    c = [synth.leaf_image(i) for i in range(5)]
    c += list(synth.adversarial_images(64, 64).values())
    c += list(synth.adversarial_images(61, 97).values())
    return c


@pytest.mark.parametrize("strategy", ["hsv_h", "lab"])
def test_make_mask_contour(strategy):
    """make_mask returns the same (mask, contour) as the spec; the contour equals
    cv2.findContours(RETR_EXTERNAL, CHAIN_APPROX_SIMPLE) of the largest component."""
    cv2 = pytest.importorskip("cv2")
    cfg = transform.default_config(mask_strategy=strategy, grabcut_refine=False, mask_upscale_factor=1.0,
                                   mask_upscale_long_side=0, fill_size=50)
    scfg = sm.Cfg(mask_strategy=strategy, fill_size=50)
    for im in _cases():
        mask, cnt = transform.make_mask(im, cfg)
        emask, info = sm.make_mask(im, scfg)
        assert np.array_equal(mask, emask)
        assert (cnt is None) == (info is None)
        if cnt is None:
            continue
        cnts, _ = cv2.findContours(emask, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        ref = max(cnts, key=cv2.contourArea)
        assert cnt.dtype == np.int32 and cnt.shape == ref.shape and np.array_equal(cnt, ref)
        assert transform.bounding_rect(cnt) == tuple(cv2.boundingRect(ref)) == info["bbox"]


def test_trace_contour_moments(dev):
    import torch
    from leaffliction_b200 import ops
    imgs = synth.leaf_batch(6)
    x = torch.from_numpy(imgs).to(dev)
    mask, info = ops.make_mask(x, ops.mask_cfg("hsv_h"))
    pts, cnt, sums = ops.trace_contour(mask, info, 4096)
    pts, cnt, sums, info_h = pts.cpu().numpy(), cnt.cpu().numpy(), sums.cpu().numpy(), info.cpu().numpy()
    for i in range(len(imgs)):
        m, _ = sm.make_mask(imgs[i], sm.Cfg(mask_strategy="hsv_h"))
        start = (int(info_h[i, 7] >> 8), int(info_h[i, 2]))
        exp = spc.trace_external(m, start)
        got = pts[i, :cnt[i]].reshape(-1, 1, 2)
        assert np.array_equal(got, exp)
        m00, m10, m01 = spc.moments_polygon(exp)
        a00, a10, a01 = (float(v) for v in sums[i])
        sgn = 1.0 if a00 > 0 else -1.0
        assert (a00 * 0.5 * sgn, a10 * (sgn * 0.16666666666666666), a01 * (sgn * 0.16666666666666666)) == (m00, m10, m01)
    # buffer too small -> negative count = -needed
    _, cnt2, _ = ops.trace_contour(mask, info, 8)
    assert (cnt2.cpu().numpy() == -cnt).all()


def test_apply_mask_errors_and_values():
    im = synth.leaf_image(0)
    m = sm.mask_hsv_green(im, sm.Cfg())
    assert np.array_equal(transform.apply_mask(im, m, "white"), sm.apply_mask(im, m, "white"))
    assert np.array_equal(transform.apply_mask(im, m, "BLACK"), sm.apply_mask(im, m, "black"))
    before = im.copy()
    transform.apply_mask(im, m)
    assert np.array_equal(im, before)                       # inputs are never mutated
    with pytest.raises(ValueError):
        transform.apply_mask(im, m, "red")
    with pytest.raises(TypeError):
        transform.apply_mask([1, 2], m)
    with pytest.raises(ValueError):
        transform.apply_mask(im, np.zeros((2, 2, 2, 2), np.uint8))


def test_roi_filter_dropin():
    cfg = transform.default_config(mask_strategy="hsv_h")
    for i in range(4):
        im = synth.leaf_image(i)
        mask, cnt = transform.make_mask(im, cfg)
        masked = transform.apply_mask(im, mask, "white")
        canvas, vis, bbox = transform.apply_roi_filter(masked, cnt, cfg)
        assert bbox == transform.bounding_rect(cnt)
        assert np.array_equal(canvas, sm.roi_letterbox(masked, bbox, cfg.roi_size))
        assert vis.shape == im.shape
    canvas, vis, bbox = transform.apply_roi_filter(im, None, cfg)     # roi.py:23-24: no contour -> (rgb, None, None)
    assert np.array_equal(canvas, im) and vis is None and bbox is None


def test_image_augmenter_files(tmp_path):
    """File-in/file-out drop-in: same seeds -> same parameters as the reference's call sequence
    (Augmentation.py single_image_mode), pixel results equal to Pillow's on the decoded image."""
    from PIL import Image
    im = synth.leaf_image(3)
    src = tmp_path / "leaf.jpg"
    Image.fromarray(im).save(src, quality=95)
    dec = np.array(Image.open(src).convert("RGB"))
    aug = augment.ImageAugmenter(seed=42)
    outs = {}
    for t in augment.TRANSFORMATIONS:
        out = tmp_path / f"{t}_leaf.jpg"
        assert getattr(aug, t)(str(src), str(out)) is True
        outs[t] = out
    # replay the reference's draws with the library calls (oracle/refcalls.py) and compare the JPEGs
    random.seed(42); np.random.seed(42)
    exp = {}
    exp["flip"] = rc.flip(dec, random.choice([True, False]))
    exp["rotate"] = rc.rotate(dec, random.uniform(-30, 30))
    s = random.uniform(0.05, 0.15)
    exp["skew"] = rc.warp(dec, sa.skew_coeffs(s, 256, 256), True)
    k = random.uniform(-0.2, 0.2); hz = random.choice([True, False])
    exp["shear"] = rc.warp(dec, sa.shear_coeffs(k, hz), False)
    r = random.uniform(0.8, 0.95); nw, nh = int(256 * r), int(256 * r)
    left = random.randint(0, 256 - nw); top = random.randint(0, 256 - nh)
    exp["crop"] = rc.crop_resize(dec, left, top, nw, nh)
    noise = np.random.normal(0, 5, dec.shape)
    exp["distortion"] = rc.distortion(dec, noise, random.uniform(0, 2))
    for t in augment.TRANSFORMATIONS:
        ref_path = tmp_path / f"ref_{t}.jpg"
        Image.fromarray(exp[t]).save(ref_path, quality=95)
        assert ref_path.read_bytes() == outs[t].read_bytes(), t      # identical pixels -> identical JPEG bytes
    # error convention: never raises, returns False
    assert aug.flip(str(tmp_path / "missing.jpg"), str(tmp_path / "o.jpg")) is False
    assert aug.flip(str(tmp_path / "x.png"), str(tmp_path / "o.jpg")) is False


def test_augment_arrays_matches_per_task_reference():
    imgs = [synth.leaf_image(i) for i in range(12)]
    transforms = [augment.TRANSFORMATIONS[i % 6] for i in range(12)]
    seeds = [1000 + 37 * i for i in range(12)]
    got = augment.augment_arrays(imgs, transforms, seeds)
    for i in range(12):
        random.seed(seeds[i]); np.random.seed(seeds[i])
        t = transforms[i]
        if t == "flip":
            e = rc.flip(imgs[i], random.choice([True, False]))
        elif t == "rotate":
            e = rc.rotate(imgs[i], random.uniform(-30, 30))
        elif t == "skew":
            e = rc.warp(imgs[i], sa.skew_coeffs(random.uniform(0.05, 0.15), 256, 256), True)
        elif t == "shear":
            k = random.uniform(-0.2, 0.2); hz = random.choice([True, False])
            e = rc.warp(imgs[i], sa.shear_coeffs(k, hz), False)
        elif t == "crop":
            e = rc.crop_resize(imgs[i], *sa.draw_crop(256, 256))
        else:
            noise = np.random.normal(0, 5, imgs[i].shape)
            e = rc.distortion(imgs[i], noise, random.uniform(0, 2))
        assert np.array_equal(got[i], e), (i, t)


def test_analyze_record_hull_and_pca():
    """analyze.py:43-98 numeric record: centroid of the contour polygon, extreme points, convex hull, PCA axes --
    against OpenCV on the same contour."""
    cv2 = pytest.importorskip("cv2")
    from leaffliction_b200 import filters
    cfg = transform.default_config(mask_strategy="hsv_h", grabcut_refine=False, mask_upscale_factor=1.0, mask_upscale_long_side=0)
    for i in range(4):
        im = synth.leaf_image(i)
        mask, cnt = transform.make_mask(im, cfg)
        rec = filters.analyze_record(im, mask, cnt)
        M = cv2.moments(cnt)
        assert rec["centroid"] == (int(M["m10"] / M["m00"]), int(M["m01"] / M["m00"]))
        hull = cv2.convexHull(cnt)
        assert {tuple(p) for p in hull[:, 0, :]} == {tuple(p) for p in rec["hull"][:, 0, :]}
        data = cnt[:, 0, :].astype(np.float32)
        mean, evec, evals = cv2.PCACompute2(data, mean=None)
        assert np.allclose(rec["pca_mean"], mean[0], atol=1e-3)
        assert np.allclose(rec["pca_eigenvalues"], evals[:, 0], rtol=1e-4)
        for k in range(2):
            assert abs(abs(float(np.dot(rec["pca_eigenvectors"][k], evec[k]))) - 1.0) < 1e-4       # same axis up to sign
            proj = data @ evec[k]
            got = {float(np.dot(np.array(p, np.float32), evec[k])) for p in rec["axes"][k]}
            assert abs(min(got) - float(proj.min())) < 1e-2 and abs(max(got) - float(proj.max())) < 1e-2


def test_resize_cubic_and_nearest_kernels():
    """lfx_resize_cubic == the oracle's restatement of OpenCV's 8-bit cubic path (bit-exact) and within 1 LSB of cv2 itself;
    lfx_resize_nearest == cv2 exactly (mask.py:29-50, 526-545)."""
    import torch
    cv2 = pytest.importorskip("cv2")
    from leaffliction_b200 import ops
    from oracle import spec_filters as sf
    d = torch.device("cuda:0")
    rng = np.random.default_rng(8)
    for (h, w), (oh, ow) in (((256, 256), (333, 333)), ((61, 97), (79, 126)), ((64, 64), (200, 150)), ((40, 32), (40, 32))):
        imgs = np.stack([synth.leaf_image(i, h, w) for i in range(2)] + [rng.integers(0, 256, (h, w, 3), dtype=np.uint8)])
        got = ops.resize_cubic(torch.from_numpy(imgs).to(d), (oh, ow)).cpu().numpy()
        for i in range(len(imgs)):
            assert np.array_equal(got[i], sf.resize_cubic_u8(imgs[i], (ow, oh))), (h, w, i)
            ref = cv2.resize(imgs[i], (ow, oh), interpolation=cv2.INTER_CUBIC).astype(int)
            assert np.abs(got[i].astype(int) - ref).max() <= 1                      # the +-1 LSB class of SURVEY A.12
        m = (rng.integers(0, 2, (3, oh, ow), dtype=np.uint8) * 255)
        back = ops.resize_nearest(torch.from_numpy(m).to(d), (h, w)).cpu().numpy()
        for i in range(3):
            assert np.array_equal(back[i], cv2.resize(m[i], (w, h), interpolation=cv2.INTER_NEAREST))
        rgb_back = ops.resize_nearest(torch.from_numpy(got).to(d), (h, w)).cpu().numpy()
        assert np.array_equal(rgb_back[0], cv2.resize(got[0], (w, h), interpolation=cv2.INTER_NEAREST))


def test_make_mask_with_cubic_upscale_close_to_reference_calls():
    """mask_upscale_factor 1.3 (the reference YAML's value): upscale -> mask -> nearest back.  The working image is only
    +-1 LSB from cv2's, so the mask is compared by IoU with the same steps done by OpenCV + the oracle."""
    cv2 = pytest.importorskip("cv2")
    from oracle import spec_mask as sm
    cfg = transform.default_config(mask_strategy="hsv_h", grabcut_refine=False, mask_upscale_factor=1.3, mask_upscale_long_side=1500)
    native, _ = transform.make_mask(synth.leaf_image(0), cfg)              # default: profile P0, the keys are ignored
    assert np.array_equal(native, sm.make_mask(synth.leaf_image(0), sm.Cfg(mask_strategy="hsv_h"))[0])
    transform.set_upscale(True)
    try:
        _check_upscaled_masks(cv2, cfg)
    finally:
        transform.set_upscale(None)


def _check_upscaled_masks(cv2, cfg):
    for i in range(3):
        im = synth.leaf_image(i)
        mask, cnt = transform.make_mask(im, cfg)
        assert mask.shape == im.shape[:2] and cnt is not None
        work = cv2.resize(im, (333, 333), interpolation=cv2.INTER_CUBIC)
        m_work, _ = sm.make_mask(work, sm.Cfg(mask_strategy="hsv_h"))
        exp = cv2.resize(m_work, (256, 256), interpolation=cv2.INTER_NEAREST)
        inter = np.logical_and(mask > 0, exp > 0).sum()
        union = np.logical_or(mask > 0, exp > 0).sum()
        assert inter / union > 0.995, (i, inter / union)
        x, y, w, h = transform.bounding_rect(cnt)
        ys, xs = np.nonzero(mask)
        assert abs(x - xs.min()) <= 2 and abs(y - ys.min()) <= 2 and abs(x + w - 1 - xs.max()) <= 2 and abs(y + h - 1 - ys.max()) <= 2
