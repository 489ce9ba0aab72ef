"""GPU parity: every CUDA op, called through the C ABI (libleafx.so via ctypes), against the
CPU oracle on the same seeded inputs.  Bit-exact unless a tolerance is written in the test."""
import random

import numpy as np
import pytest
import torch

from leaffliction_b200 import ops, synth
from oracle import spec_augment as sa
from oracle import spec_color as sc
from oracle import spec_filters as sf
from oracle import spec_mask as sm

pytestmark = pytest.mark.gpu


def up(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def leaf_set(n=6, h=256, w=256):
    return synth.leaf_batch(n, h, w)


SHAPES = [(256, 256), (64, 64), (61, 97), (33, 130)]


# ----------------------------------------------------------------------------- colour
def test_cvt_color_exhaustive(dev):
    img = sc.all_colours()[None]                      # all 2^24 colours
    x = up(img, dev)
    assert np.array_equal(ops.cvt_color(x, "gray").cpu().numpy()[0], sc.rgb_to_gray(img[0]))
    assert np.array_equal(ops.cvt_color(x, "hsv").cpu().numpy()[0], sc.rgb_to_hsv(img[0]))
    assert np.array_equal(ops.cvt_color(x, "lab").cpu().numpy()[0], sc.rgb_to_lab(img[0]))


@pytest.mark.parametrize("hw", SHAPES)
def test_cvt_color_ragged(dev, hw):
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (3, *hw, 3), dtype=np.uint8)
    x = up(img, dev)
    for code, f in (("gray", sc.rgb_to_gray), ("hsv", sc.rgb_to_hsv), ("lab", sc.rgb_to_lab)):
        got = ops.cvt_color(x, code).cpu().numpy()
        for i in range(3):
            assert np.array_equal(got[i], f(img[i])), code


@pytest.mark.parametrize("strategy", ["hsv_h", "lab"])
def test_threshold_mask(dev, strategy):
    imgs = np.concatenate([leaf_set(4), np.random.default_rng(1).integers(0, 256, (2, 256, 256, 3), dtype=np.uint8)])
    cfg = ops.mask_cfg(strategy=strategy)
    got = ops.threshold_mask(up(imgs, dev), cfg).cpu().numpy()
    scfg = sm.Cfg(mask_strategy=strategy)
    for i in range(len(imgs)):
        exp = sm.mask_hsv_green(imgs[i], scfg) if strategy == "hsv_h" else sm.mask_lab(imgs[i])
        assert np.array_equal(got[i], exp)


def test_threshold_mask_exhaustive(dev):
    img = sc.all_colours()[None]
    x = up(img, dev)
    assert np.array_equal(ops.threshold_mask(x, ops.mask_cfg("hsv_h")).cpu().numpy()[0], sm.mask_hsv_green(img[0], sm.Cfg()))
    assert np.array_equal(ops.threshold_mask(x, ops.mask_cfg("lab")).cpu().numpy()[0], sm.mask_lab(img[0]))


@pytest.mark.parametrize("hw", SHAPES)
def test_apply_mask(dev, hw):
    rng = np.random.default_rng(4)
    img = rng.integers(0, 256, (3, *hw, 3), dtype=np.uint8)
    mask = rng.integers(0, 256, (3, *hw), dtype=np.uint8)   # non-binary: exercises the >127 rule
    for val, name in ((255, "white"), (0, "black")):
        got = ops.apply_mask(up(img, dev), up(mask, dev), val).cpu().numpy()
        for i in range(3):
            assert np.array_equal(got[i], sm.apply_mask(img[i], mask[i], name))


@pytest.mark.parametrize("hw", SHAPES)
def test_color_stats(dev, hw):
    rng = np.random.default_rng(5)
    imgs = synth.leaf_batch(3, *hw)
    imgs[2] = rng.integers(0, 256, (*hw, 3), dtype=np.uint8)
    masks = np.stack([sm.mask_hsv_green(im, sm.Cfg()) for im in imgs])
    h9, h3, cn = ops.color_stats(up(imgs, dev), up(masks, dev))
    h9, h3, cn = h9.cpu().numpy(), h3.cpu().numpy(), cn.cpu().numpy()
    for i in range(3):
        masked = sm.apply_mask(imgs[i], masks[i], "white")
        assert np.array_equal(h9[i], sm.hist9(imgs[i], masks[i]))
        assert np.array_equal(h3[i], sm.hsv_hist_leaf(masked))
        assert np.array_equal(cn[i, :14], sm.hist_counters(masked))
        assert h9[i, :3].sum(axis=1).tolist() == [int((masks[i] > 0).sum())] * 3   # property: sums = masked px
    # no mask: whole-image histograms (PIL histogram inside autocontrast)
    h9n, _, _ = ops.color_stats(up(imgs, dev), None, True, False, False)
    for i in range(3):
        assert np.array_equal(h9n.cpu().numpy()[i], sm.hist9(imgs[i], None))


# ----------------------------------------------------------------------------- Gaussian
@pytest.mark.parametrize("hw", SHAPES + [(300, 500), (48, 640), (40, 1024), (33, 512)])     # the last three: column-tiled TMA kernel
@pytest.mark.parametrize("ks", [(5, 1.5), (15, 0.0), (3, 0.0), (7, 2.0)])
def test_gauss(dev, hw, ks):
    rng = np.random.default_rng(6)
    k, s = ks
    img = rng.integers(0, 256, (2, *hw, 3), dtype=np.uint8)
    got = ops.gauss_u8(up(img, dev), k, s).cpu().numpy()
    for i in range(2):
        assert np.array_equal(got[i], sf.gaussian_blur_u8(img[i], k, s))      # bit-exact (north star allows +-1)
    gray = img[..., 0].copy()
    got = ops.gauss_u8(up(gray, dev), k, s).cpu().numpy()
    assert np.array_equal(got[0], sf.gaussian_blur_u8(gray[0], k, s))


# TMA-tiled blur (rows that are a multiple of 16 bytes): sizes that cross tile borders, tall / wide / tiny images
@pytest.mark.parametrize("hw", [(256, 256), (64, 64), (100, 48), (17, 32), (300, 128), (64, 1024), (9, 16)])
@pytest.mark.parametrize("ks", [(5, 1.5), (15, 0.0), (5, 0.8)])
def test_gauss_tma_shapes(dev, hw, ks):
    rng = np.random.default_rng(hw[0] * 7 + ks[0])
    k, s = ks
    img = rng.integers(0, 256, (3, *hw, 3), dtype=np.uint8)
    img[1] = 255
    img[1, ::2, ::3] = 0                                   # worst case for the packed 16-bit sums
    got = ops.gauss_u8(up(img, dev), k, s).cpu().numpy()
    for i in range(3):
        assert np.array_equal(got[i], sf.gaussian_blur_u8(img[i], k, s)), (hw, ks, i)
    gray = np.ascontiguousarray(img[..., 1])
    got = ops.gauss_u8(up(gray, dev), k, s).cpu().numpy()
    for i in range(3):
        assert np.array_equal(got[i], sf.gaussian_blur_u8(gray[i], k, s)), (hw, ks, i, "gray")


# ----------------------------------------------------------------------------- augment
@pytest.mark.parametrize("hw", SHAPES)
def test_flip(dev, hw):
    rng = np.random.default_rng(7)
    img = rng.integers(0, 256, (4, *hw, 3), dtype=np.uint8)
    lr = [True, False, False, True]
    got = ops.flip(up(img, dev), lr).cpu().numpy()
    for i in range(4):
        assert np.array_equal(got[i], sa.flip(img[i], lr[i]))
    # property: flip o flip = id
    again = ops.flip(up(got, dev), lr).cpu().numpy()
    assert np.array_equal(again, img)


@pytest.mark.parametrize("hw", SHAPES)
def test_rotate(dev, hw):
    rng = np.random.default_rng(8)
    h, w = hw
    angles = [-30.0, 30.0, 17.1235, 3.3, -0.0001, 29.999, -12.5]
    img = rng.integers(0, 256, (len(angles), h, w, 3), dtype=np.uint8)
    params = []
    for a in angles:
        m, nw, nh = sa.rotate_params(a, w, h)
        params.append(list(sa.affine_fixed_coeffs(m)) + [nw, nh])
    slab, stride = ops.rotate_nn(up(img, dev), np.array(params, np.int32))
    slab = slab.cpu().numpy()
    for i, a in enumerate(angles):
        nw, nh = params[i][6], params[i][7]
        got = slab[i, :nh * nw * 3].reshape(nh, nw, 3)
        assert np.array_equal(got, sa.rotate_nn(img[i], a)), a


@pytest.mark.parametrize("hw", SHAPES)
def test_warp_bicubic(dev, hw):
    rng = np.random.default_rng(9)
    h, w = hw
    random.seed(11)
    coeffs, persp = [], []
    for _ in range(4):
        coeffs.append(sa.skew_coeffs(sa.draw_skew(), w, h)); persp.append(True)
        coeffs.append(sa.shear_coeffs(*sa.draw_shear())); persp.append(False)
    n = len(coeffs)
    img = rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)
    img[0] = 255; img[1] = 0; img[2, : h // 2] = 255          # flat / saturated regions hit the exact path
    got = ops.warp_bicubic(up(img, dev), np.array(coeffs), persp).cpu().numpy()
    for i in range(n):
        assert np.array_equal(got[i], sa.warp_bicubic(img[i], coeffs[i], persp[i])), i   # bit-exact (allowed +-1)


@pytest.mark.parametrize("hw", [(40, 300), (64, 515), (96, 1024), (33, 259)])
def test_warp_bicubic_tiles_and_general_maps(dev, hw):
    """Several column tiles / slices per band (W > 256, odd widths), general affine maps (both axes mixed) and true
    perspective maps (taps from global memory) -- against Pillow itself."""
    from PIL import Image
    rng = np.random.default_rng(hw[0] * 1000 + hw[1])
    h, w = hw
    random.seed(hw[1])
    coeffs, persp = [], []
    coeffs.append(sa.skew_coeffs(0.11, w, h)); persp.append(True)
    coeffs.append(sa.shear_coeffs(0.2, True)); persp.append(False)
    coeffs.append(sa.shear_coeffs(-0.2, False)); persp.append(False)                    # tall source span: column slices
    coeffs.append([0.9, 0.25, 3.0, -0.3, 1.05, 7.5, 0.0, 0.0]); persp.append(False)      # general affine
    coeffs.append([1.02, 0.03, -2.0, 0.01, 0.97, 1.0, 1e-4, -2e-4]); persp.append(True)   # true perspective
    coeffs.append([1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0]); persp.append(False)          # identity: dx = dy = 0
    img = rng.integers(0, 256, (len(coeffs), h, w, 3), dtype=np.uint8)
    got = ops.warp_bicubic(up(img, dev), np.array(coeffs, np.float64), persp).cpu().numpy()
    for i, (co, p) in enumerate(zip(coeffs, persp)):
        exp = np.asarray(Image.fromarray(img[i]).transform((w, h), Image.Transform.PERSPECTIVE if p else Image.Transform.AFFINE,
                                                         co if p else co[:6], Image.Resampling.BICUBIC))
        assert np.array_equal(got[i], exp), (hw, i, int((got[i] != exp).sum()))


@pytest.mark.parametrize("hw", SHAPES)
def test_crop_lanczos(dev, hw):
    rng = np.random.default_rng(10)
    h, w = hw
    random.seed(12)
    boxes = [sa.draw_crop(w, h) for _ in range(5)] + [(0, 0, w, h)]
    img = rng.integers(0, 256, (len(boxes), h, w, 3), dtype=np.uint8)
    got = ops.crop_lanczos(up(img, dev), np.array(boxes), (h, w)).cpu().numpy()
    for i, b in enumerate(boxes):
        assert np.array_equal(got[i], sa.crop_resize(img[i], *b)), b


def test_resize_normalize_224(dev):
    imgs = leaf_set(3)
    boxes = np.array([(0, 0, 256, 256)] * 3)
    u8, f32 = ops.crop_lanczos(up(imgs, dev), boxes, (224, 224), want_f32=True)
    for i in range(3):
        eu8, ef = sa.resize_normalize(imgs[i], 224)
        assert np.array_equal(u8.cpu().numpy()[i], eu8)
        assert np.array_equal(f32.cpu().numpy()[i], ef)        # x/255.0f is exact in fp32


@pytest.mark.parametrize("hw", SHAPES)
def test_distort(dev, hw):
    h, w = hw
    imgs = synth.leaf_batch(3, h, w)
    seeds = [42, 7, 999983]
    noises, cuts, cutoffs = [], [], []
    for s in seeds:
        noises.append(sa.noise_u8(sa.MT19937(s).normals(h * w * 3, 0, 5).reshape(h, w, 3)))
        random.seed(s)
        c = random.uniform(0, 2)
        cutoffs.append(c)
        cuts.append(int(h * w * c // 100))
    noise = np.stack(noises)
    got = ops.distort(up(imgs, dev), up(noise, dev), cuts).cpu().numpy()
    for i in range(3):
        assert np.array_equal(got[i], sa.distortion(imgs[i], noise[i], cutoffs[i]))


# ----------------------------------------------------------------------------- mask path
def _mask_cases():
    cases = [("leaf%d" % i, im) for i, im in enumerate(leaf_set(6))]
    cases += [("adv64_" + k, v) for k, v in synth.adversarial_images(64, 64).items()]
    cases += [("adv_61x97_" + k, v) for k, v in synth.adversarial_images(61, 97).items()]
    return cases


@pytest.mark.parametrize("fill", [1000, 50])
def test_postprocess_mask(dev, fill):
    for name, im in _mask_cases():
        raw = sm.mask_hsv_green(im, sm.Cfg())
        got, info = ops.postprocess_mask(up(raw[None], dev), fill, 3)
        exp, einfo = sm.postprocess(raw, sm.Cfg(fill_size=fill))
        info = info.cpu().numpy()[0]
        assert np.array_equal(got.cpu().numpy()[0], exp), name
        assert bool(info[0]) == (einfo is not None), name
        if einfo is not None:
            assert tuple(info[1:5]) == einfo["bbox"] and info[5] == einfo["area2"] and info[6] == einfo["npix"], name


@pytest.mark.parametrize("strategy", ["hsv_h", "lab", "hsv_s", "hsv_v_dark"])
@pytest.mark.parametrize("fill", [1000, 50])
def test_make_mask(dev, strategy, fill):
    cfg = ops.mask_cfg(strategy=strategy, fill_size=fill)
    scfg = sm.Cfg(mask_strategy=strategy, fill_size=fill)
    by_shape = {}
    for name, im in _mask_cases():
        by_shape.setdefault(im.shape, []).append((name, im))
    for shape, items in by_shape.items():
        batch = np.stack([im for _, im in items])
        mask, info = ops.make_mask(up(batch, dev), cfg)
        mask, info = mask.cpu().numpy(), info.cpu().numpy()
        for i, (name, im) in enumerate(items):
            exp, einfo = sm.make_mask(im, scfg)
            assert np.array_equal(mask[i], exp), (strategy, name, int((mask[i] != exp).sum()))
            assert bool(info[i, 0]) == (einfo is not None), (strategy, name)
            if einfo is not None:
                assert tuple(info[i, 1:5]) == einfo["bbox"] and info[i, 5] == einfo["area2"], (strategy, name)


def test_make_mask_external_raw(dev):
    """strategy 'external': the inclusive front end computed by the oracle, post-processed on the GPU."""
    imgs = leaf_set(3)
    scfg = sm.Cfg(mask_strategy="inclusive")
    raws = np.stack([sm.mask_inclusive(im, scfg) for im in imgs])
    cfg = ops.mask_cfg(strategy="external")
    mask, info = ops.make_mask(up(imgs, dev), cfg, raw=up(raws, dev))
    for i in range(3):
        exp, einfo = sm.make_mask(imgs[i], scfg)
        assert np.array_equal(mask.cpu().numpy()[i], exp)
        assert tuple(info.cpu().numpy()[i, 1:5]) == einfo["bbox"]


def test_roi_letterbox(dev):
    imgs = leaf_set(8)
    cfg = ops.mask_cfg("hsv_h")
    x = up(imgs, dev)
    mask, info = ops.make_mask(x, cfg)
    roi = ops.roi_letterbox(x, mask, info, (256, 256)).cpu().numpy()
    for i in range(len(imgs)):
        m, einfo = sm.make_mask(imgs[i], sm.Cfg(mask_strategy="hsv_h"))
        exp = sm.roi_letterbox(sm.apply_mask(imgs[i], m, "white"), einfo["bbox"], (256, 256))
        assert np.array_equal(roi[i], exp), i                     # bit-exact (north star allows +-1)


def test_roi_no_contour(dev):
    imgs = np.stack([synth.adversarial_images(64, 64)["black"]] * 2)
    x = up(imgs, dev)
    mask, info = ops.make_mask(x, ops.mask_cfg("hsv_h"))
    assert info.cpu().numpy()[:, 0].tolist() == [0, 0]
    assert int(ops.roi_letterbox(x, mask, info, (64, 64)).sum()) == 0


def test_pipeline_core(dev):
    imgs = leaf_set(8)
    cfg = ops.mask_cfg("hsv_h")
    out = ops.pipeline_core(up(imgs, dev), cfg, 1.5, (256, 256))
    scfg = sm.Cfg(mask_strategy="hsv_h")
    for i in range(len(imgs)):
        m, einfo = sm.make_mask(imgs[i], scfg)
        masked = sm.apply_mask(imgs[i], m, "white")
        assert np.array_equal(out.blur.cpu().numpy()[i], sf.gaussian_blur_u8(imgs[i], 5, 1.5))
        assert np.array_equal(out.mask.cpu().numpy()[i], m)
        assert tuple(out.info.cpu().numpy()[i, 1:5]) == einfo["bbox"]
        assert np.array_equal(out.roi.cpu().numpy()[i], sm.roi_letterbox(masked, einfo["bbox"], (256, 256)))
        assert np.array_equal(out.hist9.cpu().numpy()[i], sm.hist9(imgs[i], m))
        assert np.array_equal(out.hsv3.cpu().numpy()[i], sm.hsv_hist_leaf(masked))
        assert np.array_equal(out.counters.cpu().numpy()[i, :14], sm.hist_counters(masked))


def _check_core(out, imgs, scfg, roi_size):
    for i in range(len(imgs)):
        m, einfo = sm.make_mask(imgs[i], scfg)
        masked = sm.apply_mask(imgs[i], m, "white")
        assert np.array_equal(out.blur[i].cpu().numpy(), sf.gaussian_blur_u8(imgs[i], 5, scfg.gaussian_sigma)), f"blur {i}"
        assert np.array_equal(out.mask[i].cpu().numpy(), m), f"mask {i}"
        if einfo is None:
            assert int(out.info[i, 0]) == 0, f"found {i}"
            assert int(out.roi[i].sum()) == 0, f"roi (no contour) {i}"
        else:
            assert tuple(out.info[i, 1:5].cpu().numpy()) == einfo["bbox"], f"bbox {i}"
            assert np.array_equal(out.roi[i].cpu().numpy(), sm.roi_letterbox(masked, einfo["bbox"], roi_size)), f"roi {i}"
        assert np.array_equal(out.hist9[i].cpu().numpy(), sm.hist9(imgs[i], m)), f"hist9 {i}"
        assert np.array_equal(out.hsv3[i].cpu().numpy(), sm.hsv_hist_leaf(masked)), f"hsv3 {i}"
        assert np.array_equal(out.counters[i, :14].cpu().numpy(), sm.hist_counters(masked)), f"counters {i}"


# fused kernel (W % 32 == 0, H*W <= 65536) and the general path (other shapes) through the same entry point
@pytest.mark.parametrize("hw,roi", [((256, 256), (256, 256)), ((64, 64), (64, 64)), ((96, 64), (128, 64)), ((40, 32), (48, 48)),
                                    ((35, 128), (64, 128)), ((128, 512), (128, 512)), ((61, 97), (64, 100))])
@pytest.mark.parametrize("strategy", ["hsv_h", "lab"])
def test_pipeline_core_shapes(dev, hw, roi, strategy):
    imgs = synth.leaf_batch(5, hw[0], hw[1], seed=77)
    cfg = ops.mask_cfg(strategy, fill_size=60)
    out = ops.pipeline_core(up(imgs, dev), cfg, 1.5, roi)
    _check_core(out, imgs, sm.Cfg(mask_strategy=strategy, fill_size=60, roi_size=roi), roi)


@pytest.mark.parametrize("strategy,bias", [("hsv_s", "light_bg"), ("hsv_s", "dark_bg"), ("hsv_v_dark", "light_bg")])
@pytest.mark.parametrize("hw", [(256, 256), (96, 64), (61, 97)])
def test_pipeline_core_otsu_strategies(dev, strategy, bias, hw):
    """Otsu candidates (mask.py:76-84) through the fused kernel (extra histogram + threshold passes in phase B) and the
    general path (61x97)."""
    imgs = synth.leaf_batch(4, hw[0], hw[1], seed=5)
    cfg = ops.mask_cfg(strategy, fill_size=60, bg_bias=bias)
    out = ops.pipeline_core(up(imgs, dev), cfg, 1.5, (hw[0], hw[1] if hw[1] % 16 == 0 else 112))
    roi = (hw[0], hw[1] if hw[1] % 16 == 0 else 112)
    _check_core(out, imgs, sm.Cfg(mask_strategy=strategy, fill_size=60, bg_bias=bias, roi_size=roi), roi)


def test_pipeline_core_adversarial(dev):
    adv = synth.adversarial_images(64, 64)
    imgs = np.stack(list(adv.values()))
    cfg = ops.mask_cfg("hsv_h", fill_size=20)
    out = ops.pipeline_core(up(imgs, dev), cfg, 1.5, (64, 64))
    _check_core(out, imgs, sm.Cfg(mask_strategy="hsv_h", fill_size=20, roi_size=(64, 64)), (64, 64))


@pytest.mark.parametrize("kw", [dict(extend_brown=False), dict(use_lab_brown=True), dict(gaussian_sigma=0.8),
                                dict(brown_hue_range=(5, 40), brown_min_area_px=5), dict(hsv_channel_for_mask="v")])
def test_pipeline_core_configs(dev, kw):
    imgs = np.concatenate([leaf_set(3), np.stack([synth.adversarial_images(256, 256)[k] for k in ("black", "salt", "frame")])])
    sigma = kw.pop("gaussian_sigma", 1.5)
    extend = kw.pop("extend_brown", True)
    cfg = ops.mask_cfg("hsv_h", extend_brown=extend, **kw)
    out = ops.pipeline_core(up(imgs, dev), cfg, sigma, (256, 256))
    scfg = sm.Cfg(mask_strategy="hsv_h", gaussian_sigma=sigma, **kw)
    if extend:
        _check_core(out, imgs, scfg, (256, 256))
    else:  # the spec always extends; compare the pre-extension stage
        for i in range(len(imgs)):
            raw = sm.raw_candidate(imgs[i], scfg)
            m, einfo = sm.postprocess(raw, scfg)
            if einfo is None or einfo["area2"] <= 2:
                m, einfo = sm.postprocess(sm.mask_hsv_otsu(imgs[i], scfg.hsv_channel_for_mask, "light"), scfg)
            assert np.array_equal(out.mask[i].cpu().numpy(), m)


def test_pipeline_core_many_images(dev):
    """More images than resident thread blocks (dynamic image queue), each checked against the batch-1 result."""
    base = synth.leaf_batch(24, 256, 256, seed=5)
    imgs = np.concatenate([base] * 30)          # 720 images > 296 resident blocks
    out = ops.pipeline_core(up(imgs, dev), ops.mask_cfg("hsv_h"), 1.5, (256, 256))
    ref = ops.pipeline_core(up(base, dev), ops.mask_cfg("hsv_h"), 1.5, (256, 256))
    for name in ("blur", "mask", "roi", "hist9", "hsv3", "counters"):
        a = getattr(out, name).view(30, 24, *getattr(ref, name).shape[1:])
        assert bool((a == getattr(ref, name)[None]).all()), name
    assert bool((out.info.view(30, 24, 8)[:, :, :7] == ref.info[None, :, :7]).all())
    _check_core(ref, base[:4], sm.Cfg(mask_strategy="hsv_h"), (256, 256))


@pytest.mark.parametrize("seed", [42, 7, 999983, 1, 123456])
def test_device_legacy_normal_stream(dev, seed):
    """GPU MT19937 + polar gauss == np.random.seed(seed); np.random.normal(0, 5, n).astype(np.uint8) (image_augmenter.py:121-123)."""
    n = 256 * 256 * 3
    got = ops.legacy_normal_noise([seed], n, 5.0, dev).cpu().numpy()[0]
    np.random.seed(seed)
    exp = np.random.normal(0, 5, n).astype(np.uint8)
    assert np.array_equal(got, exp), int((got != exp).sum())
    assert np.array_equal(got, sa.noise_u8(sa.MT19937(seed).normals(n, 0.0, 5.0)))   # and == the oracle's restatement


def test_device_legacy_normal_batch_and_sizes(dev):
    seeds = [3, 999999, 31337, 2**31 + 5, 17, 65536, 8, 77, 1000000]
    for n in (1, 2, 311, 64 * 64 * 3):
        got = ops.legacy_normal_noise(seeds, n, 5.0, dev).cpu().numpy()
        for i, sd in enumerate(seeds):
            np.random.seed(sd)
            assert np.array_equal(got[i], np.random.normal(0, 5, n).astype(np.uint8)), (n, sd)


def test_device_legacy_normal_stress(dev):
    """25 M samples (128 streams x 196,608): the fp32 fast path with its fp64 re-evaluation must reproduce NumPy's
    fp64 stream bit for bit; also the general (loc, scale) form, which always runs in fp64."""
    n = 256 * 256 * 3
    seeds = [1000 + 7919 * i for i in range(128)]
    got = ops.legacy_normal_noise(seeds, n, 5.0, dev).cpu().numpy()
    bad = 0
    for i, sd in enumerate(seeds):
        np.random.seed(sd)
        bad += int((got[i] != np.random.normal(0, 5, n).astype(np.uint8)).sum())
    assert bad == 0, bad
    for loc, scale in ((0.0, 8.0), (0.0, 0.37), (100.0, 20.0), (3.5, 5.0)):
        got = ops.legacy_normal_noise([5, 99], 40001, scale, dev, loc=loc).cpu().numpy()     # odd n: unaligned second row
        for i, sd in enumerate((5, 99)):
            np.random.seed(sd)
            assert np.array_equal(got[i], np.random.normal(loc, scale, 40001).astype(np.uint8)), (loc, scale, sd)


def test_empty_batch(dev):
    x = torch.empty((0, 64, 64, 3), dtype=torch.uint8, device=dev)
    assert ops.cvt_color(x, "hsv").shape == (0, 64, 64, 3)
    assert ops.gauss_u8(x, 5, 1.5).shape == (0, 64, 64, 3)
    m, info = ops.make_mask(x, ops.mask_cfg("hsv_h"))
    assert m.shape == (0, 64, 64) and info.shape == (0, 8)


# ----------------------------------------------------------------------------- register-only kernels: tails, unaligned buffers
def _unaligned(arr, dev, off=1):
    """Device tensor holding `arr` at an odd byte offset from the allocation (forces the shared-tile fallback kernels)."""
    flat = torch.empty(arr.size + 16, dtype=torch.uint8, device=dev)
    view = flat[off:off + arr.size].view(arr.shape)
    view.copy_(torch.from_numpy(np.ascontiguousarray(arr)).to(dev))
    assert view.data_ptr() % 16 != 0
    return view


@pytest.mark.parametrize("hw", [(7, 5), (64, 64), (33, 130)])
def test_colour_kernels_tail_and_unaligned(dev, hw):
    rng = np.random.default_rng(hw[0])
    imgs = rng.integers(0, 256, (3, *hw, 3), dtype=np.uint8)
    masks = rng.integers(0, 2, (3, *hw), dtype=np.uint8) * 255
    masks[1] = rng.integers(0, 256, hw, dtype=np.uint8)                  # non-binary mask: > 127 rule
    for x, m in ((up(imgs, dev), up(masks, dev)), (_unaligned(imgs, dev), _unaligned(masks, dev, 3))):
        for code, fn in (("gray", sc.rgb_to_gray), ("hsv", sc.rgb_to_hsv), ("lab", sc.rgb_to_lab)):
            got = ops.cvt_color(x, code).cpu().numpy()
            assert np.array_equal(got, np.stack([fn(im) for im in imgs])), code
        assert np.array_equal(ops.threshold_mask(x, ops.mask_cfg("hsv_h")).cpu().numpy(),
                              np.stack([sm.mask_hsv_green(im, sm.Cfg()) for im in imgs]))
        for color in (255, 0):
            exp = np.stack([sm.apply_mask(im, mk, "white" if color else "black") for im, mk in zip(imgs, masks)])
            assert np.array_equal(ops.apply_mask(x, m, color).cpu().numpy(), exp)


def test_color_stats_large_image_counter_flush(dev):
    """1024 x 1024: every thread sees > 128 pixels, so the packed 8-bit category counters are flushed many times."""
    img = synth.leaf_image(4, 1024, 1024)
    mask = sm.mask_hsv_green(img, sm.Cfg())
    h9, h3, cn = ops.color_stats(up(img[None], dev), up(mask[None], dev))
    masked = sm.apply_mask(img, mask, "white")
    assert np.array_equal(h9.cpu().numpy()[0], sm.hist9(img, mask))
    assert np.array_equal(h3.cpu().numpy()[0], sm.hsv_hist_leaf(masked))
    assert np.array_equal(cn.cpu().numpy()[0, :14], sm.hist_counters(masked))


@pytest.mark.parametrize("hw,roi", [((64, 64), (96, 512)), ((96, 64), (128, 64)), ((61, 97), (300, 100))])
def test_roi_letterbox_canvas_shapes(dev, hw, roi):
    """Canvases wider than the thread block (two column passes), taller than one row block, and the general-shape path."""
    imgs = synth.leaf_batch(4, hw[0], hw[1], seed=31)
    cfg = ops.mask_cfg("hsv_h", fill_size=60)
    x = up(imgs, dev)
    mask, info = ops.make_mask(x, cfg)
    got = ops.roi_letterbox(x, mask, info, roi).cpu().numpy()
    for i in range(len(imgs)):
        m, einfo = sm.make_mask(imgs[i], sm.Cfg(mask_strategy="hsv_h", fill_size=60))
        if einfo is None:
            assert int(got[i].sum()) == 0
        else:
            assert np.array_equal(got[i], sm.roi_letterbox(sm.apply_mask(imgs[i], m, "white"), einfo["bbox"], roi)), i
