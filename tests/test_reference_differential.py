"""Differential test oracle <-> REFERENCE at the BASELINE sizes (build container only: `needs_reference`).

tests/golden/golden_v1.npz pins the oracle on small images plus one 256x256 leaf; the GPU suite then compares CUDA with
the oracle on many 256x256 / 1024x1024 images.  This file closes the gap in between: the reference's OWN functions
(/root/reference, imported through tests/golden/ref_harness.py) against oracle/spec_* on 64 seeded 256x256 leaves and
4 seeded 1024x1024 leaves -- make_mask for all six deterministic strategies, apply_mask, apply_roi_filter,
apply_blur_filter (+-1 LSB: three float32 normalisations), apply_brown_filter and the six ImageAugmenter ops.
Bit-exact unless noted.  (VERDICT r1 "What's weak" 1.)
"""
import os
import random
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

import make_golden as mg  # noqa: E402
import ref_harness  # noqa: E402
from leaffliction_b200 import synth  # noqa: E402
from oracle import spec_augment as sa  # noqa: E402
from oracle import spec_mask as sm  # noqa: E402

pytestmark = pytest.mark.needs_reference

N256 = int(os.environ.get("LFX_DIFF_N256", "64"))
N1024 = int(os.environ.get("LFX_DIFF_N1024", "4"))


@pytest.fixture(scope="module")
def ns():
    return ref_harness.load()


@pytest.fixture(scope="module")
def leaves256():
    return synth.leaf_batch(N256, 256, 256, seed=97531)


@pytest.fixture(scope="module")
def leaves1024():
    return synth.leaf_batch(N1024, 1024, 1024, seed=86420)


def _check_make_mask(ns, imgs, strategies):
    import cv2
    for strat in strategies:
        cfg = ref_harness.ref_config(ns, mask_strategy=strat)
        scfg = sm.Cfg(mask_strategy=strat)
        for i, img in enumerate(imgs):
            m, cnt = ns.mask.make_mask(img, cfg)
            om, info = sm.make_mask(img, scfg)
            assert np.array_equal(m, om), f"{strat} image {i}: {(m != om).sum()} px differ"
            assert (cnt is None) == (info is None), f"{strat} image {i}"
            if cnt is not None:
                assert tuple(cv2.boundingRect(cnt)) == tuple(info["bbox"]), f"{strat} image {i}"
                assert int(round(2 * cv2.contourArea(cnt))) == info["area2"], f"{strat} image {i}"


@pytest.mark.parametrize("strategy", mg.STRATEGIES)
def test_make_mask_256(ns, leaves256, strategy):
    """mask.py:548-582 under parity profile P0, every deterministic strategy, 64 leaves."""
    _check_make_mask(ns, leaves256, [strategy])


def test_make_mask_1024(ns, leaves1024):
    _check_make_mask(ns, leaves1024, ("hsv_h", "lab", "inclusive"))


def test_filters_256(ns, leaves256):
    """apply_mask (mask_utils.py:10-83), apply_roi_filter (roi.py:20-46), apply_brown_filter (brown.py:21-89),
    apply_blur_filter (blur.py:18-79) on the default-YAML strategy (inclusive) and on hsv_h."""
    for strat in ("inclusive", "hsv_h"):
        cfg = ref_harness.ref_config(ns, mask_strategy=strat)
        scfg = sm.Cfg(mask_strategy=strat)
        for i, img in enumerate(leaves256[:24]):
            m, cnt = ns.mask.make_mask(img, cfg)
            white = ns.mask_utils.apply_mask(img, m, "white")
            assert np.array_equal(white, sm.apply_mask(img, m, "white")), (strat, i)
            assert np.array_equal(ns.mask_utils.apply_mask(img, m, "black"), sm.apply_mask(img, m, "black")), (strat, i)
            if cnt is not None:
                canvas, _vis, box = ns.roi.apply_roi_filter(white, cnt, cfg)
                assert np.array_equal(canvas, sm.roi_letterbox(white, box, cfg.roi_size)), (strat, i)
            _vis, pct, count = ns.brown.apply_brown_filter(white, m, cfg)
            _spots, spct, scount = sm.brown_spots(white, m, scfg)
            assert count == scount and abs(pct - spct) < 1e-9, (strat, i, pct, count, spct, scount)
            if i < 8:   # the slowest reference call (it runs make_mask again inside)
                ref = ns.blur.apply_blur_filter(white, cfg, lambda rgb: ns.mask.make_mask(rgb, cfg))
                m2, _ = sm.make_mask(white, scfg)
                got = sm.saliency_blur(white, m2, scfg)
                assert np.abs(ref.astype(int) - got.astype(int)).max() <= 1, (strat, i)


def _oracle_augment(img, seed):
    h, w = img.shape[:2]
    random.seed(seed)
    out = {"flip": sa.flip(img, sa.draw_flip()), "rotate": sa.rotate_nn(img, sa.draw_rotate()),
           "skew": sa.warp_bicubic(img, sa.skew_coeffs(sa.draw_skew(), w, h), True)}
    k, horiz = sa.draw_shear()
    out["shear"] = sa.warp_bicubic(img, sa.shear_coeffs(k, horiz), False)
    out["crop"] = sa.crop_resize(img, *sa.draw_crop(w, h))
    noise = sa.MT19937(seed).normals(img.size, 0.0, 5.0).reshape(img.shape)
    out["distortion"] = sa.distortion(img, sa.noise_u8(noise), random.uniform(0, 2))
    return out


def test_augment_256(ns, leaves256):
    """ImageAugmenter's six methods verbatim (image_augmenter.py:20-133; file I/O patched to arrays) on 16 leaves,
    a different seed each -- outputs and output SHAPES (rotate expands) equal the oracle's."""
    for i, img in enumerate(leaves256[:16]):
        seed = 1000 + 7919 * i
        ref = mg.run_augment(ns, img, seed)
        got = _oracle_augment(img, seed)
        for t in ("flip", "rotate", "skew", "shear", "crop", "distortion"):
            assert ref[t].shape == got[t].shape, (t, i, ref[t].shape, got[t].shape)
            assert np.array_equal(ref[t], got[t]), f"{t} image {i}: {(ref[t] != got[t]).sum()} bytes differ"


def test_augment_1024(ns, leaves1024):
    img = leaves1024[0]
    ref = mg.run_augment(ns, img, 424242)
    got = _oracle_augment(img, 424242)
    for t in ("flip", "rotate", "skew", "shear", "crop"):
        assert ref[t].shape == got[t].shape and np.array_equal(ref[t], got[t]), t


def test_all_background_mask(ns):
    """a flat image gives an all-zero candidate mask.  With the PlantCV >= 3.14 fill semantics the
    harness restates (skimage.remove_small_objects on a boolean image: a single-valued mask is legal), make_mask never
    raises and returns (mask, None) -- mask.py:66-67, :507-515; the oracle (and the CUDA path through it) pins exactly
    that behaviour.  PlantCV 3.x builds that raise "Image is not binary" on single-valued masks are outside the
    reference's own requirement (requirements.txt:11 plantcv>=3.14 resolves to 4.x)."""
    for val in (0, 255, 128):
        img = np.full((64, 64, 3), val, np.uint8)
        for strat in mg.STRATEGIES:
            cfg = ref_harness.ref_config(ns, mask_strategy=strat)
            m, cnt = ns.mask.make_mask(img, cfg)
            om, info = sm.make_mask(img, sm.Cfg(mask_strategy=strat))
            assert np.array_equal(m, om), (val, strat)
            assert (cnt is None) == (info is None), (val, strat)


def test_auto_strategy_without_kmeans(ns, leaves256):
    """`mask_strategy: auto` (mask.py:435-461) with the k-means candidate disabled on the reference side (it returns an
    empty mask: score -1, never selected; Tier C) against the oracle's six-candidate restatement: per-candidate
    _score_mask values (1e-6: float32 mean vs the same formula) and the final mask / contour statistics (bit-exact)."""
    import dataclasses

    import cv2
    orig = ns.mask._create_kmeans_mask
    ns.mask._create_kmeans_mask = lambda rgb, cfg: np.zeros(rgb.shape[:2], np.uint8)
    try:
        cfg = ref_harness.ref_config(ns, mask_strategy="auto")
        scfg = sm.Cfg(mask_strategy="auto")
        choices = set()
        for i, img in enumerate(list(leaves256[:24]) + [np.full((64, 64, 3), 200, np.uint8)]):
            for st in sm.AUTO_CANDIDATES:
                raw = sm.raw_candidate(img, dataclasses.replace(scfg, mask_strategy=st))
                rm, rcnt = ns.mask._postprocess_mask(raw, cfg)
                om, oinfo = sm.postprocess(raw, scfg)
                assert abs(ns.mask._score_mask(rm, rcnt, img, cfg) - sm.score_mask(om, oinfo, img, scfg)) < 1e-6, (i, st)
            m, cnt = ns.mask.make_mask(img, cfg)
            om, info, choice, _score = sm.make_mask_auto(img, scfg, True, with_kmeans=False)
            choices.add(choice)
            assert np.array_equal(m, om), f"image {i}: {(m != om).sum()} px differ (oracle chose {choice})"
            assert (cnt is None) == (info is None)
            if cnt is not None:
                assert tuple(cv2.boundingRect(cnt)) == tuple(info["bbox"])
        assert len(choices - {None}) >= 2          # the selection rule is exercised, not one strategy winning everywhere
    finally:
        ns.mask._create_kmeans_mask = orig


def test_kmeans_candidate_and_full_auto(ns, leaves256):
    """Tier C row c2 pinned: the reference's `_create_kmeans_mask` (mask.py:109-140: cv2.setRNGSeed(12345) + cv2.kmeans)
    against oracle/spec_kmeans.py -- cv2.kmeans' labels and centres bit for bit, the candidate mask for the three bg_bias
    settings, non-256 sizes (INTER_AREA working copy + INTER_NEAREST back), degenerate images (empty-cluster rule), then
    `mask_strategy: kmeans` and the full seven-candidate `auto` through the reference's make_mask."""
    import dataclasses

    import cv2

    from oracle import spec_kmeans as sk
    crit = (cv2.TERM_CRITERIA_EPS + cv2.TERM_CRITERIA_MAX_ITER, 20, 0.5)
    g = np.random.default_rng(11)
    extra = [np.full((256, 256, 3), 77, np.uint8), np.zeros((64, 256, 3), np.uint8),
             g.integers(0, 256, (256, 200, 3), dtype=np.uint8), (g.integers(0, 2, (256, 256, 1), dtype=np.uint8) * 200).repeat(3, 2)]
    for i, img in enumerate(list(leaves256[:32]) + extra):
        cv2.setRNGSeed(12345)
        _c, labels, centers = cv2.kmeans(img.reshape(-1, 3).astype(np.float32), 3, None, crit, 1, cv2.KMEANS_PP_CENTERS)
        ol, oc, _it = sk.kmeans3(img.reshape(-1, 3))
        assert np.array_equal(labels.ravel(), ol) and np.array_equal(centers, oc), i
    sizes = [(256, 256), (256, 192), (128, 128), (300, 400), (512, 512)]
    from leaffliction_b200 import synth
    for bias in ("light_bg", "dark_bg", "auto"):
        cfg = ref_harness.ref_config(ns, mask_strategy="kmeans", bg_bias=bias)
        scfg = sm.Cfg(mask_strategy="kmeans", bg_bias=bias)
        for k, (h, w) in enumerate(sizes):
            img = synth.leaf_image(40 + k, h, w)
            assert np.array_equal(ns.mask._create_kmeans_mask(img, cfg), sk.kmeans_mask(img, scfg)), (bias, h, w)
        for i, img in enumerate(leaves256[:8]):
            m, cnt = ns.mask.make_mask(img, cfg)
            om, info = sm.make_mask(img, scfg)
            assert np.array_equal(m, om), (bias, i)
            assert (cnt is None) == (info is None)
    cfg = ref_harness.ref_config(ns, mask_strategy="auto")
    scfg = sm.Cfg(mask_strategy="auto")
    choices = set()
    for i, img in enumerate(leaves256[:24]):
        raw = sm.raw_candidate(img, dataclasses.replace(scfg, mask_strategy="kmeans"))
        rm, rcnt = ns.mask._postprocess_mask(raw, cfg)
        omk, oinfo = sm.postprocess(raw, scfg)
        assert abs(ns.mask._score_mask(rm, rcnt, img, cfg) - sm.score_mask(omk, oinfo, img, scfg)) < 1e-6, i
        m, cnt = ns.mask.make_mask(img, cfg)
        om, info, choice, _score = sm.make_mask_auto(img, scfg, True)
        choices.add(choice)
        assert np.array_equal(m, om), f"image {i}: {(m != om).sum()} px differ (oracle chose {choice})"
    assert len(choices - {None}) >= 2
