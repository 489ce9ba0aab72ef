"""GPU tests at the sizes of BASELINE.json's other configs: config 4 (1024x1024 warps + blur), config 5 (fused
resize-224 + /255 normalise handed over via DLPack), config 3 (device-resident class balancing with the device
MT19937 noise) -- each against the oracle / the reference's own library calls."""
import random

import numpy as np
import pytest
import torch

from leaffliction_b200 import augment, balance, ops, synth
from oracle import refcalls as rc
from oracle import spec_augment as sa
from oracle import spec_filters as sf

pytestmark = pytest.mark.gpu


def up(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def test_config4_1024_warps_and_blur(dev):
    img = synth.leaf_image(3, 1024, 1024)
    x = up(img[None], dev)
    # blur: bit-exact against cv2.GaussianBlur's fixed-point path (oracle) at both kernel sizes
    assert np.array_equal(ops.gauss_u8(x, 5, 1.5).cpu().numpy()[0], sf.gaussian_blur_u8(img, 5, 1.5))
    assert np.array_equal(ops.gauss_u8(x, 15, 0.0).cpu().numpy()[0], sf.gaussian_blur_u8(img, 15, 0.0))
    # flip / rotate / skew / shear / crop against Pillow itself (refcalls = the reference's own calls)
    assert np.array_equal(ops.flip(x, [True]).cpu().numpy()[0], rc.flip(img, True))
    assert np.array_equal(ops.flip(x, [False]).cpu().numpy()[0], rc.flip(img, False))
    ang = -23.4567
    m, nw, nh = augment.rotate_matrix(ang, 1024, 1024)
    slab, _ = ops.rotate_nn(x, np.array([augment.fixed_affine(m) + [nw, nh]], np.int32))
    assert np.array_equal(slab.cpu().numpy()[0, : nh * nw * 3].reshape(nh, nw, 3), rc.rotate(img, ang))
    skew = sa.skew_coeffs(0.0831, 1024, 1024)
    assert np.array_equal(ops.warp_bicubic(x, np.array([skew]), [True]).cpu().numpy()[0], rc.warp(img, skew, True))
    for k, horiz in ((0.171, True), (-0.139, False)):     # vertical shear spans too many rows for the staged band: global path
        sh = sa.shear_coeffs(k, horiz)
        assert np.array_equal(ops.warp_bicubic(x, np.array([sh]), [False]).cpu().numpy()[0], rc.warp(img, sh, False)), (k, horiz)
    box = (37, 101, 870, 870)
    assert np.array_equal(ops.crop_lanczos(x, np.array([box], np.int32), (1024, 1024)).cpu().numpy()[0], rc.crop_resize(img, *box))


def test_config5_resize_normalize_dlpack(dev):
    imgs = synth.leaf_batch(8, 256, 256, seed=9)
    x = up(imgs, dev)
    u8, f32 = ops.crop_lanczos(x, np.tile(np.array([0, 0, 256, 256], np.int32), (8, 1)), (224, 224), want_f32=True)
    cap = torch.utils.dlpack.to_dlpack(f32)              # the hand-over train.py's input pipeline would consume
    back = torch.utils.dlpack.from_dlpack(cap)
    assert back.data_ptr() == f32.data_ptr() and back.shape == (8, 224, 224, 3) and back.dtype == torch.float32
    for i in range(8):
        eu8, ef = rc.resize_normalize(imgs[i], 224)       # PIL Lanczos + /255.0 (sequence.py:84-88)
        assert np.array_equal(u8[i].cpu().numpy(), eu8)
        assert np.array_equal(back[i].cpu().numpy(), ef)
    # 1024 -> 224 (29-tap downscale, general kernel)
    big = synth.leaf_image(1, 1024, 1024)
    u8b, f32b = ops.crop_lanczos(up(big[None], dev), np.array([[0, 0, 1024, 1024]], np.int32), (224, 224), want_f32=True)
    eu8, ef = rc.resize_normalize(big, 224)
    assert np.array_equal(u8b[0].cpu().numpy(), eu8) and np.array_equal(f32b[0].cpu().numpy(), ef)


def test_config3_device_balancing_matches_reference_calls(dev):
    """tasks_for_labels -> augment_device (images in HBM, noise from the device MT19937) == the reference's Pillow /
    NumPy calls with the parameters a fresh ImageAugmenter(seed) would draw for each task."""
    counts = {"PlantA": {"A_c0": 12, "A_c1": 5}, "PlantB": {"B_c0": 7, "B_c1": 9}}
    names = [c for p in counts.values() for c in p]
    plants = {p: list(c) for p, c in counts.items()}
    labels = np.repeat(np.arange(len(names)), [n for p in counts.values() for n in p.values()])
    imgs = synth.leaf_batch(len(labels), 96, 96, seed=21)
    plan, tasks = balance.tasks_for_labels(labels, names, plants, seed=42)
    assert sum(sum(v.values()) for v in plan.values()) == (12 - 5) + (9 - 7) == len(tasks)
    res = augment.augment_device(up(imgs, dev), tasks, device_noise=True)
    seen = 0
    for key, val in res.items():
        ids = val[0]
        for k, ti in enumerate(ids):
            t = tasks[ti]
            src = imgs[t.source_index]
            p = augment.draw_task_params(t.transform_name, t.seed, 96, 96)
            if t.transform_name == "flip":
                exp = rc.flip(src, p[0])
            elif t.transform_name == "rotate":
                exp = rc.rotate(src, p[0])
            elif t.transform_name in ("skew", "shear"):
                exp = rc.warp(src, p[0], p[1])
            elif t.transform_name == "crop":
                exp = rc.crop_resize(src, *p)
            else:
                np.random.seed(t.seed)
                noise = np.random.normal(0, 5, src.shape)
                exp = rc.distortion(src, noise, p[1])
            if key == "rotate":
                nh, nw = val[2][k]
                got = val[1][k, : nh * nw * 3].view(nh, nw, 3).cpu().numpy()
            else:
                got = val[1][k].cpu().numpy()
            assert got.shape == exp.shape and np.array_equal(got, exp), (t.transform_name, t.seed)
            seen += 1
    assert seen == len(tasks)


def test_warp_bicubic_stress_against_pillow(dev):
    """6.3 M interpolated values on pure-noise images (the worst case for the fp32 fast path: taps alternate over the
    whole range) must equal Pillow's fp64 result bit for bit -- the fp32 path may only be used where it provably
    cannot change the truncated byte."""
    rng = np.random.default_rng(2024)
    imgs = rng.integers(0, 256, (32, 256, 256, 3), dtype=np.uint8)
    imgs[0] = (np.indices((256, 256)).sum(0) % 2 * 255)[..., None]          # checkerboard 0/255
    prng = random.Random(5)
    coeffs, persp = [], []
    for i in range(32):
        if i % 2 == 0:
            coeffs.append(sa.skew_coeffs(prng.uniform(0.05, 0.15), 256, 256)); persp.append(True)
        else:
            coeffs.append(sa.shear_coeffs(prng.uniform(-0.2, 0.2), prng.random() < 0.5)); persp.append(False)
    got = ops.warp_bicubic(up(imgs, dev), np.array(coeffs, np.float64), persp).cpu().numpy()
    for i in range(32):
        exp = rc.warp(imgs[i], coeffs[i], persp[i])
        assert np.array_equal(got[i], exp), (i, int((got[i] != exp).sum()))
