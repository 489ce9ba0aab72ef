"""GPU tests of the two CLIs and the dataset balancer end to end (JPEG in, JPEG out)."""
import io
import random

import numpy as np
import pytest
from PIL import Image

from leaffliction_b200 import balance, synth
from leaffliction_b200.cli import Augmentation as A
from leaffliction_b200.cli import Transformation as TC
from oracle import spec_augment as sa
from oracle import spec_mask as sm

pytestmark = pytest.mark.gpu


def _jpeg_roundtrip(arr, quality=95):
    buf = io.BytesIO()
    Image.fromarray(arr).save(buf, format="JPEG", quality=quality)
    return np.asarray(Image.open(io.BytesIO(buf.getvalue())).convert("RGB"))


def _write_leaf(path, idx, size=128):
    Image.fromarray(synth.leaf_image(idx, size, size)).save(path, quality=95)
    return np.asarray(Image.open(path).convert("RGB"))


def test_augmentation_cli_single_image(tmp_path):
    src = tmp_path / "leaf.jpg"
    img = _write_leaf(src, 3)
    out = tmp_path / "out"
    A.main([str(src), "-out", str(out), "-seed", "42"])
    assert (out / "original_leaf.jpg").read_bytes() == src.read_bytes()
    # the oracle with the reference's draw order, then the same JPEG encode
    random.seed(42)
    h, w = img.shape[:2]
    exp = {"flip": sa.flip(img, sa.draw_flip()), "rotate": sa.rotate_nn(img, sa.draw_rotate())}
    exp["skew"] = sa.warp_bicubic(img, sa.skew_coeffs(sa.draw_skew(), w, h), True)
    k, horiz = sa.draw_shear()
    exp["shear"] = sa.warp_bicubic(img, sa.shear_coeffs(k, horiz), False)
    exp["crop"] = sa.crop_resize(img, *sa.draw_crop(w, h))
    noise = sa.MT19937(42).normals(img.size, 0.0, 5.0).reshape(img.shape)
    exp["distortion"] = sa.distortion(img, sa.noise_u8(noise), random.uniform(0, 2))
    for t in A.TRANSFORMATIONS:
        got = np.asarray(Image.open(out / f"{t}_leaf.jpg").convert("RGB"))
        assert np.array_equal(got, _jpeg_roundtrip(exp[t])), t


def test_transformation_cli_single_and_folder(tmp_path):
    src = tmp_path / "in"
    src.mkdir()
    imgs = [_write_leaf(src / f"image ({i}).JPG", i, 256) for i in range(3)]
    cfgp = tmp_path / "cfg.yaml"
    txt = TC.PACKAGED_CONFIG.read_text().replace("mask_strategy: inclusive", "mask_strategy: hsv_h")
    txt = txt.replace("grabcut_refine: true", "grabcut_refine: false").replace("mask_upscale_factor: 1.3", "mask_upscale_factor: 1.0")
    txt = txt.replace("mask_upscale_long_side: 1500", "mask_upscale_long_side: 0")
    cfgp.write_text(txt)
    dst = tmp_path / "out"
    TC.main(["-src", str(src), "-dst", str(dst), "--types", "mask,roi,brown,hist", "--config", str(cfgp)])
    scfg = sm.Cfg(mask_strategy="hsv_h")
    for i, img in enumerate(imgs):
        m, info = sm.make_mask(img, scfg)
        got = np.asarray(Image.open(dst / f"image ({i})__T_Mask.jpg").convert("RGB"))
        exp = sm.apply_mask(img, m, "black")
        # cv2-style JPEG write in the reference vs Pillow here: compare before encoding tolerance-free is not
        # possible, so check against the same encoder
        assert np.array_equal(got, _jpeg_roundtrip(exp, quality=95)), i
        assert (dst / f"image ({i})__T_ROI.jpg").exists() and (dst / f"image ({i})__T_Brown.jpg").exists()
        assert (dst / f"image ({i})__T_Hist.json").exists()
    # --skip-existing keeps files, single-image mode writes next to --out-dir
    before = (dst / "image (0)__T_Mask.jpg").stat().st_mtime_ns
    TC.main(["-src", str(src), "-dst", str(dst), "--types", "mask", "--config", str(cfgp), "--skip-existing"])
    assert (dst / "image (0)__T_Mask.jpg").stat().st_mtime_ns == before
    one = tmp_path / "one"
    TC.main([str(src / "image (1).JPG"), "--out-dir", str(one), "--types", "Mask", "--config", str(cfgp)])
    assert (one / "image (1)__T_Mask.jpg").exists()


def test_dataset_balancer_balances(tmp_path):
    root = tmp_path / "images"
    spec = {"Apple": {"Apple_healthy": 5, "Apple_scab": 2}, "Grape": {"Grape_spot": 4, "Grape_esca": 1}}
    k = 0
    for plant, classes in spec.items():
        for cls, n in classes.items():
            (root / plant / cls).mkdir(parents=True)
            for i in range(n):
                _write_leaf(root / plant / cls / f"img{i}.JPG", k, 64)
                k += 1
    target = tmp_path / "balanced"
    b = balance.DatasetBalancer(source_dir=str(root), target_dir=str(target), seed=42, workers=2)
    b.run()
    rows = balance.count_images(target)
    assert {(p, c): n for p, c, n in rows} == {("Apple", "Apple_healthy"): 5, ("Apple", "Apple_scab"): 5,
                                               ("Grape", "Grape_esca"): 4, ("Grape", "Grape_spot"): 4}
    assert b.completed == 6 and b.failed == 0
    names = sorted(p.name for p in (target / "Apple" / "Apple_scab").iterdir())
    assert sum("_aug_" in n for n in names) == 3


def test_transformation_cli_default_config_large_and_broken_images(tmp_path):
    """the reference's DEFAULT config (mask_strategy inclusive, roi_size 256) on images larger than 256x256
    (512x512: front-end planes in global scratch, ROI shrink path), mixed shapes in one folder, plus a file that cannot
    be decoded: the folder run logs it and goes on (Transformation.py:700-705), the other images get all their outputs."""
    src = tmp_path / "in"
    src.mkdir()
    big = _write_leaf(src / "big.JPG", 1, 512)
    small = _write_leaf(src / "small.JPG", 2, 256)
    (src / "broken.JPG").write_bytes(b"not a jpeg")
    cfgp = tmp_path / "cfg.yaml"
    txt = TC.PACKAGED_CONFIG.read_text().replace("grabcut_refine: true", "grabcut_refine: false")
    txt = txt.replace("mask_upscale_factor: 1.3", "mask_upscale_factor: 1.0").replace("mask_upscale_long_side: 1500", "mask_upscale_long_side: 0")
    cfgp.write_text(txt)
    dst = tmp_path / "out"
    TC.main(["-src", str(src), "-dst", str(dst), "--types", "mask,roi,analyze,brown,blur", "--config", str(cfgp)])
    scfg = sm.Cfg(mask_strategy="inclusive")
    for name, img in (("big", big), ("small", small)):
        m, _info = sm.make_mask(img, scfg)
        got = np.asarray(Image.open(dst / f"{name}__T_Mask.jpg").convert("RGB"))
        assert np.array_equal(got, _jpeg_roundtrip(sm.apply_mask(img, m, "black"), quality=95)), name
        for t in ("ROI", "Analyze", "Brown", "Blur"):
            assert (dst / f"{name}__T_{t}.jpg").exists(), (name, t)
    assert not list(dst.glob("broken__T_*"))
