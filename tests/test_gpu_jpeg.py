"""GPU tests of the nvJPEG file boundary (include/leafx_jpeg.h, leaffliction_b200/jpegio.py).

Parity is to JPEG tolerance: the reference decodes with Pillow / libjpeg-turbo (srcs/utils/image_utils.py:19-47) and
encodes with Pillow quality 95 (:49-59) or cv2.imwrite (srcs/cli/Transformation.py:196-205).  nvJPEG's inverse DCT and
chroma upsampling differ from libjpeg-turbo's by a few LSB, and its forward path by rounding only, so the bounds are:
  decode: mean |nvJPEG - Pillow| <= 2.5 LSB per image (measured: 0.3 for 4:4:4, 1.6-2.0 for noisy 4:2:0 leaves), PSNR(nvJPEG, Pillow) >= 34 dB (4:2:0: nvJPEG replicates chroma samples, libjpeg-turbo
          interpolates them, so single values on chroma edges differ by tens of LSB); 4:4:4 streams: every value within 4 LSB;
  encode: PSNR against the source within 1 dB of Pillow's own encode at the same quality / subsampling, stream length
          within 10 % of Pillow's.
"""
import io

import numpy as np
import pytest
import torch
from PIL import Image

from leaffliction_b200 import balance, jpegio, synth
from leaffliction_b200.cli import Transformation as TC

pytestmark = pytest.mark.gpu

DECODE_MEAN_LSB = 2.5
DECODE_PSNR_DB = 34.0
DECODE_444_MAX_LSB = 4
ENCODE_PSNR_SLACK_DB = 1.0
ENCODE_SIZE_SLACK = 0.10


def _pil_jpeg(arr, quality=95, **kw):
    buf = io.BytesIO()
    Image.fromarray(arr).save(buf, format="JPEG", quality=quality, **kw)
    return buf.getvalue()


def _pil_decode(blob):
    return np.asarray(Image.open(io.BytesIO(blob)).convert("RGB"))


def _psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


def test_decode_batch_matches_pillow_to_jpeg_tolerance():
    imgs = synth.leaf_batch(24, 256, 256)
    blobs = [_pil_jpeg(imgs[i], 95 if i % 2 else 80) for i in range(len(imgs))]
    blobs[3] = _pil_jpeg(imgs[3], 95, subsampling=0)      # 4:4:4
    blobs[5] = _pil_jpeg(imgs[5], 95, subsampling=1)      # 4:2:2
    x, status = jpegio.decode_batch(blobs, 256, 256)
    assert x.is_cuda and tuple(x.shape) == (24, 256, 256, 3)
    assert (status == 0).all()
    got = x.cpu().numpy()
    for i, b in enumerate(blobs):
        d = np.abs(got[i].astype(np.int32) - _pil_decode(b).astype(np.int32))
        assert d.mean() <= DECODE_MEAN_LSB, (i, d.mean())
        assert _psnr(got[i], _pil_decode(b)) >= DECODE_PSNR_DB, (i, _psnr(got[i], _pil_decode(b)))
        if i == 3:       # 4:4:4: no chroma upsampling involved, only the inverse DCT / colour conversion rounding differs
            assert d.mean() <= 0.6 and d.max() <= DECODE_444_MAX_LSB, (d.mean(), d.max())


def test_decode_grey_wrong_size_and_broken_streams():
    imgs = synth.leaf_batch(4, 128, 128)
    grey = np.asarray(Image.fromarray(imgs[0]).convert("L"))
    buf = io.BytesIO()
    Image.fromarray(grey).save(buf, format="JPEG", quality=95)
    blobs = [_pil_jpeg(imgs[1]), buf.getvalue(), _pil_jpeg(synth.leaf_image(9, 64, 64)), b"not a jpeg", None, _pil_jpeg(imgs[2])]
    x, status = jpegio.decode_batch(blobs, 128, 128)
    assert status[0] == 0 and status[1] == 0 and status[5] == 0
    assert status[2] < 0 and status[3] < 0 and status[4] < 0
    got = x.cpu().numpy()
    assert not got[2].any() and not got[3].any() and not got[4].any()          # refused slots stay zero
    g = got[1]
    assert np.array_equal(g[..., 0], g[..., 1]) and np.array_equal(g[..., 1], g[..., 2])   # convert("RGB") of a grey file
    assert np.abs(g[..., 0].astype(np.int32) - _pil_decode(blobs[1])[..., 0].astype(np.int32)).mean() <= DECODE_MEAN_LSB
    for i in (0, 5):
        assert np.abs(got[i].astype(np.int32) - _pil_decode(blobs[i]).astype(np.int32)).mean() <= DECODE_MEAN_LSB
    h, w, nc = jpegio.probe(blobs[2])
    assert (h, w, nc) == (64, 64, 3)


@pytest.mark.parametrize("shape", [(256, 256), (96, 160), (350, 322)])
def test_encode_batch_against_pillow(shape):
    H, W = shape
    imgs = np.stack([synth.leaf_image(i, H, W) for i in range(12)])
    blobs = jpegio.encode_batch(torch.from_numpy(imgs).cuda(), quality=95)
    assert len(blobs) == 12
    for i, b in enumerate(blobs):
        assert b[:2] == b"\xff\xd8" and b[-2:] == b"\xff\xd9"
        with Image.open(io.BytesIO(b)) as im:
            assert im.size == (W, H) and im.mode == "RGB"
        ref = _pil_jpeg(imgs[i], 95)
        p_gpu, p_ref = _psnr(_pil_decode(b), imgs[i]), _psnr(_pil_decode(ref), imgs[i])
        assert p_gpu >= p_ref - ENCODE_PSNR_SLACK_DB, (i, p_gpu, p_ref)
        assert abs(len(b) - len(ref)) <= ENCODE_SIZE_SLACK * len(ref), (i, len(b), len(ref))


def test_encode_decode_round_trip_on_device():
    imgs = synth.leaf_batch(64, 256, 256)
    x = torch.from_numpy(imgs).cuda()
    blobs = jpegio.encode_batch(x, quality=95)
    y, status = jpegio.decode_batch(blobs, 256, 256)
    assert (status == 0).all()
    # the loss of a GPU encode + GPU decode is the codec's own (quality 95, 4:2:0 on noisy leaves), not larger than
    # Pillow's encode + decode of the same image
    for i in (0, 17, 63):
        assert _psnr(y[i].cpu().numpy(), imgs[i]) >= _psnr(_pil_decode(_pil_jpeg(imgs[i], 95)), imgs[i]) - 1.5, i


def _write_leaf(path, idx, size):
    Image.fromarray(synth.leaf_image(idx, size, size)).save(path, quality=95)


def test_dataset_balancer_with_gpu_jpeg(tmp_path):
    """Same plan, same output names and counts as the host-codec path; pixels equal to JPEG tolerance."""
    root = tmp_path / "images"
    spec = {"Apple": {"Apple_healthy": 8, "Apple_scab": 2}, "Grape": {"Grape_spot": 7, "Grape_esca": 1}}
    k = 0
    for plant, classes in spec.items():
        for cls, n in classes.items():
            (root / plant / cls).mkdir(parents=True)
            for i in range(n):
                _write_leaf(root / plant / cls / f"img{i}.JPG", k, 64)
                k += 1
    t_host, t_gpu = tmp_path / "host", tmp_path / "gpu"
    b1 = balance.DatasetBalancer(source_dir=str(root), target_dir=str(t_host), seed=42, workers=2)
    b1.run()
    b2 = balance.DatasetBalancer(source_dir=str(root), target_dir=str(t_gpu), seed=42, workers=2, gpu_jpeg=True)
    b2.run()
    assert b2.completed == b1.completed == 12 and b2.failed == 0
    names1 = sorted(str(p.relative_to(t_host)) for p in t_host.rglob("*_aug_*"))
    names2 = sorted(str(p.relative_to(t_gpu)) for p in t_gpu.rglob("*_aug_*"))
    assert names1 == names2 and len(names1) == 12
    for n in names1:
        a = np.asarray(Image.open(t_host / n).convert("RGB"))
        b = np.asarray(Image.open(t_gpu / n).convert("RGB"))
        assert a.shape == b.shape, n
        if "_aug_distortion_" in n:
            continue      # autocontrast cut-offs are histogram ranks: one LSB of decode difference moves the LUT
        assert _psnr(a, b) > 30.0, (n, _psnr(a, b))


def test_transformation_folder_with_gpu_jpeg(tmp_path, monkeypatch):
    src = tmp_path / "in"
    src.mkdir()
    for i in range(3):
        _write_leaf(src / f"image ({i}).JPG", i, 256)
    _write_leaf(src / "other.JPG", 7, 128)
    (src / "broken.JPG").write_bytes(b"not a jpeg")
    cfgp = tmp_path / "cfg.yaml"
    txt = TC.PACKAGED_CONFIG.read_text().replace("mask_strategy: inclusive", "mask_strategy: hsv_h")
    txt = txt.replace("grabcut_refine: true", "grabcut_refine: false").replace("mask_upscale_factor: 1.3", "mask_upscale_factor: 1.0")
    txt = txt.replace("mask_upscale_long_side: 1500", "mask_upscale_long_side: 0")
    cfgp.write_text(txt)
    d_host, d_gpu = tmp_path / "host", tmp_path / "gpu"
    TC.main(["-src", str(src), "-dst", str(d_host), "--types", "mask,roi,blur", "--config", str(cfgp)])
    monkeypatch.setenv("LEAFX_GPU_JPEG", "1")
    TC.main(["-src", str(src), "-dst", str(d_gpu), "--types", "mask,roi,blur", "--config", str(cfgp)])
    monkeypatch.delenv("LEAFX_GPU_JPEG")
    n1 = sorted(p.name for p in d_host.iterdir())
    n2 = sorted(p.name for p in d_gpu.iterdir())
    assert n1 == n2 and len(n1) == 12 and not any(n.startswith("broken") for n in n1)
    for n in n1:
        if "__T_Mask" in n:
            a = np.asarray(Image.open(d_host / n).convert("RGB"))
            b = np.asarray(Image.open(d_gpu / n).convert("RGB"))
            # masks of the two decodes differ only where a pixel sits on a threshold
            assert np.mean(np.any(a > 127, axis=2) != np.any(b > 127, axis=2)) < 0.03, n
