"""The reference's own third-party calls on in-memory arrays (no JPEG I/O).

Each function performs the same Pillow / OpenCV / NumPy call the reference makes at the
cited line, on arrays instead of files.  Used (a) to cross-check ``spec_*`` at test time on
both the build container and the GPU box (same image, same library versions) and (b) as
the timed CPU baseline in bench.py (``cpu_baseline`` and ``--impl reference``).
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
from __future__ import annotations

import numpy as np
from PIL import Image, ImageOps


def flip(img, left_right):                       # image_augmenter.py:24,26
    im = Image.fromarray(img)
    t = Image.FLIP_LEFT_RIGHT if left_right else Image.FLIP_TOP_BOTTOM
    return np.asarray(im.transpose(t))


def rotate(img, angle):                          # image_augmenter.py:37
    return np.asarray(Image.fromarray(img).rotate(angle, expand=True, fillcolor="white"))


def warp(img, coeffs, perspective):              # image_augmenter.py:61-66, 84-89
    im = Image.fromarray(img)
    if perspective:
        return np.asarray(im.transform(im.size, Image.PERSPECTIVE, list(coeffs[:8]), Image.BICUBIC))
    return np.asarray(im.transform(im.size, Image.AFFINE, list(coeffs[:6]), Image.BICUBIC))


def crop_resize(img, left, top, nw, nh):         # image_augmenter.py:108-109
    im = Image.fromarray(img)
    w, h = im.size
    return np.asarray(im.crop((left, top, left + nw, top + nh)).resize((w, h), Image.LANCZOS))


def distortion(img, noise_f64, cutoff):          # image_augmenter.py:121-127
    noise = noise_f64.astype(np.uint8)
    arr = np.clip(img + noise, 0, 255)
    return np.asarray(ImageOps.autocontrast(Image.fromarray(arr), cutoff=cutoff))


def resize_normalize(img, size=224):             # sequence.py:84-88
    u8 = np.asarray(Image.fromarray(img).resize((size, size), Image.Resampling.LANCZOS))
    return u8, u8.astype(np.float32) / 255.0
