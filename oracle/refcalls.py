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


AUGMENT_OPS = ("flip", "rotate", "skew", "shear", "crop", "distortion")


def augment_task(img, op, seed):
    """One balancing task on an array: `_process_single_transformation` (dataset_balancer.py:201-207) builds
    ImageAugmenter(seed) (seeds `random` and `np.random` when seed is truthy, image_augmenter.py:16-18), then the
    method draws its parameters in the reference's order and makes the Pillow / NumPy calls."""
    import random
    if seed:
        random.seed(seed)
        np.random.seed(seed)
    h, w = img.shape[:2]
    if op == "flip":
        return flip(img, random.choice([True, False]))                                  # :23
    if op == "rotate":
        return rotate(img, random.uniform(-30, 30))                                     # :36
    if op == "skew":
        s = random.uniform(0.05, 0.15)                                                  # :48
        return warp(img, [1 + s, 0, -s * w, 0, 1 + s, -s * h, 0, 0], True)
    if op == "shear":
        k = random.uniform(-0.2, 0.2)                                                   # :77
        c = [1, k, 0, 0, 1, 0, 0, 0] if random.choice([True, False]) else [1, 0, 0, k, 1, 0, 0, 0]
        return warp(img, c, False)
    if op == "crop":
        r = random.uniform(0.8, 0.95)                                                   # :101
        nw, nh = int(w * r), int(h * r)
        left = random.randint(0, w - nw)
        top = random.randint(0, h - nh)
        return crop_resize(img, left, top, nw, nh)
    if op == "distortion":
        noise = np.random.normal(0, 5, img.shape)                                       # :121
        return distortion(img, noise, random.uniform(0, 2))                             # :126
    raise ValueError(op)


def augment_set(img, seeds6):
    """The six augmentations of one image (BASELINE metric "transform+augment"), one task seed each."""
    return [augment_task(img, op, int(sd)) for op, sd in zip(AUGMENT_OPS, seeds6)]


def resize_normalize(img, size=224):             # sequence.py:84-88
    u8 = np.asarray(Image.fromarray(img).resize((size, size), Image.Resampling.LANCZOS))
    return u8, u8.astype(np.float32) / 255.0


# --------------------------------------------------------------------------- transform path
# The reference's OpenCV call sequence for the core transform profile (parity profile P1:
# mask_strategy hsv_h, grabcut_refine false, no upscale), on arrays.  PlantCV's fill is restated
# with scipy (skimage.remove_small_objects semantics), as in tests/golden/ref_harness.py.
def _cv2():
    import cv2
    return cv2


def _fill(bin_img, size):                              # pcv.fill, mask.py:59
    from scipy import ndimage as ndi
    b = bin_img.astype(bool)
    lab, _ = ndi.label(b)
    sizes = np.bincount(lab.ravel())
    small = sizes < size
    small[0] = False
    b[small[lab]] = False
    return b.astype(np.uint8) * 255


def _postprocess(bin_img, fill_size=1000, k=3):        # mask.py:53-69 + Transformation.py:285-299
    cv2 = _cv2()
    b = (bin_img > 0).astype(np.uint8) * 255
    filled = _fill(b, fill_size)
    el = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (k, k))
    opened = cv2.morphologyEx(cv2.morphologyEx(filled, cv2.MORPH_CLOSE, el), cv2.MORPH_OPEN, el)
    cnts, _ = cv2.findContours(opened, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    if not cnts:
        return opened, None
    cnt = max(cnts, key=cv2.contourArea)
    out = np.zeros(opened.shape, np.uint8)
    cv2.drawContours(out, [cnt], -1, color=255, thickness=-1)
    return out, cnt


def make_mask_hsv_h(rgb, green=(25, 100), fill_size=1000, brown=(0, 30, 20, 200), min_area=25):
    """make_mask with mask_strategy hsv_h (mask.py:86-91, :53-69, :395-411, :335-392)."""
    cv2 = _cv2()
    hsv = cv2.cvtColor(rgb, cv2.COLOR_RGB2HSV)
    h, s, v = cv2.split(hsv)
    raw = ((h >= green[0]) & (h <= green[1]) & (s >= 40)).astype(np.uint8) * 255
    m, cnt = _postprocess(raw, fill_size)
    if cnt is None or cv2.contourArea(cnt) <= 1:
        _, th = cv2.threshold(s, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)
        m, cnt = _postprocess(th, fill_size)
    search = cv2.dilate(m, cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (20, 20)), iterations=2) > 0
    br = ((h >= brown[0]) & (h <= brown[1]) & (s >= brown[2]) & (v <= brown[3])) & search
    k3 = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (3, 3))
    clean = cv2.morphologyEx(cv2.morphologyEx(br.astype(np.uint8) * 255, cv2.MORPH_OPEN, k3), cv2.MORPH_CLOSE, k3)
    n, labels, stats, _ = cv2.connectedComponentsWithStats(clean, connectivity=8)
    filt = np.zeros_like(clean)
    for i in range(1, n):
        if stats[i, cv2.CC_STAT_AREA] >= min_area:
            filt[labels == i] = 255
    ext = ((m > 0) | (filt > 0)).astype(np.uint8) * 255
    cnts, _ = cv2.findContours(ext, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    if cnts:
        return ext, max(cnts, key=cv2.contourArea)
    return m, None


def core_transform(rgb, sigma=1.5, roi_size=(256, 256)):
    """Core transform profile on one image with the reference's library calls:
    GaussianBlur 5x5 (blur.py:72), make_mask (mask.py:548), apply_mask white (mask_utils.py:10),
    ROI letterbox (roi.py:26-40), RGB/HSV/LAB histograms under the mask and hist.py's counters."""
    cv2 = _cv2()
    blur = cv2.GaussianBlur(rgb, (5, 5), sigma)
    mask, cnt = make_mask_hsv_h(rgb)
    masked = rgb.copy()
    masked[~(mask > 127)] = 255
    H, W = roi_size
    canvas = np.zeros((H, W, 3), np.uint8)
    bbox = None
    if cnt is not None:
        x, y, w, h = cv2.boundingRect(cnt)
        bbox = (x, y, w, h)
        scale = min(W / max(w, 1), H / max(h, 1))
        nw, nh = max(int(w * scale), 1), max(int(h * scale), 1)
        res = cv2.resize(masked[y:y + h, x:x + w], (nw, nh), interpolation=cv2.INTER_AREA)
        oy, ox = (H - nh) // 2, (W - nw) // 2
        canvas[oy:oy + nh, ox:ox + nw] = res
    hsv = cv2.cvtColor(rgb, cv2.COLOR_RGB2HSV)
    lab = cv2.cvtColor(rgb, cv2.COLOR_RGB2LAB)
    sel = mask > 0
    planes = np.concatenate([rgb, hsv, lab], axis=-1)
    hist9 = np.stack([np.bincount(planes[..., c][sel], minlength=256) for c in range(9)])
    mh = cv2.cvtColor(masked, cv2.COLOR_RGB2HSV)
    hh, ss, vv = cv2.split(mh)
    leaf = (ss > 10) & (vv > 15) & (vv < 245)                                # hist.py:188
    hsv3 = np.stack([np.bincount(mh[..., c][leaf], minlength=256) for c in range(3)])
    return blur, mask, bbox, canvas, hist9, hsv3


def _core_worker(args):
    import cv2
    cv2.setNumThreads(1)
    imgs, seeds = args
    for i, im in enumerate(imgs):
        core_transform(im)
        if seeds is not None:
            augment_set(im, seeds[:, i])
    return len(imgs)


def core_transform_pool(images, pool, nworkers, seeds=None):
    """Run core_transform (+ the six augmentations when `seeds` [6, n] is given) over `images` on `nworkers`
    processes (one task slice per worker, the reference's own parallel model: Transformation.py:691-696,
    dataset_balancer.py:137-141)."""
    n = len(images)
    per = (n + nworkers - 1) // nworkers
    chunks = [(images[i:i + per], None if seeds is None else seeds[:, i:i + per]) for i in range(0, n, per)]
    return sum(pool.map(_core_worker, chunks))
