"""Drop-in for srcs/preprocessing/image_augmenter.py: the six class-balancing augmentations.

`ImageAugmenter` keeps the reference's constructor and method signatures
(`ImageAugmenter(seed=None)`, `.flip/.rotate/.skew/.shear/.crop/.distortion(image_path,
output_path) -> bool`, image_augmenter.py:15-133) and its error convention (catch everything, log,
return False).  Every parameter is drawn on the host with the SAME `random` / `np.random` calls in
the SAME order as the reference, so a given seed yields the reference's parameters by
construction; the pixel work runs in libleafx's CUDA kernels.  JPEG decode/encode stays on the
host (Pillow, quality 95 -- image_utils.py:19-59; out of scope per SURVEY.md section 8).

`augment_arrays` is the batched array-in/array-out form used by the dataset balancer and benches.
"""
from __future__ import annotations

import os

import logging
import math
import random
from pathlib import Path
from typing import List, Sequence

import numpy as np

logger = logging.getLogger(__name__)

TRANSFORMATIONS = ["flip", "rotate", "skew", "shear", "crop", "distortion"]
SUPPORTED_EXTENSIONS = {".jpg"}
NOISE_LEVEL = 5


# --------------------------------------------------------------------------- host parameter maths
def rotate_matrix(angle: float, w: int, h: int):
    """PIL Image.rotate(angle, expand=True) geometry -> (inverse affine matrix[6], nw, nh),
    or a transpose tag for multiples of 90 degrees (PIL/Image.py rotate)."""
    angle = angle % 360.0
    if angle == 0:
        return "copy", w, h
    if angle == 180:
        return "rot180", w, h
    if angle == 90:
        return "rot90", h, w
    if angle == 270:
        return "rot270", h, w
    cx, cy = w / 2.0, h / 2.0
    a = -math.radians(angle)
    m = [round(math.cos(a), 15), round(math.sin(a), 15), 0.0, round(-math.sin(a), 15), round(math.cos(a), 15), 0.0]

    def tf(x, y):
        return m[0] * x + m[1] * y + m[2], m[3] * x + m[4] * y + m[5]
    m[2], m[5] = tf(-cx, -cy)
    m[2] += cx
    m[5] += cy
    xs, ys = zip(*(tf(x, y) for x, y in ((0, 0), (w, 0), (w, h), (0, h))))
    nw = math.ceil(max(xs)) - math.floor(min(xs))
    nh = math.ceil(max(ys)) - math.floor(min(ys))
    m[2], m[5] = tf(-(nw - w) / 2.0, -(nh - h) / 2.0)
    return m, nw, nh


def fixed_affine(m):
    """libImaging affine_fixed 16.16 coefficients (half-pixel centre folded into a2, a5)."""
    fix = lambda v: int(math.floor(v * 65536.0 + 0.5))  # noqa: E731
    return [fix(m[0]), fix(m[1]), fix(m[2] + m[0] * 0.5 + m[1] * 0.5),
            fix(m[3]), fix(m[4]), fix(m[5] + m[3] * 0.5 + m[4] * 0.5)]


# --------------------------------------------------------------------------- GPU execution
def _ops():
    from . import ops  # imports torch; raises loudly without CUDA when an op is called
    return ops


def _to_dev(arrs: Sequence[np.ndarray]):
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("leaffliction_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.from_numpy(np.ascontiguousarray(np.stack(arrs))).cuda()


def gpu_flip(imgs, left_right):
    return _ops().flip(_to_dev(imgs), list(left_right)).cpu().numpy()


def gpu_rotate(imgs, angles) -> List[np.ndarray]:
    ops = _ops()
    h, w = imgs[0].shape[:2]
    outs: List = [None] * len(imgs)
    idx, params = [], []
    for i, ang in enumerate(angles):
        m, nw, nh = rotate_matrix(ang, w, h)
        if isinstance(m, str):  # PIL's transpose fast paths: pure index permutations
            im = imgs[i]
            outs[i] = {"copy": im.copy(), "rot180": im[::-1, ::-1].copy(),
                       "rot90": np.transpose(im, (1, 0, 2))[::-1].copy(),
                       "rot270": np.transpose(im, (1, 0, 2))[:, ::-1].copy()}[m]
        else:
            idx.append(i)
            params.append(fixed_affine(m) + [nw, nh])
    if idx:
        slab, _ = ops.rotate_nn(_to_dev([imgs[i] for i in idx]), np.array(params, np.int32), 255)
        slab = slab.cpu().numpy()
        for k, i in enumerate(idx):
            nw, nh = params[k][6], params[k][7]
            outs[i] = slab[k, : nh * nw * 3].reshape(nh, nw, 3).copy()
    return outs


def gpu_warp(imgs, coeffs, perspective):
    return _ops().warp_bicubic(_to_dev(imgs), np.asarray(coeffs, np.float64), list(perspective)).cpu().numpy()


def gpu_crop(imgs, boxes):
    h, w = imgs[0].shape[:2]
    return _ops().crop_lanczos(_to_dev(imgs), np.asarray(boxes, np.int32), (h, w)).cpu().numpy()


def gpu_distort(imgs, noises_u8, cutoffs):
    h, w = imgs[0].shape[:2]
    cuts = [int(h * w * c // 100) for c in cutoffs]      # ImageOps.autocontrast: int(n * cutoff // 100)
    return _ops().distort(_to_dev(imgs), _to_dev(noises_u8), cuts).cpu().numpy()


# --------------------------------------------------------------------------- file-level drop-in
def _load_rgb(image_path) -> np.ndarray:
    """ImageLoader.load_pil_image semantics (image_utils.py:19-39): .jpg only, RGB."""
    from PIL import Image
    p = Path(image_path)
    if not p.exists():
        raise FileNotFoundError(f"Image not found: {p}")
    if p.suffix.lower() not in SUPPORTED_EXTENSIONS:
        raise ValueError(f"Unsupported image format: {p.suffix}")
    with Image.open(p) as im:
        if im.mode != "RGB":
            im = im.convert("RGB")
        return np.array(im)


def _save_rgb(arr: np.ndarray, output_path, quality: int = 95) -> None:
    from PIL import Image
    out = Path(output_path)
    out.parent.mkdir(parents=True, exist_ok=True)
    Image.fromarray(arr).save(out, quality=quality)


class ImageAugmenter:
    NOISE_LEVEL = NOISE_LEVEL

    def __init__(self, seed=None):
        if seed:  # reference quirk: seed 0 / None leaves both RNGs unseeded (image_augmenter.py:16-18)
            random.seed(seed)
            np.random.seed(seed)

    def _run(self, image_path, output_path, fn) -> bool:
        try:
            _save_rgb(fn(_load_rgb(image_path)), output_path)
            return True
        except Exception as e:  # same convention as the reference: log, return False, never raise
            logger.error(f"Failed to process {image_path} - {e}")
            return False

    def flip(self, image_path, output_path):
        return self._run(image_path, output_path, lambda im: gpu_flip([im], [random.choice([True, False])])[0])

    def rotate(self, image_path, output_path):
        return self._run(image_path, output_path, lambda im: gpu_rotate([im], [random.uniform(-30, 30)])[0])

    def skew(self, image_path, output_path):
        def fn(im):
            h, w = im.shape[:2]
            s = random.uniform(0.05, 0.15)
            return gpu_warp([im], [[1 + s, 0, -s * w, 0, 1 + s, -s * h, 0, 0]], [True])[0]
        return self._run(image_path, output_path, fn)

    def shear(self, image_path, output_path):
        def fn(im):
            k = random.uniform(-0.2, 0.2)
            c = [1, k, 0, 0, 1, 0, 0, 0] if random.choice([True, False]) else [1, 0, 0, k, 1, 0, 0, 0]
            return gpu_warp([im], [c], [False])[0]
        return self._run(image_path, output_path, fn)

    def crop(self, image_path, output_path):
        def fn(im):
            h, w = im.shape[:2]
            r = random.uniform(0.8, 0.95)
            nw, nh = int(w * r), int(h * r)
            left = random.randint(0, w - nw)
            top = random.randint(0, h - nh)
            return gpu_crop([im], [(left, top, nw, nh)])[0]
        return self._run(image_path, output_path, fn)

    def distortion(self, image_path, output_path):
        def fn(im):
            noise = np.random.normal(0, self.NOISE_LEVEL, im.shape).astype(np.uint8)
            return gpu_distort([im], [noise], [random.uniform(0, 2)])[0]
        return self._run(image_path, output_path, fn)


def augment_arrays(images: Sequence[np.ndarray], transforms: Sequence[str], seeds: Sequence[int]) -> List[np.ndarray]:
    """Batched form of `_process_single_transformation` (dataset_balancer.py:201-207): task i applies
    transforms[i] to images[i] with a fresh ImageAugmenter(seed=seeds[i]).  Parameters are drawn per
    task exactly as the reference would, then tasks are grouped per transform and executed as one
    GPU batch each."""
    n = len(images)
    outs: List = [None] * n
    groups = {t: [] for t in TRANSFORMATIONS}
    params = [None] * n
    for i, (t, sd) in enumerate(zip(transforms, seeds)):
        if sd:
            random.seed(sd)
            np.random.seed(sd)
        h, w = images[i].shape[:2]
        if t == "flip":
            params[i] = random.choice([True, False])
        elif t == "rotate":
            params[i] = random.uniform(-30, 30)
        elif t == "skew":
            s = random.uniform(0.05, 0.15)
            params[i] = ([1 + s, 0, -s * w, 0, 1 + s, -s * h, 0, 0], True)
        elif t == "shear":
            k = random.uniform(-0.2, 0.2)
            params[i] = (([1, k, 0, 0, 1, 0, 0, 0] if random.choice([True, False]) else [1, 0, 0, k, 1, 0, 0, 0]), False)
        elif t == "crop":
            r = random.uniform(0.8, 0.95)
            nw, nh = int(w * r), int(h * r)
            left = random.randint(0, w - nw)
            params[i] = (left, random.randint(0, h - nh), nw, nh)
        elif t == "distortion":
            noise = np.random.normal(0, NOISE_LEVEL, images[i].shape).astype(np.uint8)
            params[i] = (noise, random.uniform(0, 2))
        else:
            raise ValueError(f"unknown transform {t!r}")
        groups[t].append(i)

    def by_shape(ids):
        d = {}
        for i in ids:
            d.setdefault(images[i].shape, []).append(i)
        return d.values()
    for ids in by_shape(groups["flip"]):
        for i, o in zip(ids, gpu_flip([images[i] for i in ids], [params[i] for i in ids])):
            outs[i] = o
    for ids in by_shape(groups["rotate"]):
        for i, o in zip(ids, gpu_rotate([images[i] for i in ids], [params[i] for i in ids])):
            outs[i] = o
    for ids in by_shape(groups["skew"] + groups["shear"]):
        res = gpu_warp([images[i] for i in ids], [params[i][0] for i in ids], [params[i][1] for i in ids])
        for i, o in zip(ids, res):
            outs[i] = o
    for ids in by_shape(groups["crop"]):
        for i, o in zip(ids, gpu_crop([images[i] for i in ids], [params[i] for i in ids])):
            outs[i] = o
    for ids in by_shape(groups["distortion"]):
        res = gpu_distort([images[i] for i in ids], [params[i][0] for i in ids], [params[i][1] for i in ids])
        for i, o in zip(ids, res):
            outs[i] = o
    return outs


# --------------------------------------------------------------------------- device-resident batches (BASELINE config 3)
def draw_task_params(transform: str, seed: int, h: int, w: int, want_noise: bool = True):
    """The parameters `_process_single_transformation` (dataset_balancer.py:201-207) would draw for one task:
    a fresh ImageAugmenter(seed) seeds `random` / `np.random` (only `if seed:`), then the method draws in the
    reference's order.  Returns a tuple whose layout depends on the transform."""
    if seed:
        random.seed(seed)
        np.random.seed(seed)
    if transform == "flip":
        return (random.choice([True, False]),)
    if transform == "rotate":
        return (random.uniform(-30, 30),)
    if transform == "skew":
        s = random.uniform(0.05, 0.15)
        return ([1 + s, 0, -s * w, 0, 1 + s, -s * h, 0, 0], True)
    if transform == "shear":
        k = random.uniform(-0.2, 0.2)
        return (([1, k, 0, 0, 1, 0, 0, 0] if random.choice([True, False]) else [1, 0, 0, k, 1, 0, 0, 0]), False)
    if transform == "crop":
        r = random.uniform(0.8, 0.95)
        nw, nh = int(w * r), int(h * r)
        left = random.randint(0, w - nw)
        return (left, random.randint(0, h - nh), nw, nh)
    if transform == "distortion":
        noise = np.random.normal(0, NOISE_LEVEL, (h, w, 3)).astype(np.uint8) if want_noise else None
        return (noise, random.uniform(0, 2))
    raise ValueError(f"unknown transform {transform!r}")


TRANSFORM_CODE = {name: i for i, name in enumerate(TRANSFORMATIONS)}     # == LFX_AUG_* of include/leafx.h


def draw_params_batch(transforms, seeds, h: int, w: int, threads: int = 0):
    """Parameters of a batch of balancing tasks, drawn natively (lfx_draw_augment_params: CPython's MT19937
    `random` stream of each task seed, consumed in the reference's order; no GPU involved).
    transforms: int codes (TRANSFORM_CODE) or names; seeds: ints.  Returns (iparams int32 [B,8], dparams float64 [B,8])
    laid out as include/leafx.h documents.  Tasks with seed 0 (unseeded in the reference, image_augmenter.py:16)
    are drawn from the interpreter's current global stream by draw_task_params."""
    import ctypes as C
    from . import _lib
    tr = np.ascontiguousarray([TRANSFORM_CODE[t] if isinstance(t, str) else int(t) for t in transforms]
                              if not isinstance(transforms, np.ndarray) else transforms, dtype=np.int32)
    sd = np.ascontiguousarray(seeds, dtype=np.int64)
    if sd.size and (sd.min() < 0 or sd.max() > 0xFFFFFFFF):
        raise ValueError("task seeds must be in [0, 2^32)")
    sd32 = sd.astype(np.uint32)
    B = len(tr)
    ip, dp = np.zeros((B, 8), np.int32), np.zeros((B, 8), np.float64)
    P = C.c_void_p
    _lib.check(_lib.load().lfx_draw_augment_params(tr.ctypes.data_as(P), sd32.ctypes.data_as(P), B, int(h), int(w),
                                                   ip.ctypes.data_as(P), dp.ctypes.data_as(P), int(threads)))
    for i in np.nonzero(sd == 0)[0]:
        name = TRANSFORMATIONS[tr[i]]
        p = draw_task_params(name, 0, h, w, want_noise=False)
        ip[i], dp[i] = 0, 0.0
        if name == "flip":
            ip[i, 0] = 0 if p[0] else 1
        elif name == "rotate":
            m, nw, nh = rotate_matrix(p[0], w, h)
            if isinstance(m, str):
                m, nw, nh = [1.0, 0.0, 0.0, 0.0, 1.0, 0.0], w, h
            ip[i], dp[i, 0] = fixed_affine(m) + [nw, nh], p[0]
        elif name in ("skew", "shear"):
            dp[i], ip[i, 0] = p[0], int(p[1])
        elif name == "crop":
            ip[i, :4] = p
        else:
            ip[i, 0], dp[i, 0] = int(h * w * p[1] // 100), p[1]
    return ip, dp


def draw_params_from_words(transforms: np.ndarray, seeds32: np.ndarray, words: np.ndarray, h: int, w: int):
    """draw_params_batch with the MT19937 seeding already done on the device: words uint32 [B, n] = the first outputs of every
    task's `random` stream (ops.seed_words).  Same results (tests/test_params_cpu.py); seeds must be non-zero."""
    import ctypes as C
    from . import _lib
    tr = np.ascontiguousarray(transforms, dtype=np.int32)
    sd = np.ascontiguousarray(seeds32, dtype=np.uint32)
    wd = np.ascontiguousarray(words).view(np.uint32)
    B = len(tr)
    if wd.shape[0] != B or len(sd) != B:
        raise ValueError("draw_params_from_words: one row of words and one seed per task")
    ip, dp = np.zeros((B, 8), np.int32), np.zeros((B, 8), np.float64)
    P = C.c_void_p
    _lib.check(_lib.load().lfx_draw_augment_params_words(tr.ctypes.data_as(P), sd.ctypes.data_as(P), wd.ctypes.data_as(P), int(wd.shape[1]),
                                                         B, int(h), int(w), ip.ctypes.data_as(P), dp.ctypes.data_as(P)))
    return ip, dp


class TaskArrays:
    """Struct-of-arrays view of a task list: transform codes, seeds, dataset index of the source image."""

    def __init__(self, tasks):
        self.transform = np.fromiter((TRANSFORM_CODE[t.transform_name] for t in tasks), np.int32, len(tasks))
        self.seed = np.fromiter((t.seed for t in tasks), np.int64, len(tasks))
        self.source_index = np.fromiter((t.source_index for t in tasks), np.int64, len(tasks))

    def __len__(self):
        return len(self.transform)

    def slice(self, lo, hi, step=1):
        o = object.__new__(TaskArrays)
        o.transform, o.seed, o.source_index = self.transform[lo:hi:step], self.seed[lo:hi:step], self.source_index[lo:hi:step]
        return o

    def shard(self, rank: int, world: int):
        """Tasks i with i % world == rank."""
        return self.slice(rank, None, world)


def augment_device(x, tasks, device_noise: bool = True):
    """Run balance tasks on images already resident in HBM.  `x`: uint8 [N,H,W,3] CUDA tensor; `tasks`: a TaskArrays or
    a list of objects with .transform_name, .seed, .source_index.  Returns {key: (task indices, output tensor)} with
    key in flip / rotate / warp (skew + shear) / crop / distortion; rotate returns (indices, slab, [(nh, nw)]).
    Parameters come from one native call (draw_params_batch); with `device_noise` the distortion noise (NumPy's
    legacy MT19937 normal stream of the task seed) is generated on the GPU (ops.legacy_normal_noise) instead of by
    np.random on the host."""
    import torch
    ops = _ops()
    h, w = int(x.shape[1]), int(x.shape[2])
    ta = tasks if isinstance(tasks, TaskArrays) else TaskArrays(tasks)
    dev = x.device

    def ids_of(*names):
        return np.nonzero(np.isin(ta.transform, [TRANSFORM_CODE[n] for n in names]))[0]

    up = ops.PackedUpload(dev)
    # the noise streams need only the seeds: launched first, they run on the GPU while the host draws the parameters
    dist_ids = ids_of("distortion")
    noise = None
    if len(dist_ids) and device_noise:
        sd = ta.seed[dist_ids]
        dseeds = up.upload({"seeds": (sd & 0xFFFFFFFF).astype(np.uint32).view(np.int32)})["seeds"]
        noise = ops.legacy_normal_noise(sd, h * w * 3, NOISE_LEVEL, dev, dseeds=dseeds).view(len(dist_ids), h, w, 3)
    few_cores = len(os.sched_getaffinity(0)) < 8     # with >= 8 cores the native host drawer is quicker than a device round trip
    if device_noise and few_cores and len(ta) >= 512 and (ta.seed != 0).all() and int(ta.seed.max()) <= 0xFFFFFFFF:
        # the seedings of all tasks as one small kernel (lfx_seed_words); the host consumes the first words of each stream
        # (a rank of an 8-GPU box owns 4 cores: 4608 seedings there cost as much as the rank's augment kernels)
        sd32 = (ta.seed & 0xFFFFFFFF).astype(np.uint32)
        d_all = up.upload({"all_seeds": sd32.view(np.int32)})["all_seeds"]
        words = ops.seed_words(d_all, AugmentSet.SEED_WORDS).cpu().numpy()
        ip, dp = draw_params_from_words(ta.transform, sd32, words, h, w)
    else:
        ip, dp = draw_params_batch(ta.transform, ta.seed, h, w)
    # every per-task parameter array of every op goes to the device in ONE pinned, non-blocking transfer; the kernel
    # launches that follow never wait for the stream to drain
    groups = {"flip": ids_of("flip"), "rotate": ids_of("rotate"), "warp": ids_of("skew", "shear"), "crop": ids_of("crop"),
              "distortion": dist_ids}
    host = {f"src_{k}": ta.source_index[ids].astype(np.int32) for k, ids in groups.items() if len(ids)}
    plan = None
    if len(groups["flip"]):
        host["flip_mode"] = ip[groups["flip"], 0]
    if len(groups["rotate"]):
        host["rot"] = ip[groups["rotate"]]
    if len(groups["warp"]):
        host["warp_coef"], host["warp_persp"] = dp[groups["warp"]], ip[groups["warp"], 0]
    if len(groups["crop"]):
        plan = ops.CropPlan(ip[groups["crop"], :4], (h, w), dev, upload=False)
        host["crop_box"], host["crop_off"] = plan.h_box, plan.h_off
    if len(groups["distortion"]):
        host["cuts"] = ip[groups["distortion"], 0]
    d = up.upload(host) if host else {}
    out = {}

    # the kernels read source image src_index[i] in place (no gather copy of the sources)
    ids = groups["flip"]
    if len(ids):
        out["flip"] = (ids, ops.flip(x, d["flip_mode"], src_index=d["src_flip"]))
    ids = groups["rotate"]
    if len(ids):
        slab, stride = ops.rotate_nn(x, ip[ids], 255, dparams=d["rot"], src_index=d["src_rotate"])
        out["rotate"] = (ids, slab, ip[ids][:, [7, 6]])
    ids = groups["warp"]
    if len(ids):
        out["warp"] = (ids, ops.warp_bicubic(x, d["warp_coef"], d["warp_persp"], src_index=d["src_warp"]))
    ids = groups["crop"]
    if len(ids):
        plan.box, plan.off = d["crop_box"], d["crop_off"]
        out["crop"] = (ids, ops.crop_lanczos(x, plan, src_index=d["src_crop"]))
    ids = dist_ids
    if len(ids):
        if noise is None:
            noises = []
            for sd in ta.seed[ids]:
                if sd:
                    np.random.seed(int(sd))
                noises.append(np.random.normal(0, NOISE_LEVEL, (h, w, 3)).astype(np.uint8))
            noise = torch.from_numpy(np.stack(noises)).to(dev)
        out["distortion"] = (ids, ops.distort(x, noise, d["cuts"], src_index=d["src_distortion"]))
    return out


# --------------------------------------------------------------------------- the 6-op augment set on a resident batch
class AugmentSet:
    """Every image of a device-resident batch through all six ImageAugmenter ops (BASELINE metric "transform+augment":
    6 augmentations per image, image_augmenter.py:20-133), each with its own task seed exactly as
    `_process_single_transformation` runs it (dataset_balancer.py:201-207: a fresh ImageAugmenter(seed) per task).
    Outputs, noise and scratch are allocated once; `run` draws the parameters natively (lfx_draw_augment_params), uploads
    them in one pinned transfer and launches the kernels: noise, flip, rotate, skew, shear, crop, distort.
    Seeds must be non-zero (seed 0 leaves the reference unseeded: such tasks go through `augment_arrays`)."""

    OPS = ("flip", "rotate", "skew", "shear", "crop", "distortion")

    SEED_WORDS = 16    # stream outputs fetched per task; draws that need more (~1 in 10^4 tasks) are seeded on the host

    def __init__(self, B: int, H: int, W: int, device, concurrent: bool = False, device_seeding: bool = True, pipelined: bool = False):
        """`device_seeding`: the 6 B `random.seed(task seed)` calls of a step run as one small kernel (lfx_seed_words) on a
        side stream; the host only consumes the first words of each stream (a host core spends ~1.8 us per
        seeding, which outlasts the GPU step once a rank owns only a few cores).
        `concurrent`: launch the noise generator and the five geometric kernels on two side streams, so that they share
        the SMs with each other and with whatever the caller queues on its own stream between start() and finish()
        (every kernel of this path is issue- or latency-bound, none fills the machine alone)."""
        import torch
        self.B, self.H, self.W, self.device = int(B), int(H), int(W), device
        self.concurrent = bool(concurrent)
        # pipelined (with `concurrent`): the distortion runs on the noise stream, right behind its noise, and finish() does NOT
        # join the side streams into the caller's stream -- consecutive steps overlap (the memory-bound distortion of step i runs
        # under the issue-bound geometric kernels and k_core of step i + 1).  Each side stream is in order with itself, so the
        # set's own buffers are safe; the CALLER must call join() before it reads any output or overwrites `x`.
        self.pipelined = bool(pipelined) and self.concurrent
        self._done = None
        self.device_seeding = bool(device_seeding) and os.environ.get("LFX_DEVICE_SEEDING", "1") != "0"
        if not self.concurrent and len(os.sched_getaffinity(0)) >= 8:
            # in stream order (see start()) the host would wait for the stream to drain before every chunk: with eight or more
            # cores the native host drawer (lfx_draw_augment_params) is quicker than that wait
            self.device_seeding = False
        self._seed_bufs = None
        self._streams = None
        self._pending = None
        u8 = dict(dtype=torch.uint8, device=device)
        self.flip = torch.empty((B, H, W, 3), **u8)
        self.skew = torch.empty((B, H, W, 3), **u8)
        self.shear = torch.empty((B, H, W, 3), **u8)
        self.crop = torch.empty((B, H, W, 3), **u8)
        self.distortion = torch.empty((B, H, W, 3), **u8)
        self.noise = torch.empty((B, H * W * 3), **u8)
        self.hist_ws = torch.empty((B, 3, 256), dtype=torch.int32, device=device)
        # rotate outputs: pitched slab, image i = rotate[i, :nh_i*nw_i*3].  nw*nh = WH + (W^2+H^2)/2 * sin(2a) grows with
        # |a| up to 45 degrees, so the reference's +-30 degree draw (image_augmenter.py:36) is bounded by the 30-degree size
        _, nw30, nh30 = rotate_matrix(30.0, W, H)
        self.rotate_stride = (((nw30 + 1) * (nh30 + 1) * 3 + 15) // 16) * 16
        self.rotate = torch.empty((B, self.rotate_stride), **u8)
        self.rotate_hw = None              # int32 [B,2] (nh, nw) on the host
        self._up = None
        # every Lanczos table a crop of this image size can need (nw = int(W * r), r in [0.8, 0.95], image_augmenter.py:100-103),
        # so that the device copy of the tables is uploaded once and never replaced under queued kernels
        _ops()._lanczos.prefill(int(W * 0.8), int(W * 0.95) + 1, W)
        _ops()._lanczos.prefill(int(H * 0.8), int(H * 0.95) + 1, H)

    def out_bytes_per_image(self, rotate_px_mean: float) -> float:
        n = self.H * self.W * 3
        return 5 * n + 3 * rotate_px_mean

    def run(self, x, seeds: np.ndarray, timings: dict = None):
        """x: uint8 [B,H,W,3] on the device; seeds: int [6,B] (one task seed per op and image, non-zero).
        `timings`: optional dict filled with per-kernel milliseconds (synchronising CUDA events: profiling runs only)."""
        self.start(x, seeds, timings)
        return self.finish(timings)

    def start(self, x, seeds: np.ndarray, timings: dict = None):
        """Queue noise, flip, rotate, skew, shear and crop (on the side streams when `concurrent`); finish() queues the
        distortion, which needs the noise, on the caller's stream and joins the side streams."""
        import torch
        ops = _ops()
        B, H, W = self.B, self.H, self.W
        seeds = np.asarray(seeds, np.int64).reshape(6, B)
        if (seeds == 0).any():
            raise ValueError("AugmentSet: task seeds must be non-zero (seed 0 is 'unseeded' in the reference)")
        if self._up is None:
            self._up = ops.PackedUpload(self.device)
        tr = np.repeat(np.arange(6, dtype=np.int32), B)
        cur = torch.cuda.current_stream(self.device)
        side = self.concurrent and timings is None
        if side:
            if self._streams is None:
                self._streams = (torch.cuda.Stream(self.device), torch.cuda.Stream(self.device))
            ev0 = torch.cuda.Event()
            ev0.record(cur)            # inputs ready, previous readers of our output buffers done
            for st in self._streams:
                st.wait_event(ev0)
        s_noise, s_geo = self._streams if side else (cur, cur)
        if self.device_seeding:
            # the step's 6 B seedings: a 25-us kernel on a stream of its own (it must not queue behind the previous step's
            # kernels), words back to pinned memory while the noise kernel is being launched
            if self._seed_bufs is None:
                nw = self.SEED_WORDS
                # A stream of its own only in `concurrent` mode (the bench step).  Inside TransformEngine.run_host (chunks on one
                # kernel stream, copy streams on both sides) a separate seeding stream ended in an illegal-address fault on
                # hosts with few cores -- at normal and at high stream priority, never with the kernel on the caller's stream --
                # so the pipelined host path keeps it in stream order; that path is bound by its device-to-host copies anyway.
                self._seed_bufs = (torch.cuda.Stream(self.device) if self.concurrent else None,
                                   torch.empty(6 * B, dtype=torch.int32).pin_memory(),
                                   torch.empty(6 * B, dtype=torch.int32, device=self.device),
                                   torch.empty((6 * B, nw), dtype=torch.int32, device=self.device),
                                   torch.empty((6 * B, nw), dtype=torch.int32).pin_memory())
            s_seed, h_seeds, d_seeds, d_words, h_words = self._seed_bufs
            if s_seed is None or timings is not None:
                s_seed = cur
            sd32 = (seeds.reshape(-1) & 0xFFFFFFFF).astype(np.uint32)
            h_seeds.numpy()[:] = sd32.view(np.int32)
            with torch.cuda.stream(s_seed):
                d_seeds.copy_(h_seeds, non_blocking=True)
                ops.seed_words(d_seeds, self.SEED_WORDS, out=d_words)
                h_words.copy_(d_words, non_blocking=True)
                ev_words = torch.cuda.Event()
                ev_words.record(s_seed)
        # noise first: it needs only the seeds and runs while the host draws the other parameters
        with torch.cuda.stream(s_noise):
            d0 = self._up.upload({"seeds": (seeds[5] & 0xFFFFFFFF).astype(np.uint32).view(np.int32)})

        def timed(name, fn):
            if timings is None:
                return fn()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            r = fn()
            b.record()
            b.synchronize()
            timings[name] = timings.get(name, 0.0) + a.elapsed_time(b)
            return r

        with torch.cuda.stream(s_noise):
            timed("k_legacy_normal_u8", lambda: ops.legacy_normal_noise(seeds[5], H * W * 3, NOISE_LEVEL, self.device,
                                                                        dseeds=d0["seeds"], out=self.noise))
            ev_noise = torch.cuda.Event()
            ev_noise.record(s_noise)
        if self.device_seeding:
            ev_words.synchronize()
            ip, dp = draw_params_from_words(tr, sd32, h_words.numpy(), H, W)
        else:
            ip, dp = draw_params_batch(tr, seeds.reshape(-1), H, W)
        ip, dp = ip.reshape(6, B, 8), dp.reshape(6, B, 8)
        plan = ops.CropPlan(ip[4, :, :4], (H, W), self.device, upload=False)
        self.rotate_hw = ip[1][:, [7, 6]]
        self.crop_px = int((ip[4, :, 2].astype(np.int64) * ip[4, :, 3]).sum())
        if side and self.pipelined:
            with torch.cuda.stream(s_noise):
                # the cut-offs go up on the noise stream (their buffer is then allocated, used and recycled on ONE stream)
                dc = self._up.upload({"cuts": ip[5, :, 0]})
                ops.distort(x, self.noise.view(B, H, W, 3), dc["cuts"], out=self.distortion, hist_ws=self.hist_ws)
                ev_noise = torch.cuda.Event()
                ev_noise.record(s_noise)
        with torch.cuda.stream(s_geo):
            d = self._up.upload({"flip_mode": ip[0, :, 0], "rot": ip[1], "skew_coef": dp[2], "skew_persp": ip[2, :, 0],
                                 "shear_coef": dp[3], "shear_persp": ip[3, :, 0], "crop_box": plan.h_box, "crop_off": plan.h_off,
                                 "cuts": ip[5, :, 0]})
            plan.box, plan.off = d["crop_box"], d["crop_off"]
            timed("k_flip_vec", lambda: ops.flip(x, d["flip_mode"], out=self.flip))
            timed("k_rotate_nn", lambda: ops.rotate_nn(x, ip[1], 255, dparams=d["rot"], out=self.rotate))
            timed("k_warp_bicubic(skew)", lambda: ops.warp_bicubic(x, d["skew_coef"], d["skew_persp"], out=self.skew))
            timed("k_warp_bicubic(shear)", lambda: ops.warp_bicubic(x, d["shear_coef"], d["shear_persp"], out=self.shear))
            timed("k_lanczos_dp4a", lambda: ops.crop_lanczos(x, plan, out=self.crop))
            ev_geo = torch.cuda.Event()
            ev_geo.record(s_geo)
        self._pending = (x, d, ev_noise, ev_geo, side)
        if side and self.pipelined:
            self._done = (ev_noise, ev_geo)
        return self

    def join(self):
        """pipelined mode: make the caller's stream wait for everything queued so far (no-op otherwise)."""
        import torch
        if self._done is not None:
            cur = torch.cuda.current_stream(self.device)
            for ev in self._done:
                cur.wait_event(ev)
            self._done = None
        return self

    def finish(self, timings: dict = None):
        import torch
        ops = _ops()
        if self._pending is None:
            raise RuntimeError("AugmentSet.finish() without start()")
        x, d, ev_noise, ev_geo, side = self._pending
        self._pending = None
        if side and self.pipelined:
            return self                    # the distortion is already queued behind its noise; join() orders the caller's stream
        B, H, W = self.B, self.H, self.W
        cur = torch.cuda.current_stream(self.device)
        if side:
            cur.wait_event(ev_noise)
            cur.wait_event(ev_geo)     # (also orders the upload of `cuts`, made on the geometry stream)

        def timed(name, fn):
            if timings is None:
                return fn()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            r = fn()
            b.record()
            b.synchronize()
            timings[name] = timings.get(name, 0.0) + a.elapsed_time(b)
            return r
        timed("k_distort_hist+lut+apply", lambda: ops.distort(x, self.noise.view(B, H, W, 3), d["cuts"], out=self.distortion,
                                                               hist_ws=self.hist_ws))
        return self

    def algo_bytes(self):
        """Algorithmic HBM bytes per kernel for the last run, whole batch."""
        n = self.H * self.W * 3
        B = self.B
        rot_px = int((self.rotate_hw[:, 0].astype(np.int64) * self.rotate_hw[:, 1]).sum())
        return {"k_flip_vec": 2 * n * B, "k_rotate_nn": n * B + 3 * rot_px, "k_warp_bicubic(skew)": 2 * n * B,
                "k_warp_bicubic(shear)": 2 * n * B, "k_lanczos_dp4a": 3 * self.crop_px + n * B, "k_distort_hist+lut+apply": 3 * n * B,
                "k_legacy_normal_u8": n * B}

    launches_per_run = 10  # seed words, noise, flip, rotate, skew, shear, crop, distort hist / lut / apply (+ one memset node)
