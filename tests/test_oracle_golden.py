"""CPU tests: the oracle (oracle/spec_*.py) against the committed golden fixture
tests/golden/golden_v1.npz, which tests/golden/make_golden.py produced by running the REFERENCE's own
functions (/root/reference) on the same seeded inputs.  Bit-exact unless a tolerance is written here.
"""
import os
import random
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

import make_golden as mg  # noqa: E402  (inputs() only; never touches /root/reference at import)
from oracle import spec_augment as sa  # noqa: E402
from oracle import spec_color as sc  # noqa: E402
from oracle import spec_filters as sf  # noqa: E402
from oracle import spec_mask as sm  # noqa: E402

G = np.load(os.path.join(HERE, "golden", "golden_v1.npz"))
IMGS = mg.inputs()


def _keys(prefix):
    return sorted(k for k in G.files if k.startswith(prefix))


# ----------------------------------------------------------------------------- augment (image_augmenter.py:20-133)
def _oracle_augment(img, seed):
    """The six ops with the reference's draw order, all arithmetic from the oracle."""
    h, w = img.shape[:2]
    if seed:
        random.seed(seed)
    out = {}
    out["flip"] = sa.flip(img, sa.draw_flip())
    out["rotate"] = sa.rotate_nn(img, sa.draw_rotate())
    out["skew"] = sa.warp_bicubic(img, sa.skew_coeffs(sa.draw_skew(), w, h), True)
    k, horiz = sa.draw_shear()
    out["shear"] = sa.warp_bicubic(img, sa.shear_coeffs(k, horiz), False)
    out["crop"] = sa.crop_resize(img, *sa.draw_crop(w, h))
    noise = sa.MT19937(seed).normals(img.size, 0.0, 5.0).reshape(img.shape)
    cutoff = random.uniform(0, 2)
    out["distortion"] = sa.distortion(img, sa.noise_u8(noise), cutoff)
    return out


@pytest.mark.parametrize("seed", mg.AUG_SEEDS)
@pytest.mark.parametrize("name", ["leaf64_0", "leaf96x64", "leaf256"])
def test_augment_matches_reference(seed, name):
    got = _oracle_augment(IMGS[name], seed)
    for t in ("flip", "rotate", "skew", "shear", "crop", "distortion"):
        exp = G[f"aug/{seed}/{name}/{t}"]
        assert got[t].shape == exp.shape, (t, got[t].shape, exp.shape)
        assert np.array_equal(got[t], exp), f"{t} seed {seed} {name}: {(got[t] != exp).sum()} bytes differ"


def test_rng_trace_seed42():
    """Draw order flip -> rotate -> skew -> shear(k, choice) -> crop ratio."""
    random.seed(42)
    t = [float(not sa.draw_flip()), sa.draw_rotate(), sa.draw_skew()]
    k, horiz = sa.draw_shear()
    t += [k, float(not horiz), random.uniform(0.8, 0.95)]
    # the fixture stored choice([0, 1]) draws; choice([True, False]) consumes the stream identically
    assert [float(v) for v in t] == [float(v) for v in G["aug/trace42"][:6]]      # all six draws, exactly


# ----------------------------------------------------------------------------- make_mask (mask.py:548-582)
@pytest.mark.parametrize("strategy", mg.STRATEGIES)
def test_make_mask_matches_reference(strategy):
    cfg = sm.Cfg(mask_strategy=strategy)
    keys = _keys(f"mask/{strategy}/")
    assert keys
    for k in keys:
        name = k.split("/")[-1]
        m, info = sm.make_mask(IMGS[name], cfg)
        assert np.array_equal(m, G[k]), f"{k}: {(m != G[k]).sum()} px differ"
        bk = f"bbox/{strategy}/{name}"
        assert (info is not None) == (bk in G.files), k
        if info is not None:
            assert tuple(int(v) for v in G[bk]) == info["bbox"], k
            assert int(G[f"area2/{strategy}/{name}"][0]) == info["area2"], k


def test_kmeans_candidate_and_auto_match_reference():
    """oracle/spec_kmeans.py (cv2.kmeans + cv::RNG restated) against the reference's `_create_kmeans_mask` outputs, and the
    seven-candidate `mask_strategy: auto` against the reference's make_mask -- from the fixture, so it runs anywhere."""
    from oracle import spec_kmeans as sk
    keys = _keys("kmeans_raw/")
    assert len(keys) == 15
    for k in keys:
        _, bias, name = k.split("/")
        got = sk.kmeans_mask(IMGS[name], sm.Cfg(mask_strategy="kmeans", bg_bias=bias))
        assert np.array_equal(got, G[k]), f"{k}: {(got != G[k]).sum()} px differ"
    keys = _keys("mask/auto/")
    assert len(keys) == 5
    for k in keys:
        name = k.split("/")[-1]
        m, info = sm.make_mask_auto(IMGS[name], sm.Cfg(mask_strategy="auto"))
        assert np.array_equal(m, G[k]), f"{k}: {(m != G[k]).sum()} px differ"
        bk = f"bbox/auto/{name}"
        assert (info is not None) == (bk in G.files), k
        if info is not None:
            assert tuple(int(v) for v in G[bk]) == info["bbox"], k


@pytest.mark.parametrize("name", ["leaf64_0", "leaf96x64", "leaf256", "adv_frame"])
def test_transform_filters_match_reference(name):
    img = IMGS[name]
    cfg = sm.Cfg(mask_strategy="hsv_h")
    m, info = sm.make_mask(img, cfg)
    white = sm.apply_mask(img, m, "white")
    assert np.array_equal(white, G[f"applymask/white/{name}"])
    assert np.array_equal(sm.apply_mask(img, m, "black"), G[f"applymask/black/{name}"])
    # roi.py:20-46 canvas: interpolated -> +-1 LSB allowed by the north star, measured exact here
    if info is not None:
        canvas = sm.roi_letterbox(white, info["bbox"], (256, 256))
        assert np.array_equal(canvas, G[f"roi/{name}"])
    else:  # roi.py:23-24: no contour -> the input image is handed back unchanged
        assert np.array_equal(white, G[f"roi/{name}"])
    # blur.py:18-79: three float32 min-max normalisations -> +-1 LSB
    blur = sm.saliency_blur(white, sm.make_mask(white, cfg)[0], cfg)
    d = np.abs(blur.astype(np.int16) - G[f"blurfilter/{name}"].astype(np.int16))
    assert d.max() <= 1, f"saliency blur max diff {d.max()}"
    # brown.py:21-89 percentage and spot count
    _, spct, scount = sm.brown_spots(white, m, cfg)
    pct, count = G[f"brown/{name}"]
    assert int(count) == scount
    assert abs(pct - spct) < 1e-9
    # hist.py: leaf pixel count, 8 category percentages, 5 hue-range counts
    cnt = sm.hist_counters(white)
    assert int(G[f"hist/leafpx/{name}"][0]) == int(cnt[0])
    if cnt[0] > 0:
        assert np.allclose(cnt[1:9] / cnt[0] * 100.0, G[f"hist/cats_pct/{name}"], rtol=0, atol=1e-9)
    assert np.array_equal(cnt[9:14], G[f"hist/hue_ranges/{name}"])


# ----------------------------------------------------------------------------- library kernels
@pytest.mark.parametrize("name", ["leaf64_0", "leaf96x64"])
def test_gauss_and_canny_match_opencv(name):
    img = IMGS[name]
    assert np.array_equal(sf.gaussian_blur_u8(img, 5, 1.5), G[f"gauss5/{name}"])
    assert np.array_equal(sf.gaussian_blur_u8(img, 15, 0.0), G[f"gauss15/{name}"])
    gray = sc.rgb_to_gray(img)
    for k in _keys("canny/"):
        if not k.endswith(name):
            continue
        lo, hi, l2 = k.split("/")[1].split("_")
        assert np.array_equal(sf.canny(gray, float(lo), float(hi), bool(int(l2))), G[k]), k


def test_cvt_color_lattice():
    lattice = np.stack(np.meshgrid(*[np.arange(0, 256, 17, dtype=np.uint8)] * 3, indexing="ij"), -1).reshape(1, -1, 3)
    assert np.array_equal(sc.rgb_to_gray(lattice), G["cvt/gray"])
    assert np.array_equal(sc.rgb_to_hsv(lattice), G["cvt/hsv"])
    assert np.array_equal(sc.rgb_to_lab(lattice), G["cvt/lab"])


def test_known_answer_constants():
    """SURVEY 8c KATs: Gaussian taps, ellipse footprints, rotate sizes."""
    assert sf.gaussian_kernel_q8(15, 0.0).tolist() == [1, 3, 6, 12, 20, 30, 36, 40, 36, 30, 20, 12, 6, 3, 1]
    assert sf.gaussian_kernel_q8(5, 1.5).tolist() == [31, 60, 74, 60, 31]
    assert sf.ellipse_footprint(3).astype(int).tolist() == [[0, 1, 0], [1, 1, 1], [0, 1, 0]]
    assert sf.ellipse_footprint(9).sum(axis=1).tolist() == [1, 7, 7, 9, 9, 9, 7, 7, 1]
    assert sf.ellipse_footprint(20).sum(axis=1).tolist() == [1, 9, 13, 15, 17, 19, 19, 20, 20, 20, 20, 20, 20, 20, 19, 19, 17, 15, 13, 9]
    for angle, size in ((-30, 350), (17.1235, 322), (3.3, 272), (45, 364)):
        _, nw, nh = sa.rotate_params(angle, 256, 256)
        assert (nw, nh) == (size, size), angle


def test_exhaustive_colour_tables_and_cutoff_kats():
    """SURVEY 8c KATs: SHA-256 of the GRAY / HSV / LAB conversion of ALL 2^24 colours (digests taken from cv2.cvtColor of
    opencv-python-headless 4.13.0 in the build container; the oracle must reproduce them, and the GPU suite compares the CUDA
    kernels with the oracle on the same exhaustive lattice), and the autocontrast cut-offs of seeds 42 / 7 / 999983."""
    import hashlib
    import random
    r, g, b = np.meshgrid(*[np.arange(256, dtype=np.uint8)] * 3, indexing="ij")
    lat = np.stack([r, g, b], -1).reshape(4096, 4096, 3)
    want = {"gray": "6d4f6d7f4301c52d2672db66451b4a06a5502bef956dd81b577660f956f410ae",
            "hsv": "a5b38b214f65aed9d382cc07bf40311aff2e3e324c98da86df77bef285f224eb",
            "lab": "b3067516fa862bb008af4ffb4da5555722524a70776bf7ffd523a90f4331480c"}
    for name, fn in (("gray", sc.rgb_to_gray), ("hsv", sc.rgb_to_hsv), ("lab", sc.rgb_to_lab)):
        assert hashlib.sha256(np.ascontiguousarray(fn(lat)).tobytes()).hexdigest() == want[name], name
    for seed, cutoff in ((42, 1.2788535969157675), (7, 0.6476655296663247), (999983, 1.9837290261359528)):
        random.seed(seed)
        assert random.uniform(0, 2) == cutoff                      # image_augmenter.py:126 on a freshly seeded augmenter
        from leaffliction_b200 import augment
        ip, dp = augment.draw_params_batch(np.array([augment.TRANSFORM_CODE["distortion"]], np.int32), [seed], 256, 256)
        assert dp[0, 0] == cutoff and ip[0, 0] == int(256 * 256 * cutoff // 100)


# ----------------------------------------------------------------------------- live cross-checks (same image on the GPU box)
def test_oracle_vs_live_libraries():
    """The spec functions against Pillow / OpenCV in this interpreter (both boxes carry the same versions)."""
    cv2 = pytest.importorskip("cv2")
    from oracle import refcalls as rc
    rng = np.random.default_rng(11)
    img = rng.integers(0, 256, (48, 80, 3), dtype=np.uint8)
    assert np.array_equal(sc.rgb_to_hsv(img), cv2.cvtColor(img, cv2.COLOR_RGB2HSV))
    assert np.array_equal(sc.rgb_to_lab(img), cv2.cvtColor(img, cv2.COLOR_RGB2LAB))
    assert np.array_equal(sf.gaussian_blur_u8(img, 5, 1.5), cv2.GaussianBlur(img, (5, 5), 1.5))
    assert np.array_equal(sa.rotate_nn(img, 17.1235), rc.rotate(img, 17.1235))
    assert np.array_equal(sa.warp_bicubic(img, sa.skew_coeffs(0.1, 80, 48), True), rc.warp(img, sa.skew_coeffs(0.1, 80, 48), True))
    assert np.array_equal(sa.crop_resize(img, 3, 5, 64, 40), rc.crop_resize(img, 3, 5, 64, 40))


@pytest.mark.needs_reference
def test_golden_is_reproducible_from_reference():
    """Re-run a slice of the generator against /root/reference (build container only)."""
    import ref_harness
    ns = ref_harness.load()
    out = mg.run_augment(ns, IMGS["leaf64_0"], 42)
    for t, arr in out.items():
        assert np.array_equal(arr, G[f"aug/42/leaf64_0/{t}"]), t
    cfg = ref_harness.ref_config(ns, mask_strategy="inclusive")
    m, _ = ns.mask.make_mask(IMGS["leaf64_1"], cfg)
    assert np.array_equal(m, G["mask/inclusive/leaf64_1"])


def test_resize_specs_against_opencv():
    """spec_filters.resize_cubic_u8 is within 1 LSB of cv2.resize(INTER_CUBIC) (OpenCV's own code paths differ by that much,
    SURVEY A.12); resize_nearest_u8 is exact."""
    cv2 = pytest.importorskip("cv2")
    from oracle import spec_filters as sf
    rng = np.random.default_rng(4)
    for (h, w), (oh, ow) in (((96, 131), (125, 170)), ((64, 64), (83, 83)), ((50, 70), (100, 140))):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        ref = cv2.resize(img, (ow, oh), interpolation=cv2.INTER_CUBIC).astype(int)
        assert np.abs(sf.resize_cubic_u8(img, (ow, oh)).astype(int) - ref).max() <= 1
        m = rng.integers(0, 2, (oh, ow), dtype=np.uint8) * 255
        assert np.array_equal(sf.resize_nearest_u8(m, (w, h)), cv2.resize(m, (w, h), interpolation=cv2.INTER_NEAREST))


def test_resize_area_down_against_opencv():
    """The INTER_AREA shrink restatement (ROI letterbox of a bounding box larger than roi_size, roi.py:35-38) equals
    cv2.resize bit for bit: integer factors (2x2, 3x3, 4x4, mixed), general ratios, one axis unchanged."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(6)
    for (w, h), (nw, nh) in (((300, 400), (256, 192)), ((512, 400), (256, 200)), ((700, 650), (256, 237)), ((384, 384), (128, 128)),
                             ((257, 300), (219, 256)), ((512, 256), (128, 128)), ((301, 299), (256, 254)), ((256, 300), (256, 100)),
                             ((97, 61), (33, 21))):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        assert np.array_equal(sm.resize_area_down(img, nw, nh), cv2.resize(img, (nw, nh), interpolation=cv2.INTER_AREA)), ((w, h), (nw, nh))
    big = rng.integers(0, 256, (300, 340, 3), dtype=np.uint8)      # through the letterbox itself
    x, y, w, h = 10, 20, 320, 270
    sc_ = min(128 / w, 128 / h)
    nw, nh = max(int(w * sc_), 1), max(int(h * sc_), 1)
    exp = np.zeros((128, 128, 3), np.uint8)
    exp[(128 - nh) // 2:(128 - nh) // 2 + nh, (128 - nw) // 2:(128 - nw) // 2 + nw] = cv2.resize(big[y:y + h, x:x + w], (nw, nh), interpolation=cv2.INTER_AREA)
    assert np.array_equal(sm.roi_letterbox(big, (x, y, w, h), (128, 128)), exp)
