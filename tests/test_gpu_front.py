"""GPU parity of the per-image front ends: Canny, inclusive/enhanced strategies, brown spots,
saliency 'Blur', analyze record, histogram statistics."""
import numpy as np
import pytest
import torch

from leaffliction_b200 import filters, ops, synth, transform
from oracle import spec_color as sc
from oracle import spec_contour as spc
from oracle import spec_filters as sf
from oracle import spec_mask as sm

pytestmark = pytest.mark.gpu


def up(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def _images():
    ims = [("leaf%d" % i, synth.leaf_image(i)) for i in range(6)]
    ims += [("adv64_" + k, v) for k, v in synth.adversarial_images(64, 64).items()]
    ims += [("adv61x97_" + k, v) for k, v in synth.adversarial_images(61, 97).items()]
    return ims


@pytest.mark.parametrize("thr", [(30, 100, False), (50, 150, False), (50, 150, True), (80, 160, True), (30, 100, True)])
def test_canny(dev, thr):
    lo, hi, l2 = thr
    for name, im in _images():
        g = sc.rgb_to_gray(im)
        got = ops.canny(up(g[None], dev), lo, hi, l2).cpu().numpy()[0]
        assert np.array_equal(got, sf.canny(g, lo, hi, l2)), (name, thr)


def test_canny_noise_many_runs(dev):
    g = np.random.default_rng(0).integers(0, 256, (2, 256, 256), dtype=np.uint8)   # > 2048 runs: global run tables
    got = ops.canny(up(g, dev), 50, 150, True).cpu().numpy()
    for i in range(2):
        assert np.array_equal(got[i], sf.canny(g[i], 50, 150, True))


@pytest.mark.parametrize("which", ["inclusive", "enhanced"])
def test_raw_mask_front_end(dev, which):
    scfg = sm.Cfg(mask_strategy=which)
    cfg = ops.mask_cfg("external")
    for name, im in _images():
        got = ops.raw_mask_front_end(up(im[None], dev), which, cfg).cpu().numpy()[0]
        exp = sm.mask_inclusive(im, scfg) if which == "inclusive" else sm.mask_enhanced(im, scfg)
        assert np.array_equal(got, exp), (which, name, int((got != exp).sum()))


@pytest.mark.parametrize("which", ["inclusive", "enhanced"])
def test_make_mask_default_strategies(which):
    """The reference's default strategy end to end through the drop-in make_mask."""
    cfg = transform.default_config(mask_strategy=which)
    scfg = sm.Cfg(mask_strategy=which)
    for name, im in _images()[:10]:
        mask, cnt = transform.make_mask(im, cfg)
        exp, info = sm.make_mask(im, scfg)
        assert np.array_equal(mask, exp), (which, name)
        assert (cnt is None) == (info is None)
        if cnt is not None:
            assert transform.bounding_rect(cnt) == info["bbox"]


def test_brown_filter():
    cfg = transform.default_config(mask_strategy="hsv_h")
    scfg = sm.Cfg(mask_strategy="hsv_h")
    for i in range(6):
        im = synth.leaf_image(i)
        m, _ = sm.make_mask(im, scfg)
        masked = sm.apply_mask(im, m, "white")
        vis, pct, count = filters.apply_brown_filter(masked, m, cfg)
        filt, epct, ecount = sm.brown_spots(masked, m, scfg)
        ev = masked.copy(); ev[filt > 0] = (255, 100, 0)
        assert np.array_equal(vis, ev) and count == ecount and pct == epct
    assert filters.apply_brown_filter(im, None, cfg) == (im, 0.0, 0)


def test_saliency_blur():
    cfg = transform.default_config(mask_strategy="hsv_h")
    scfg = sm.Cfg(mask_strategy="hsv_h")
    worst = 0
    for i in range(6):
        im = synth.leaf_image(i)
        m, _ = sm.make_mask(im, scfg)
        masked = sm.apply_mask(im, m, "white")
        got = filters.apply_blur_filter(masked, cfg, lambda r: transform.make_mask(r, cfg))
        m2, _ = sm.make_mask(masked, scfg)
        exp = sm.saliency_blur(masked, m2, scfg)
        d = np.abs(got.astype(int) - exp.astype(int))
        worst = max(worst, int(d.max()))
        assert d.max() <= 1, i          # tolerance: +-1 LSB (float32 min-max normalisations, north_star)
        assert np.array_equal(got[..., 0], got[..., 1]) and np.array_equal(got[..., 0], got[..., 2])
        assert not got[m2 == 0].any()
    print("saliency worst abs diff", worst)


def test_analyze_record_and_hist_stats():
    cfg = transform.default_config(mask_strategy="hsv_h")
    scfg = sm.Cfg(mask_strategy="hsv_h")
    for i in range(4):
        im = synth.leaf_image(i)
        mask, cnt = transform.make_mask(im, cfg)
        masked = transform.apply_mask(im, mask, "white")
        rec = filters.analyze_record(masked, mask, cnt)
        exp = spc.analyze_record(cnt)
        for k in ("centroid", "left", "right", "top", "bottom"):
            assert tuple(int(v) for v in rec[k]) == tuple(int(v) for v in exp[k]), k
        ev = (sf.canny(sc.rgb_to_gray(masked), 80, 160, True) > 0) & (mask > 0)
        assert np.array_equal(rec["veins"], ev)
        st = filters.histogram_stats(masked)
        cn = sm.hist_counters(masked)
        assert st["total_pixels"] == cn[0]
        assert list(st["hue_ranges"].values()) == [int(v) for v in cn[9:14]]
        assert np.array_equal(st["hsv_hist"], sm.hsv_hist_leaf(masked))
        for j, k in enumerate(filters.HIST_CATEGORIES):
            assert st["color_analysis"][k] == (int(cn[1 + j]) / int(cn[0])) * 100


def test_engine_default_strategy_pipeline():
    """TransformEngine(front='inclusive'): the batched core profile with the reference's default strategy, device- and
    host-resident, against the oracle (mask, bbox, blur, ROI, histograms)."""
    import torch
    from leaffliction_b200 import engine, ops
    from oracle import spec_filters as sf
    dev = torch.device("cuda:0")
    imgs = synth.leaf_batch(6, 256, 256, seed=3)
    cfg = ops.mask_cfg("hsv_h")            # numeric fields; the raw candidate comes from the front end
    eng = engine.TransformEngine(256, 256, cfg, 1.5, (256, 256), dev, chunk=4, front="inclusive")
    out = eng.run_device(torch.from_numpy(imgs).to(dev))
    host = eng.run_host(torch.from_numpy(imgs))
    scfg = sm.Cfg(mask_strategy="inclusive")
    for i in range(len(imgs)):
        m, einfo = sm.make_mask(imgs[i], scfg)
        masked = sm.apply_mask(imgs[i], m, "white")
        for o in (out, host):
            assert np.array_equal(np.asarray(o.mask[i].cpu()), m), i
            assert np.array_equal(np.asarray(o.blur[i].cpu()), sf.gaussian_blur_u8(imgs[i], 5, 1.5))
            assert tuple(np.asarray(o.info[i, 1:5].cpu())) == einfo["bbox"]
            assert np.array_equal(np.asarray(o.roi[i].cpu()), sm.roi_letterbox(masked, einfo["bbox"], (256, 256)))
            assert np.array_equal(np.asarray(o.hist9[i].cpu()), sm.hist9(imgs[i], m))
            assert np.array_equal(np.asarray(o.counters[i, :14].cpu()), sm.hist_counters(masked))
