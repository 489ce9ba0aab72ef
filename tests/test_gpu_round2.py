"""GPU tests of the round-2 boundary additions: source-index gathers inside the augment kernels, the dataset-level
histogram accumulated by k_core, the 6-op AugmentSet against the oracle, run_host's host-side completion."""
import random

import numpy as np
import pytest
import torch

from leaffliction_b200 import augment, engine, ops, synth
from oracle import refcalls
from oracle import spec_augment as sa
from oracle import spec_mask as sm

pytestmark = pytest.mark.gpu


def up(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def _seeds(n, seed=5):
    return np.random.default_rng(seed).integers(1, 1000001, size=(6, n), dtype=np.int64)


# ----------------------------------------------------------------------------- src_index == gather + op
@pytest.mark.parametrize("hw", [(256, 256), (61, 97)])
def test_src_index_equals_gather(dev, hw):
    H, W = hw
    rng = np.random.default_rng(11)
    data = up(rng.integers(0, 256, (9, H, W, 3), dtype=np.uint8), dev)
    idx = torch.tensor([8, 0, 3, 3, 7, 1, 8], dtype=torch.int32, device=dev)
    g = data.index_select(0, idx.long())
    B = len(idx)
    mode = torch.tensor([0, 1, 0, 1, 1, 0, 0], dtype=torch.int32, device=dev)
    assert torch.equal(ops.flip(data, mode, src_index=idx), ops.flip(g, mode))
    ip, dp = augment.draw_params_batch(np.repeat(np.arange(6, dtype=np.int32), B), _seeds(B).reshape(-1), H, W)
    ip, dp = ip.reshape(6, B, 8), dp.reshape(6, B, 8)
    a, _ = ops.rotate_nn(data, ip[1], 255, src_index=idx)
    b, _ = ops.rotate_nn(g, ip[1], 255)
    for i in range(B):
        n = int(ip[1, i, 6]) * int(ip[1, i, 7]) * 3
        assert torch.equal(a[i, :n], b[i, :n])
    for k in (2, 3):
        assert torch.equal(ops.warp_bicubic(data, dp[k], ip[k, :, 0], src_index=idx), ops.warp_bicubic(g, dp[k], ip[k, :, 0]))
    assert torch.equal(ops.crop_lanczos(data, ip[4, :, :4], (H, W), src_index=idx), ops.crop_lanczos(g, ip[4, :, :4], (H, W)))
    noise = up(rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8), dev)
    assert torch.equal(ops.distort(data, noise, ip[5, :, 0], src_index=idx), ops.distort(g, noise, ip[5, :, 0]))


def test_src_index_validation(dev):
    x = torch.zeros((2, 8, 8, 3), dtype=torch.uint8, device=dev)
    with pytest.raises(ValueError):
        ops.flip(x, [True], src_index=torch.zeros(1, dtype=torch.int64, device=dev))


# ----------------------------------------------------------------------------- AugmentSet vs the reference's calls
def test_augment_set_matches_reference_calls(dev):
    """Every op of the 6-op set, with the task seeds of `_process_single_transformation`, against the reference's own
    Pillow / NumPy calls on the same arrays (oracle/refcalls.augment_task) -- parameters AND pixels, bit-exact."""
    B = 6
    imgs = synth.leaf_batch(B, 256, 256, 77)
    seeds = _seeds(B, 3)
    s = augment.AugmentSet(B, 256, 256, dev).run(up(imgs, dev), seeds)
    torch.cuda.synchronize()
    got = {k: getattr(s, k).cpu().numpy() for k in ("flip", "skew", "shear", "crop", "distortion")}
    slab = s.rotate.cpu().numpy()
    for i in range(B):
        for k, op in enumerate(refcalls.AUGMENT_OPS):
            exp = refcalls.augment_task(imgs[i], op, int(seeds[k, i]))
            if op == "rotate":
                nh, nw = s.rotate_hw[i]
                assert exp.shape == (nh, nw, 3)
                assert np.array_equal(slab[i, : nh * nw * 3].reshape(nh, nw, 3), exp), f"rotate image {i}"
            else:
                assert np.array_equal(got[op][i], exp), f"{op} image {i}"


def test_augment_set_rejects_seed_zero(dev):
    x = torch.zeros((1, 32, 32, 3), dtype=torch.uint8, device=dev)
    sd = np.ones((6, 1), np.int64)
    sd[2, 0] = 0
    with pytest.raises(ValueError):
        augment.AugmentSet(1, 32, 32, dev).run(x, sd)


# ----------------------------------------------------------------------------- dataset histogram inside k_core
@pytest.mark.parametrize("hw,n", [((256, 256), 700), ((64, 96), 40), ((61, 97), 5)])
def test_dataset_histogram_accumulates(dev, hw, n):
    """dataset_hist += sum over the batch of hist9, by the fused kernel (more images than resident blocks) and by the
    general path (61x97 is not a fused shape); two calls accumulate."""
    H, W = hw
    base = synth.leaf_batch(min(n, 24), H, W, 9)
    x = up(np.concatenate([base] * ((n + len(base) - 1) // len(base)))[:n], dev)
    ds = torch.zeros((9, 256), dtype=torch.int64, device=dev)
    out = ops.pipeline_core(x, ops.mask_cfg("hsv_h"), 1.5, (H, W) if H * W < 65536 else (256, 256), dataset_hist=ds)
    exp = out.hist9.sum(dim=0, dtype=torch.int64)
    assert torch.equal(ds, exp)
    ops.pipeline_core(x, ops.mask_cfg("hsv_h"), 1.5, (H, W) if H * W < 65536 else (256, 256), out, dataset_hist=ds)
    assert torch.equal(ds, 2 * exp)
    # and the per-image histograms themselves are the oracle's
    m, _ = sm.make_mask(base[0], sm.Cfg(mask_strategy="hsv_h"))
    assert np.array_equal(out.hist9[0].cpu().numpy(), sm.hist9(base[0], m))


# ----------------------------------------------------------------------------- run_host returns filled host buffers
def test_run_host_outputs_ready_on_return(dev):
    """run_host used to return while the device-to-host copies were still queued.  A large batch read
    IMMEDIATELY after the call must already hold the results (compared with the device-resident run)."""
    B = 1536
    base = synth.leaf_batch(32, 256, 256, 21)
    imgs = torch.from_numpy(np.concatenate([base] * (B // 32))).pin_memory()
    seeds = _seeds(B, 8)
    eng = engine.TransformEngine(256, 256, ops.mask_cfg("hsv_h"), 1.5, (256, 256), dev, chunk=512, augment=True)
    host = engine.alloc_host_outputs(B, 256, 256, (256, 256), augment=True)
    for t in (host.blur, host.mask, host.roi, host.aug["crop"], host.aug["distortion"], host.aug["rotate"]):
        t.fill_(7)
    eng.run_host(imgs, host, seeds=seeds)
    snap = {k: v.clone() for k, v in (("blur", host.blur), ("mask", host.mask), ("roi", host.roi), ("hist9", host.hist9),
                                      ("crop", host.aug["crop"]), ("dist", host.aug["distortion"]), ("rot", host.aug["rotate"]))}
    torch.cuda.synchronize()
    x = imgs.to(dev)
    ref = ops.pipeline_core(x, ops.mask_cfg("hsv_h"), 1.5, (256, 256))
    aset = augment.AugmentSet(B, 256, 256, dev).run(x, seeds)
    torch.cuda.synchronize()
    assert torch.equal(snap["blur"], ref.blur.cpu()) and torch.equal(snap["mask"], ref.mask.cpu())
    assert torch.equal(snap["roi"], ref.roi.cpu()) and torch.equal(snap["hist9"], ref.hist9.cpu())
    assert torch.equal(snap["crop"], aset.crop.cpu()) and torch.equal(snap["dist"], aset.distortion.cpu())
    rot = aset.rotate.cpu()
    for i in (0, 511, 512, B - 1):
        nh, nw = host.rotate_hw[i]
        assert tuple(aset.rotate_hw[i]) == (nh, nw)
        assert torch.equal(snap["rot"][i, : nh * nw * 3], rot[i, : nh * nw * 3])


def test_run_host_ragged_tail(dev):
    B = 70
    imgs = torch.from_numpy(synth.leaf_batch(B, 64, 64, 2))
    seeds = _seeds(B, 9)
    eng = engine.TransformEngine(64, 64, ops.mask_cfg("hsv_h"), 1.5, (64, 64), dev, chunk=32, augment=True)
    host = eng.run_host(imgs, seeds=seeds)
    x = imgs.to(dev)
    aset = augment.AugmentSet(B, 64, 64, dev).run(x, seeds)
    ref = ops.pipeline_core(x, ops.mask_cfg("hsv_h"), 1.5, (64, 64))
    torch.cuda.synchronize()
    assert torch.equal(host.mask, ref.mask.cpu()) and torch.equal(host.aug["skew"], aset.skew.cpu())
    assert torch.equal(host.aug["flip"], aset.flip.cpu()) and torch.equal(host.aug["shear"], aset.shear.cpu())
    with pytest.raises(ValueError):
        eng.run_host(imgs)          # augment engine without seeds


# ----------------------------------------------------------------------------- analyze record, strategy raws, score, auto
def _dev_masks(imgs, dev, strategy="hsv_h"):
    return ops.make_mask(up(imgs, dev), ops.mask_cfg(strategy))


def test_analyze_records_batch_vs_cv2(dev):
    """lfx_analyze_record (analyze.py:43-98) for a batch: centroid / extreme points equal OpenCV's on the same contour
    (exact), hull vertex set and hull area equal cv2.convexHull / contourArea (exact), PCA axes equal cv2.PCACompute2
    up to sign (1e-4: OpenCV computes them in float32)."""
    cv2 = pytest.importorskip("cv2")
    from leaffliction_b200 import filters
    from oracle import spec_contour
    adv = synth.adversarial_images(64, 64)
    sets = [synth.leaf_batch(12, 256, 256, 31), np.stack([adv[k] for k in ("frame", "ties", "pinch", "salt")]),
            synth.leaf_batch(3, 96, 64, 5)]
    total_seen = 0
    for imgs in sets:
        H, W = imgs.shape[1:3]
        mask, info = _dev_masks(imgs, dev)
        rec = ops.analyze_records(mask, info, max_pts=4096, max_hull=512)
        ri, rf, hull = rec["rec_i"].cpu().numpy(), rec["rec_f"].cpu().numpy(), rec["hull"].cpu().numpy()
        pts, cnt = rec["points"].cpu().numpy(), rec["counts"].cpu().numpy()
        mask_h, info_h = mask.cpu().numpy(), info.cpu().numpy()
        for i in range(len(imgs)):
            if not info_h[i, 0]:
                assert ri[i, 0] == 0
                continue
            total_seen += 1
            cs, _ = cv2.findContours(mask_h[i], cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
            c = max(cs, key=cv2.contourArea)
            assert np.array_equal(pts[i, : cnt[i]].reshape(-1, 1, 2), c)              # the traced contour is OpenCV's
            r = filters.record_from_device(ri[i], rf[i], hull[i])
            exp = spec_contour.analyze_record(c)
            for k in ("centroid", "left", "right", "top", "bottom"):
                assert tuple(int(v) for v in r[k]) == tuple(int(v) for v in exp[k]), (i, k)
            M = cv2.moments(c)
            if M["m00"] != 0:
                assert r["centroid"] == (int(M["m10"] / M["m00"]), int(M["m01"] / M["m00"]))
                assert r["area"] == M["m00"]
            hv = cv2.convexHull(c)
            assert {tuple(p) for p in hv[:, 0, :]} == {tuple(p) for p in r["hull"][:, 0, :]}, i
            assert r["hull_area"] == cv2.contourArea(hv)
            data = c[:, 0, :].astype(np.float32)
            if len(data) >= 3:
                mean, evec, evals = cv2.PCACompute2(data, mean=None)
                assert np.allclose(r["pca_mean"], mean[0], atol=1e-3)
                assert np.allclose(r["pca_eigenvalues"], evals[:, 0], rtol=1e-3, atol=1e-3)
                if evals[0, 0] - evals[1, 0] > 1e-3 * max(1.0, evals[0, 0]):
                    for k in range(2):
                        assert abs(abs(float(np.dot(r["pca_eigenvectors"][k], evec[k]))) - 1.0) < 1e-4
                        proj = data @ evec[k]
                        got = {float(np.dot(np.array(p, np.float32), evec[k])) for p in r["axes"][k]}
                        assert abs(min(got) - float(proj.min())) < 1e-2 and abs(max(got) - float(proj.max())) < 1e-2


    assert total_seen >= 14


def test_analyze_record_small_hull_buffer(dev):
    imgs = synth.leaf_batch(2, 256, 256, 31)
    mask, info = _dev_masks(imgs, dev)
    rec = ops.analyze_records(mask, info, max_pts=4096, max_hull=4)
    assert (rec["rec_i"][:, 12].cpu().numpy() < 0).all()          # -needed, nothing written past the buffer
    rec2 = ops.analyze_records(mask, info, max_pts=16)            # contour buffer too small: no record, count reported
    assert (rec2["rec_i"][:, 0].cpu().numpy() == 0).all() and (rec2["rec_i"][:, 1].cpu().numpy() < 0).all()


@pytest.mark.parametrize("hw", [(256, 256), (64, 96), (61, 97)])
def test_strategy_raw_candidates(dev, hw):
    """_build_mask_candidates (mask.py:414-434): the raw candidate of each threshold strategy, before post-processing."""
    import dataclasses
    H, W = hw
    imgs = np.concatenate([synth.leaf_batch(5, H, W, 12), np.random.default_rng(2).integers(0, 256, (1, H, W, 3), dtype=np.uint8)])
    x = up(imgs, dev)
    for st, bias in (("hsv_h", "light_bg"), ("lab", "light_bg"), ("hsv_s", "light_bg"), ("hsv_s", "dark_bg"), ("hsv_v_dark", "light_bg")):
        got = ops.strategy_raw(x, ops.mask_cfg(st, bg_bias=bias)).cpu().numpy()
        for i in range(len(imgs)):
            exp = sm.raw_candidate(imgs[i], sm.Cfg(mask_strategy=st, bg_bias=bias))
            assert np.array_equal(got[i], (exp > 0).astype(np.uint8) * 255), (st, bias, i)


def test_score_features_and_auto_strategy(dev):
    """_score_mask terms on the device (mask.py:160-177) and `mask_strategy: auto` through the drop-in make_mask
    (mask.py:435-461 minus the Tier C k-means candidate) against the oracle: boundary strength within 1e-6 (float32
    mean in the reference, fp64 here), counts exact, final mask and bounding box bit-exact, same winner."""
    import dataclasses

    from leaffliction_b200 import transform
    imgs = np.concatenate([synth.leaf_batch(10, 256, 256, 4242), np.full((1, 256, 256, 3), 180, np.uint8)])
    x = up(imgs, dev)
    scfg = sm.Cfg(mask_strategy="auto")
    cfg = transform.default_config(mask_strategy="auto", grabcut_refine=False, mask_upscale_factor=1.0, mask_upscale_long_side=0)
    # terms for two candidates
    masks = []
    for st in ("hsv_h", "lab"):
        raw = ops.strategy_raw(x, ops.mask_cfg(st))
        m, _ = ops.postprocess_mask(raw, 1000, 3)
        masks.append(m)
    stack = torch.stack(masks).contiguous()
    feat, gmax, gmin = ops.score_features(x, stack, (25, 100))
    feat, gmax, gmin = feat.cpu().numpy(), gmax.cpu().numpy(), gmin.cpu().numpy()
    for k in range(2):
        mk = masks[k].cpu().numpy()
        for i in range(len(imgs)):
            b_exp, g_exp = sm.score_features(mk[i], imgs[i], scfg)
            bsum, bcnt, mpx, gpx = feat[k, i]
            assert mpx == (mk[i] > 0).sum()
            assert abs(gpx / max(1.0, mpx) - g_exp) < 1e-12
            rng_ = float(gmax[i]) - float(gmin[i])
            b = ((bsum / bcnt) - float(gmin[i])) / rng_ if (bcnt > 0 and rng_ > 0) else 0.0
            assert abs(b - b_exp) < 1e-6, (k, i, b, b_exp)
    # the whole auto path through the drop-in API
    got_masks, info, _contours = transform.make_mask_batch(imgs, cfg)
    raw, choice, scores = transform.auto_candidate(x, cfg)
    winners = set()
    for i in range(len(imgs)):
        om, oinfo, ochoice, oscore = sm.make_mask_auto(imgs[i], scfg, True)
        assert np.array_equal(got_masks[i], om), f"image {i}: {(got_masks[i] != om).sum()} px differ"
        assert (sm.AUTO_CANDIDATES[choice[i]] if choice[i] >= 0 else None) == ochoice, i
        if ochoice is not None:
            assert abs(scores[choice[i], i] - oscore) < 1e-6
            assert tuple(info[i, 1:5]) == tuple(oinfo["bbox"])
        winners.add(ochoice)
    assert len(winners) >= 2


# ----------------------------------------------------------------------------- images larger than 256x256 / than roi_size
def test_roi_letterbox_shrink(dev):
    """roi.py:35-38 when the bounding box is larger than roi_size: cv2.resize(INTER_AREA) averages down (general ratios,
    integer factors incl. 2x2) -- bit-exact against the oracle (itself pinned on OpenCV in test_oracle_golden.py)."""
    rng = np.random.default_rng(3)
    H, W = 300, 340
    imgs = rng.integers(0, 256, (6, H, W, 3), dtype=np.uint8)
    mask = (rng.integers(0, 4, (6, H, W), dtype=np.uint8) > 0).astype(np.uint8) * 255
    boxes = [(10, 20, 320, 270), (0, 0, 256, 256), (5, 7, 300, 128), (30, 40, 64, 50), (1, 1, 338, 298), (0, 0, 340, 300)]
    info = np.zeros((6, 8), np.int32)
    for i, (x, y, w, h) in enumerate(boxes):
        info[i, :5] = (1, x, y, w, h)
    for roi in ((128, 128), (100, 160)):
        got = ops.roi_letterbox(up(imgs, dev), up(mask, dev), up(info, dev), roi).cpu().numpy()
        for i, b in enumerate(boxes):
            white = sm.apply_mask(imgs[i], mask[i], "white")
            assert np.array_equal(got[i], sm.roi_letterbox(white, b, roi)), (roi, i)


@pytest.mark.parametrize("size", [512, 1024])
def test_default_strategy_on_large_images(dev, size):
    """The reference's default strategy (inclusive, config.yaml:7) and the other front ends on images larger than
    256x256: the planes live in global scratch instead of shared memory, the results stay bit-exact."""
    from leaffliction_b200 import transform
    imgs = synth.leaf_batch(2, size, size, 99)
    x = up(imgs, dev)
    for which in ("inclusive", "enhanced"):
        raw = ops.raw_mask_front_end(x, which, ops.mask_cfg("hsv_h")).cpu().numpy()
        for i in range(len(imgs)):
            exp = sm.raw_candidate(imgs[i], sm.Cfg(mask_strategy=which))
            assert np.array_equal(raw[i], (exp > 0).astype(np.uint8) * 255), (which, i)
    cfg = transform.default_config(mask_strategy="inclusive", grabcut_refine=False, mask_upscale_factor=1.0, mask_upscale_long_side=0)
    masks, info, contours = transform.make_mask_batch(imgs, cfg)
    for i in range(len(imgs)):
        om, oinfo = sm.make_mask(imgs[i], sm.Cfg(mask_strategy="inclusive"))
        assert np.array_equal(masks[i], om)
        assert tuple(info[i, 1:5]) == tuple(oinfo["bbox"])
        # ROI of a large image into the reference's 256x256 canvas (shrink) and the brown filter
        white = sm.apply_mask(imgs[i], om, "white")
        canvas, _vis, box = transform.apply_roi_filter(white, contours[i], cfg)
        assert np.array_equal(canvas, sm.roi_letterbox(white, box, cfg.roi_size))
    spots, stats = ops.brown_spots(x, up(masks, dev), ops.mask_cfg("hsv_h"))
    for i in range(len(imgs)):
        es, _pct, ecount = sm.brown_spots(imgs[i], masks[i], sm.Cfg())
        assert np.array_equal(spots[i].cpu().numpy(), es) and int(stats[i, 1]) == ecount


@pytest.mark.parametrize("hw,out", [((116, 116), (916, 916)), ((300, 400), (224, 300)), ((256, 256), (224, 224)), ((90, 64), (128, 96))])
def test_lanczos_dp4a_tap_counts(dev, hw, out):
    """Resize ratios around the 8-tap / 12-tap limits of the dp4a Lanczos kernel (a 116 -> 916 upscale has 7-tap rows, a
    400 -> 300 downscale 9-tap rows): uint8 and /255 float outputs against the oracle (Pillow's 22-bit fixed point)."""
    H, W = hw
    OH, OW = out
    rng = np.random.default_rng(17)
    imgs = rng.integers(0, 256, (3, H, W, 3), dtype=np.uint8)
    boxes = np.array([[0, 0, W, H], [1, 2, W - 3, H - 5], [W // 4, H // 4, W // 2, H // 2]], np.int32)
    got, gotf = ops.crop_lanczos(up(imgs, dev), boxes, (OH, OW), want_f32=True)
    got, gotf = got.cpu().numpy(), gotf.cpu().numpy()
    for i, (l, t, w, h) in enumerate(boxes):
        exp = sa.resize_lanczos(np.ascontiguousarray(imgs[i][t:t + h, l:l + w]), OW, OH)
        assert np.array_equal(got[i], exp), (i, (got[i] != exp).sum())
        assert np.array_equal(gotf[i], exp.astype(np.float32) / np.float32(255.0))


def test_seed_words_match_cpython_random(dev):
    """lfx_seed_words: the first outputs of Python's `random` after random.seed(s), one thread per seed (MT19937ar
    init_by_array restated on the device), against the interpreter -- and the AugmentSet built on it against the
    host-seeded AugmentSet (same parameters, same images)."""
    import random

    from leaffliction_b200 import augment
    rng = np.random.default_rng(5)
    seeds = np.concatenate([[1, 2, 42, 999983, 1000000, 2**31 - 1, 2**31, 2**32 - 1], rng.integers(1, 2**32, 3000)]).astype(np.uint32)
    d = torch.from_numpy(seeds.view(np.int32)).to(dev)
    for nw in (1, 16, 226):
        got = ops.seed_words(d, nw).cpu().numpy().view(np.uint32)
        for i in list(range(8)) + list(range(8, len(seeds), 97)):
            random.seed(int(seeds[i]))
            assert [random.getrandbits(32) for _ in range(nw)] == got[i].tolist(), (nw, i)
    B = 96
    x = up(synth.leaf_batch(B, 256, 256, 99), dev)
    task_seeds = rng.integers(1, 1000001, (6, B)).astype(np.int64)
    sa = augment.AugmentSet(B, 256, 256, dev, concurrent=True, device_seeding=True)
    assert sa.device_seeding
    a = sa.run(x, task_seeds)
    b = augment.AugmentSet(B, 256, 256, dev, device_seeding=False).run(x, task_seeds)
    torch.cuda.synchronize()
    for name in ("flip", "skew", "shear", "crop", "distortion"):
        assert torch.equal(getattr(a, name), getattr(b, name)), name
    assert np.array_equal(a.rotate_hw, b.rotate_hw)
    for i, (nh, nw) in enumerate(a.rotate_hw):          # the slab beyond an image's own extent is never written
        n = int(nh) * int(nw) * 3
        assert torch.equal(a.rotate[i, :n], b.rotate[i, :n]), ("rotate", i)


def test_pipelined_augment_set_equals_the_joined_one(dev):
    """AugmentSet(concurrent=True, pipelined=True): steps overlap on the side streams and the caller joins once (join());
    the outputs of the LAST step equal those of a plain AugmentSet run with the same seeds."""
    from leaffliction_b200 import augment
    rng = np.random.default_rng(12)
    B = 64
    x = up(synth.leaf_batch(B, 256, 256, 31), dev)
    p = augment.AugmentSet(B, 256, 256, dev, concurrent=True, pipelined=True)
    q = augment.AugmentSet(B, 256, 256, dev)
    for _ in range(4):
        seeds = rng.integers(1, 1000001, (6, B)).astype(np.int64)
        p.start(x, seeds)
        p.finish()
    p.join()
    q.run(x, seeds)
    torch.cuda.synchronize()
    for name in ("flip", "skew", "shear", "crop", "distortion"):
        assert torch.equal(getattr(p, name), getattr(q, name)), name
    assert np.array_equal(p.rotate_hw, q.rotate_hw)
    for i, (nh, nw) in enumerate(p.rotate_hw):
        n = int(nh) * int(nw) * 3
        assert torch.equal(p.rotate[i, :n], q.rotate[i, :n]), ("rotate", i)
