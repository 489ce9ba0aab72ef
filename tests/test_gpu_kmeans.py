"""GPU parity of the k-means mask candidate (SURVEY.md 8a tier C row c2, 8f rank 4): lfx_kmeans_raw against
oracle/spec_kmeans.py, which tests/test_reference_differential.py pins bit for bit on cv2.kmeans and on the reference's
`_create_kmeans_mask` (srcs/transform/filters/mask.py:109-140).  Everything here is bit-exact: centres (float32), iteration
counts, candidate masks for the three bg_bias settings, `mask_strategy: kmeans` through the drop-in make_mask."""
import dataclasses

import numpy as np
import pytest
import torch

from leaffliction_b200 import ops, synth, transform
from oracle import spec_kmeans as sk
from oracle import spec_mask as sm

pytestmark = pytest.mark.gpu


def _images():
    g = np.random.default_rng(11)
    leaves = [synth.leaf_image(100 + i, 256, 256) for i in range(40)]
    odd = [np.full((256, 256, 3), 77, np.uint8),                                                  # one colour: two empty clusters
           (g.integers(0, 2, (256, 256, 1), dtype=np.uint8) * 200).repeat(3, 2),                    # two colours: one empty cluster
           g.integers(0, 256, (256, 256, 3), dtype=np.uint8)]                                      # noise: runs to the iteration cap
    return leaves + odd


def test_kmeans_centres_iterations_and_masks_match_the_oracle(dev):
    imgs = np.stack(_images())
    x = torch.from_numpy(imgs).to(dev)
    exp = [sk.kmeans3(im.reshape(-1, 3)) for im in imgs]
    for bias in ("auto", "light_bg", "dark_bg"):
        raw, cen, ki = ops.kmeans_raw(x, (25, 100), bias, details=True)
        raw, cen, ki = raw.cpu().numpy(), cen.cpu().numpy(), ki.cpu().numpy()
        scfg = sm.Cfg(mask_strategy="kmeans", bg_bias=bias)
        for i, (labels, centers, it) in enumerate(exp):
            assert np.array_equal(cen[i], centers), (bias, i, cen[i], centers)
            assert ki[i, 1] == it and ki[i, 3] == imgs[i].shape[0] * imgs[i].shape[1], (bias, i, ki[i], it)
            pick = sk.pick_cluster(centers, scfg)
            assert ki[i, 0] == pick, (bias, i)
            assert np.array_equal(raw[i], ((labels.reshape(256, 256) == pick) * 255).astype(np.uint8)), (bias, i)
    assert ki[40, 2] == 2 and ki[41, 2] == 1 and ki[0, 2] == 0           # empty-cluster events of the degenerate images
    assert max(e[2] for e in exp) >= 15                                  # long runs are covered


@pytest.mark.parametrize("shape", [(256, 192), (64, 256), (256, 256)])
def test_kmeans_other_shapes_with_a_256_long_side(dev, shape):
    H, W = shape
    imgs = np.stack([synth.leaf_image(7 + i, H, W) for i in range(6)])
    raw = ops.kmeans_raw(torch.from_numpy(imgs).to(dev), (25, 100), "light_bg").cpu().numpy()
    scfg = sm.Cfg(mask_strategy="kmeans")
    for i in range(len(imgs)):
        assert np.array_equal(raw[i], sk.kmeans_mask(imgs[i], scfg)), i


def test_kmeans_rejects_other_sizes(dev):
    from leaffliction_b200._lib import ERR_UNSUPPORTED, LeafxError
    with pytest.raises(LeafxError) as e:
        ops.kmeans_raw(torch.zeros((1, 128, 128, 3), dtype=torch.uint8, device=dev))
    assert e.value.code == ERR_UNSUPPORTED


def test_make_mask_with_kmeans_strategy(dev):
    """`mask_strategy: kmeans` through the drop-in make_mask (candidate -> _postprocess_mask -> fallback -> brown extension)."""
    imgs = synth.leaf_batch(12, 256, 256, 777)
    for bias in ("light_bg", "auto"):
        cfg = transform.default_config(mask_strategy="kmeans", bg_bias=bias, grabcut_refine=False, mask_upscale_factor=1.0,
                                       mask_upscale_long_side=0)
        scfg = sm.Cfg(mask_strategy="kmeans", bg_bias=bias)
        masks, info, contours = transform.make_mask_batch(imgs, cfg)
        for i in range(len(imgs)):
            om, oinfo = sm.make_mask(imgs[i], scfg)
            assert np.array_equal(masks[i], om), (bias, i, int((masks[i] != om).sum()))
            assert (contours[i] is None) == (oinfo is None)
            if oinfo is not None:
                assert tuple(info[i, 1:5]) == tuple(oinfo["bbox"])


@pytest.mark.parametrize("shape", [(128, 128), (300, 400), (512, 512), (100, 37)])
def test_kmeans_candidate_any_size(dev, shape):
    """Sizes whose longer side is not 256: INTER_AREA working copy (up or down), k-means, INTER_NEAREST back
    (mask.py:113-118,139) -- bit-exact against the oracle, which is pinned on the reference for these sizes too."""
    H, W = shape
    imgs = np.stack([synth.leaf_image(300 + i, H, W) for i in range(5)])
    cfg = transform.default_config(mask_strategy="kmeans", grabcut_refine=False, mask_upscale_factor=1.0, mask_upscale_long_side=0)
    raw = transform.kmeans_candidate(torch.from_numpy(imgs).to(dev), cfg).cpu().numpy()
    scfg = sm.Cfg(mask_strategy="kmeans")
    for i in range(len(imgs)):
        exp = sk.kmeans_mask(imgs[i], scfg)
        assert np.array_equal(raw[i], exp), (i, int((raw[i] != exp).sum()))


def test_kmeans_and_auto_against_the_reference_made_fixture(dev):
    """CUDA against tests/golden/golden_v1.npz directly (arrays produced by the reference's own `_create_kmeans_mask` and
    `make_mask(mask_strategy="auto")`, make_golden.py): 64x64 / 96x64 / 256x256 leaves and adversarial images."""
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden"))
    import make_golden as mg
    G = np.load(os.path.join(here, "golden", "golden_v1.npz"))
    imgs = mg.inputs()
    for k in [f for f in G.files if f.startswith("kmeans_raw/")]:
        _, bias, name = k.split("/")
        cfg = transform.default_config(mask_strategy="kmeans", bg_bias=bias, grabcut_refine=False, mask_upscale_factor=1.0,
                                       mask_upscale_long_side=0)
        got = transform.kmeans_candidate(torch.from_numpy(imgs[name][None]).to(dev), cfg).cpu().numpy()[0]
        assert np.array_equal(got, G[k]), f"{k}: {(got != G[k]).sum()} px differ"
    for strat in ("kmeans", "auto"):
        cfg = transform.default_config(mask_strategy=strat, grabcut_refine=False, mask_upscale_factor=1.0, mask_upscale_long_side=0)
        for k in [f for f in G.files if f.startswith(f"mask/{strat}/")]:
            name = k.split("/")[-1]
            masks, info, contours = transform.make_mask_batch(imgs[name][None], cfg)
            assert np.array_equal(masks[0], G[k]), f"{k}: {(masks[0] != G[k]).sum()} px differ"
            bk = f"bbox/{strat}/{name}"
            assert (contours[0] is not None) == (bk in G.files), k
            if contours[0] is not None:
                assert tuple(int(v) for v in info[0, 1:5]) == tuple(int(v) for v in G[bk]), k


def test_engine_with_the_kmeans_front(dev):
    """TransformEngine(front="kmeans"): the batched core profile on the k-means candidate -- mask, bbox and histograms against
    the oracle's make_mask(mask_strategy="kmeans") + hist9."""
    from leaffliction_b200 import engine
    imgs = synth.leaf_batch(10, 256, 256, 2024)
    eng = engine.TransformEngine(256, 256, ops.mask_cfg("hsv_h"), 1.5, (256, 256), dev, front="kmeans", bg_bias="light_bg")
    out = eng.run_device(torch.from_numpy(imgs).to(dev))
    torch.cuda.synchronize()
    scfg = sm.Cfg(mask_strategy="kmeans", bg_bias="light_bg")
    for i in range(len(imgs)):
        om, oinfo = sm.make_mask(imgs[i], scfg)
        assert np.array_equal(out.mask[i].cpu().numpy(), om), i
        if oinfo is not None:
            assert tuple(out.info[i, 1:5].cpu().numpy()) == tuple(oinfo["bbox"]), i
        assert np.array_equal(out.hist9[i].cpu().numpy(), sm.hist9(imgs[i], om)), i
