"""Import the reference (/root/reference) in the BUILD container with test-only stand-ins for
PlantCV and matplotlib.  Used only by make_golden.py and by tests marked
`needs_reference` (skipped when /root/reference is absent, e.g. on the GPU box)."""
from __future__ import annotations

import os
import sys
import types

import numpy as np

REF = "/root/reference"


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "srcs"))


def install_standins():
    import cv2
    from scipy import ndimage as ndi

    if "plantcv" not in sys.modules:
        pcv = types.ModuleType("plantcv.plantcv")

        def fill(bin_img, size):
            b = bin_img.astype(bool)
            lab, _ = ndi.label(b)
            sizes = np.bincount(lab.ravel())
            small = sizes < size
            small[0] = False
            b[small[lab]] = False
            return b.astype(np.uint8) * 255

        def rgb2gray_hsv(rgb_img, channel):
            return cv2.cvtColor(rgb_img, cv2.COLOR_BGR2HSV)[..., "hsv".index(channel)]

        thr = types.SimpleNamespace(otsu=lambda gray_img, object_type: cv2.threshold(
            gray_img, 0, 255,
            (cv2.THRESH_BINARY if object_type == "light" else cv2.THRESH_BINARY_INV) + cv2.THRESH_OTSU)[1])
        pcv.fill = fill
        pcv.rgb2gray_hsv = rgb2gray_hsv
        pcv.threshold = thr
        pcv.params = types.SimpleNamespace(debug=None)
        pcv.analyze_object = lambda img, obj, mask: img
        root = types.ModuleType("plantcv")
        root.plantcv = pcv
        sys.modules["plantcv"] = root
        sys.modules["plantcv.plantcv"] = pcv
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        mpl.use = lambda *a, **k: None
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    if REF not in sys.path:
        sys.path.insert(0, REF)


def load():
    """Returns a namespace with the reference modules on the hot path."""
    install_standins()
    import importlib
    ns = types.SimpleNamespace()
    ns.T = importlib.import_module("srcs.cli.Transformation")
    ns.mask = importlib.import_module("srcs.transform.filters.mask")
    ns.blur = importlib.import_module("srcs.transform.filters.blur")
    ns.roi = importlib.import_module("srcs.transform.filters.roi")
    ns.brown = importlib.import_module("srcs.transform.filters.brown")
    ns.hist = importlib.import_module("srcs.transform.filters.hist")
    ns.analyze = importlib.import_module("srcs.transform.filters.analyze")
    ns.mask_utils = importlib.import_module("srcs.utils.mask_utils")
    ns.augmenter = importlib.import_module("srcs.preprocessing.image_augmenter")
    ns.components = importlib.import_module("srcs.preprocessing.dataset_components")
    ns.image_utils = importlib.import_module("srcs.utils.image_utils")
    return ns


def ref_config(ns, **over):
    """The reference's own YAML as a TransformConfig, with parity profile P0 applied
    (grabcut_refine false, no upscale) unless overridden."""
    import dataclasses
    from pathlib import Path
    cfg = ns.T.load_config(Path(REF) / "srcs/transform/config.yaml")
    base = dict(grabcut_refine=False, mask_upscale_factor=1.0, mask_upscale_long_side=0)
    base.update(over)
    return dataclasses.replace(cfg, **base)
