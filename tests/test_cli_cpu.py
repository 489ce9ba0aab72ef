"""CPU tests of the command-line surface: the flags of srcs/cli/Augmentation.py:32-78 and
srcs/cli/Transformation.py:568-608 are preserved (compared against the reference's own parsers when
/root/reference is present), type aliases, output names, config loading."""
import os
import sys

import pytest

from leaffliction_b200 import transform as T
from leaffliction_b200.cli import Augmentation as A
from leaffliction_b200.cli import Transformation as TC

HERE = os.path.dirname(os.path.abspath(__file__))


def _flags(parser):
    out = {}
    for a in parser._actions:
        if a.dest == "help":
            continue
        out[a.dest] = (tuple(a.option_strings), a.default, a.type.__name__ if a.type else None, a.nargs)
    return out


def test_augmentation_flags():
    f = _flags(A.build_parser())
    assert f["input_path"][0] == ()
    assert f["output"][0] == ("-out", "--output")
    assert f["seed"] == (("-seed", "--seed"), 42, "int", None)
    assert f["workers"][0] == ("--workers",)
    assert set(f) == {"input_path", "output", "seed", "workers"}
    a = A.parse_args(["images/", "-out", "o", "-seed", "7", "--workers", "3"])
    assert (a.input_path, a.output, a.seed, a.workers) == ("images/", "o", 7, 3)
    assert A.TRANSFORMATIONS == ["flip", "rotate", "skew", "shear", "crop", "distortion"]


def test_transformation_flags_and_types():
    f = _flags(TC.build_parser())
    assert set(f) == {"image", "out_dir", "src", "dst", "types", "config", "workers", "skip_existing", "overwrite", "preview"}
    assert f["src"][0] == ("-src", "--src") and f["dst"][0] == ("-dst", "--dst")
    assert f["config"][1] == "srcs/transform/config.yaml"
    assert f["workers"][1] == 0
    assert T.build_types_filter(None) == T.DEFAULT_TYPES
    assert T.build_types_filter("mask, ROI,histogram,spots,bogus,mask") == ("Mask", "ROI", "Hist", "Brown")
    assert T.build_types_filter("nothing") == T.DEFAULT_TYPES
    assert T.output_names("leaf (3)")["Mask"] == "leaf (3)__T_Mask.jpg"


def test_packaged_config_has_all_reference_keys():
    cfg = T.load_config(TC.PACKAGED_CONFIG)
    assert cfg.mask_strategy == "inclusive" and tuple(cfg.roi_size) == (256, 256) and cfg.gaussian_sigma == 1.5
    assert cfg.fill_size == 1000 and cfg.morph_kernel == 3 and tuple(cfg.brown_hue_range) == (0, 30)


def test_load_config_missing_key_exits(tmp_path):
    p = tmp_path / "bad.yaml"
    p.write_text("gaussian_sigma: 1.5\n")
    with pytest.raises(SystemExit) as e:
        T.load_config(p)
    assert e.value.code == 1
    with pytest.raises(SystemExit):
        T.load_config(tmp_path / "absent.yaml")


def test_augmentation_cli_exit_code_on_missing_input():
    with pytest.raises(SystemExit) as e:
        A.main(["/nonexistent/path.jpg"])
    assert e.value.code == 1


@pytest.mark.needs_reference
def test_flags_equal_reference_parsers(monkeypatch):
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import ref_harness
    ns = ref_harness.load()
    import importlib
    ref_aug = importlib.import_module("srcs.cli.Augmentation")
    captured = {}
    import argparse
    orig = argparse.ArgumentParser.parse_args

    def grab(self, *a, **k):
        captured["p"] = self
        raise SystemExit(0)
    monkeypatch.setattr(argparse.ArgumentParser, "parse_args", grab)
    for mod, ours in ((ref_aug, A.build_parser()), (ns.T, TC.build_parser())):
        with pytest.raises(SystemExit):
            mod.parse_args()
        ref = _flags(captured["p"])
        mine = _flags(ours)
        assert set(ref) == set(mine)
        for k in ref:
            assert ref[k][0] == mine[k][0], k                        # same option strings
            assert ref[k][2:] == mine[k][2:], k                      # same type / nargs
            assert ref[k][1] == mine[k][1], k                        # same default
    monkeypatch.setattr(argparse.ArgumentParser, "parse_args", orig)
    # the packaged YAML equals the reference's values for the 28 required keys
    from pathlib import Path
    rc = ns.T.load_config(Path(ref_harness.REF) / "srcs/transform/config.yaml")
    mc = T.load_config(TC.PACKAGED_CONFIG)
    for name, _ in T._CONFIG_FIELDS:
        a, b = getattr(rc, name), getattr(mc, name)
        assert (tuple(a) if isinstance(a, (list, tuple)) else a) == (tuple(b) if isinstance(b, (list, tuple)) else b), name
