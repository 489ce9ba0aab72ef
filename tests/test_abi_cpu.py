"""CPU tests of the C-ABI boundary: libleafx.so loads without a GPU, exports every symbol include/leafx.h
declares, fails loudly (no CPU fallback) on compute calls, and its HOST helpers match the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from leaffliction_b200 import _lib
from oracle import spec_augment as sa
from oracle import spec_filters as sf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "leafx.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(lfx_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported():
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 28
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/leafx.h but not exported by libleafx.so: {missing}"
    bound = set(_lib.exported_symbols())
    assert set(names) <= bound, f"not bound in leaffliction_b200/_lib.py: {sorted(set(names) - bound)}"


def test_version_and_error_string():
    lib = _lib.load()
    assert lib.lfx_version() >= 100
    lib.lfx_last_error.restype = C.c_char_p
    assert isinstance(lib.lfx_last_error(), (bytes, type(None)))


def test_no_cpu_fallback():
    """Without a successful lfx_init on a CUDA device every compute entry point returns LFX_ERR_CUDA."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the -m gpu suite")
    lib = _lib.load()
    assert lib.lfx_init(0) == -2                               # LFX_ERR_CUDA: no device
    buf = (C.c_uint8 * 64)()
    mode = (C.c_int32 * 1)(0)
    rc = lib.lfx_flip(C.cast(buf, C.c_void_p), C.cast(buf, C.c_void_p), 1, 4, 4, C.cast(mode, C.c_void_p), None, 0, None)
    assert rc == -2
    assert b"lfx_init" in lib.lfx_last_error()
    with pytest.raises(RuntimeError):
        from leaffliction_b200 import engine
        engine.TransformEngine(256, 256)


def test_host_gauss_taps_match_oracle():
    lib = _lib.load()
    for k, sigma in ((5, 1.5), (15, 0.0), (3, 0.0), (7, 2.0), (5, 0.8)):
        taps = (C.c_int32 * 31)()
        assert lib.lfx_gauss_taps(k, C.c_double(sigma), taps) == 0
        assert list(taps)[:k] == sf.gaussian_kernel_q8(k, sigma).tolist(), (k, sigma)


def test_host_lanczos_tables_match_oracle():
    lib = _lib.load()
    for insz, outsz in ((208, 256), (243, 256), (256, 224), (1024, 224), (64, 64)):
        ks = lib.lfx_lanczos_ksize(insz, outsz)
        bounds = np.zeros((outsz, 2), np.int32)
        kk = np.zeros((outsz, ks), np.int32)
        rc = lib.lfx_lanczos_table(insz, outsz, ks, bounds.ctypes.data_as(C.c_void_p), kk.ctypes.data_as(C.c_void_p))
        assert rc == ks                                        # returns the tap count (lfx_api.cu)
        eks, eb, ek = sa.lanczos_coeffs(insz, outsz)
        assert eks == ks
        assert np.array_equal(bounds, eb)
        assert np.array_equal(kk, ek)


def test_cubic_table_matches_oracle_taps():
    """lfx_cubic_table (host side, no GPU) == the oracle's float32 restatement of OpenCV's cubic weights."""
    import ctypes as C

    import numpy as np

    from leaffliction_b200 import _lib
    from oracle import spec_filters as sf
    lib = _lib.load()
    for a, b in ((256, 333), (97, 126), (64, 1500), (333, 256), (40, 40)):
        f = np.zeros(b, np.int32)
        w = np.zeros((b, 4), np.int32)
        assert lib.lfx_cubic_table(a, b, f.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p)) == 0
        s, ww = sf.cubic_taps(a, b)
        assert np.array_equal(f, s) and np.array_equal(w, ww), (a, b)


def test_jpeg_library_exports_header_symbols_and_has_no_cpu_fallback():
    """include/leafx_jpeg.h (nvJPEG file boundary): libleafx_jpeg.so loads, exports every declared symbol, and refuses
    to initialise without a CUDA device."""
    import torch

    from leaffliction_b200 import jpegio
    txt = open(os.path.join(ROOT, "include", "leafx_jpeg.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = sorted(set(re.findall(r"\b(lfx_jpeg_[a-z0-9_]+)\s*\(", txt)))
    assert len(names) == 8
    lib = jpegio.load()
    assert not [n for n in names if not hasattr(lib, n)]
    assert set(names) == set(jpegio.exported_symbols())
    if not torch.cuda.is_available():
        assert lib.lfx_jpeg_init(0, 0, 1) == -2
        assert b"no CPU fallback" in lib.lfx_jpeg_last_error()
        assert lib.lfx_jpeg_backend() == -1
        with pytest.raises(jpegio.JpegError):
            jpegio.decode_batch([b"x"], 8, 8)
