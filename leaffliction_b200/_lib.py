"""ctypes binding of libleafx.so (include/leafx.h).  There is no CPU fallback: if the shared
library is missing and cannot be built, or no sm_100 device is present when an op is called,
this module raises -- it never substitutes another implementation."""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libleafx.so")

OK, ERR_ARG, ERR_CUDA, ERR_UNSUPPORTED, ERR_WORKSPACE = 0, -1, -2, -3, -4


class LeafxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libleafx error {code}: {msg}")
        self.code = code


class MaskCfg(C.Structure):
    """lfx_mask_cfg (include/leafx.h)."""
    _fields_ = [(n, C.c_int32) for n in (
        "strategy", "green_lo", "green_hi", "fill_size", "morph_kernel", "brown_lo", "brown_hi",
        "brown_s_min", "brown_v_max", "brown_min_area_px", "brown_morph_kernel", "use_lab_brown",
        "lab_a_min", "lab_b_min", "fallback_channel", "bg_dark", "extend_brown")] + [("reserved", C.c_int32 * 3)]


_P = C.c_void_p
_I = C.c_int
_SIGS = {
    "lfx_version": (C.c_int, []),
    "lfx_init": (C.c_int, [_I]),
    "lfx_last_error": (C.c_char_p, []),
    "lfx_flip": (C.c_int, [_P, _P, _I, _I, _I, _P, _P, _I, _P]),
    "lfx_rotate_nn": (C.c_int, [_P, _P, C.c_int64, _I, _I, _I, _P, _I, _P, _I, _P]),
    "lfx_warp_bicubic": (C.c_int, [_P, _P, _I, _I, _I, _P, _P, _P, _I, _P]),
    "lfx_lanczos_ksize": (C.c_int, [_I, _I]),
    "lfx_lanczos_table": (C.c_int, [_I, _I, _I, _P, _P]),
    "lfx_crop_lanczos": (C.c_int, [_P, _P, _P, _I, _I, _I, _P, _I, _I, _P, _P, _I, _P, _P, _I, _P]),
    "lfx_distort": (C.c_int, [_P, _P, _P, _I, _I, _I, _P, _P, _P, _I, _P]),
    "lfx_cvt_color": (C.c_int, [_P, _P, _I, _I, _I, _I, _P]),
    "lfx_threshold_mask": (C.c_int, [_P, _P, _I, _I, _I, C.POINTER(MaskCfg), _P]),
    "lfx_make_mask_workspace": (C.c_size_t, [_I, _I, _I]),
    "lfx_make_mask": (C.c_int, [_P, _P, _P, _P, _I, _I, _I, C.POINTER(MaskCfg), _P, C.c_size_t, _P]),
    "lfx_postprocess_mask": (C.c_int, [_P, _P, _P, _I, _I, _I, _I, _I, _P, C.c_size_t, _P]),
    "lfx_trace_contour": (C.c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "lfx_apply_mask": (C.c_int, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "lfx_analyze_workspace": (C.c_size_t, [_I, _I]),
    "lfx_analyze_record": (C.c_int, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, C.c_size_t, _P]),
    "lfx_analyze_overlay": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "lfx_draw_primitives": (C.c_int, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "lfx_draw_rectangles": (C.c_int, [_P, _P, _P, _I, _I, _I, C.c_uint32, _I, _P]),
    "lfx_strategy_raw": (C.c_int, [_P, _P, _I, _I, _I, C.POINTER(MaskCfg), _P, C.c_size_t, _P]),
    "lfx_kmeans_raw": (C.c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, C.c_uint32, _P]),
    "lfx_score_features": (C.c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "lfx_gauss_taps": (C.c_int, [_I, C.c_double, _P]),
    "lfx_gauss_u8": (C.c_int, [_P, _P, _I, _I, _I, _I, _I, C.c_double, _P]),
    "lfx_roi_letterbox": (C.c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "lfx_color_stats": (C.c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "lfx_front_workspace": (C.c_size_t, [_I, _I, _I]),
    "lfx_canny": (C.c_int, [_P, _P, _I, _I, _I, C.c_double, C.c_double, _I, _P, C.c_size_t, _P]),
    "lfx_raw_mask": (C.c_int, [_P, _P, _I, _I, _I, _I, C.POINTER(MaskCfg), _P, C.c_size_t, _P]),
    "lfx_brown_spots": (C.c_int, [_P, _P, _P, _P, _I, _I, _I, C.POINTER(MaskCfg), _P, C.c_size_t, _P]),
    "lfx_saliency_blur": (C.c_int, [_P, _P, _P, _I, _I, _I, C.c_double, C.POINTER(MaskCfg), _P, C.c_size_t, _P]),
    "lfx_pipeline_core_workspace": (C.c_size_t, [_I, _I, _I]),
    "lfx_legacy_normal_u8": (C.c_int, [_P, _P, _I, _I, C.c_double, C.c_double, _P]),
    "lfx_draw_augment_params": (C.c_int, [_P, _P, _I, _I, _I, _P, _P, _I]),
    "lfx_seed_words": (C.c_int, [_P, _I, _I, _P, _P]),
    "lfx_draw_augment_params_words": (C.c_int, [_P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "lfx_draw_balance_tasks": (C.c_int, [C.c_uint32, _I, _P, _P, _P, _P]),
    "lfx_cubic_table": (C.c_int, [_I, _I, _P, _P]),
    "lfx_resize_cubic": (C.c_int, [_P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "lfx_resize_nearest": (C.c_int, [_P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "lfx_pipeline_core": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, C.c_double,
                                    C.POINTER(MaskCfg), _P, C.c_size_t, _P, _P, _P]),
}

_lib = None
_lock = threading.Lock()
_inited_device = None


def exported_symbols():
    """Names declared in include/leafx.h that the library must export."""
    return sorted(_SIGS)


def load() -> C.CDLL:
    """dlopen libleafx.so (building it in-tree first when nvcc is available and it is stale)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        from . import build as _build
        try:
            _build.build()          # returns at once when the library is newer than every source
        except RuntimeError:
            if not os.path.exists(LIB_PATH):
                raise               # a prebuilt library on a box without nvcc is fine; none at all is not
        if not os.path.exists(LIB_PATH):
            raise LeafxError(ERR_CUDA, f"{LIB_PATH} is missing and could not be built; there is no CPU fallback")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)  # AttributeError here == header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def check(rc: int):
    if rc != OK:
        raise LeafxError(rc, (load().lfx_last_error() or b"").decode())


def init(device: int = 0):
    """lfx_init on `device`; raises LeafxError when no sm_100 GPU is available."""
    global _inited_device
    lib = load()
    if _inited_device != device:
        check(lib.lfx_init(int(device)))
        _inited_device = device
    return lib
