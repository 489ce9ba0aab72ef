"""JPEG files <-> device batches through libleafx_jpeg.so (nvJPEG; include/leafx_jpeg.h).

Replaces the reference's per-image host codec on the file boundary of the hot path:
`ImageLoader.load_pil_image` / `load_as_array` (srcs/utils/image_utils.py:19-47), `save_pil_image` (:49-59, quality 95)
and `imwrite_bgr` (srcs/cli/Transformation.py:196-205).  Pixels never visit the host: a folder of JPEGs becomes one
uint8 [B,H,W,3] CUDA tensor, a result batch becomes B bitstreams.  Parity is to JPEG tolerance (tests/test_gpu_jpeg.py).
There is no CPU fallback: without the library or a CUDA device every call raises."""
from __future__ import annotations

import ctypes as C
import os
import threading
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from typing import List, Optional, Sequence, Tuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libleafx_jpeg.so")

_P, _I = C.c_void_p, C.c_int
_SIGS = {
    "lfx_jpeg_init": (C.c_int, [_I, _I, _I]),
    "lfx_jpeg_shutdown": (None, []),
    "lfx_jpeg_last_error": (C.c_char_p, []),
    "lfx_jpeg_backend": (C.c_int, []),
    "lfx_jpeg_info": (C.c_int, [_P, C.c_size_t, _P, _P, _P, _P]),
    "lfx_jpeg_decode_batch": (C.c_int, [_P, _P, _P, _I, _I, _I, _P, _P]),
    "lfx_jpeg_encode_bound": (C.c_size_t, [_I, _I, _I, _I]),
    "lfx_jpeg_encode_batch": (C.c_int, [_P, _I, _I, _I, _I, _I, _P, C.c_size_t, _P, _P]),
}
_lib = None
_lock = threading.Lock()
_inited = None


class JpegError(RuntimeError):
    pass


def exported_symbols():
    return sorted(_SIGS)


def load() -> C.CDLL:
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        from . import build as _build
        try:
            _build.build_jpeg()
        except RuntimeError:
            if not os.path.exists(LIB_PATH):
                raise
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def _check(rc: int):
    if rc != 0:
        raise JpegError(f"libleafx_jpeg error {rc}: {(load().lfx_jpeg_last_error() or b'').decode()}")


def init(device: int = 0, backend: int = 2, threads: Optional[int] = None):
    """lfx_jpeg_init; backend 2 = nvJPEG GPU_HYBRID (Huffman decode on the GPU: 38 k 256x256 images/s on a B200 against 3 k
    with the default backend, profiles/r02_jpeg.json); `threads` encoder states (default: the CPU affinity, at most 16)."""
    global _inited
    lib = load()
    if threads is None:
        threads = max(1, min(16, len(os.sched_getaffinity(0))))
    key = (int(device), int(backend), int(threads))
    if _inited != key:
        _check(lib.lfx_jpeg_init(*key))
        _inited = key
    return lib


def _stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def read_files(paths: Sequence, workers: int = 8) -> List[Optional[bytes]]:
    """File bytes (None for unreadable files), read by `workers` threads."""
    def one(p):
        try:
            with open(p, "rb") as f:
                return f.read()
        except OSError:
            return None
    with ThreadPoolExecutor(max_workers=max(1, workers)) as io:
        return list(io.map(one, paths))


def probe(data: bytes) -> Tuple[int, int, int]:
    """(height, width, components) from the header."""
    lib = init(_current_device())
    w, h, nc, css = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
    _check(lib.lfx_jpeg_info(C.cast(buf, _P), len(data), C.byref(w), C.byref(h), C.byref(nc), C.byref(css)))
    return h.value, w.value, nc.value


def _current_device() -> int:
    import torch
    if not torch.cuda.is_available():
        raise JpegError("no CUDA device: the nvJPEG file boundary has no CPU fallback")
    return torch.cuda.current_device()


def decode_batch(blobs: Sequence[Optional[bytes]], H: int, W: int, out=None):
    """B bitstreams -> (uint8 [B,H,W,3] CUDA tensor, int32 status [B]); status < 0 = wrong size / undecodable (slot zeroed)."""
    import torch
    lib = init(_current_device())
    B = len(blobs)
    if out is None:
        out = torch.zeros((B, H, W, 3), dtype=torch.uint8, device="cuda")
    assert out.is_cuda and out.dtype == torch.uint8 and tuple(out.shape) == (B, H, W, 3) and out.is_contiguous()
    keep = [(C.c_uint8 * len(b)).from_buffer_copy(b) if b else None for b in blobs]
    ptrs = (C.c_void_p * B)(*[C.cast(k, _P) if k is not None else None for k in keep])
    lens = (C.c_size_t * B)(*[len(b) if b else 0 for b in blobs])
    status = np.zeros(B, np.int32)
    _check(lib.lfx_jpeg_decode_batch(C.cast(ptrs, _P), C.cast(lens, _P), C.c_void_p(out.data_ptr()), B, H, W,
                                     status.ctypes.data_as(_P), _stream_ptr()))
    torch.cuda.current_stream().synchronize()     # the bitstream buffers in `keep` may go now
    return out, status


def decode_files(paths: Sequence, H: Optional[int] = None, W: Optional[int] = None, workers: int = 8):
    """Folder boundary: paths -> (uint8 [B,H,W,3] CUDA tensor, status [B]).  H, W default to the first readable header."""
    blobs = read_files(paths, workers)
    if H is None or W is None:
        first = next((b for b in blobs if b), None)
        if first is None:
            raise JpegError("decode_files: no readable file")
        H, W, _ = probe(first)
    return decode_batch(blobs, H, W)


def encode_batch(batch, quality: int = 95, subsampling: int = 420) -> List[bytes]:
    """uint8 [B,H,W,3] CUDA tensor -> B baseline JPEG bitstreams (empty bytes for a failed image)."""
    import torch
    lib = init(_current_device())
    assert batch.is_cuda and batch.dtype == torch.uint8 and batch.dim() == 4 and batch.shape[3] == 3
    batch = batch.contiguous()
    B, H, W, _ = batch.shape
    cap = int(lib.lfx_jpeg_encode_bound(H, W, quality, subsampling))
    out = np.empty((B, cap), np.uint8)
    lens = np.zeros(B, np.uint64)
    _check(lib.lfx_jpeg_encode_batch(C.c_void_p(batch.data_ptr()), B, H, W, int(quality), int(subsampling), out.ctypes.data_as(_P),
                                     cap, lens.ctypes.data_as(_P), _stream_ptr()))
    return [out[i, :int(lens[i])].tobytes() for i in range(B)]


def encode_to_files(batch, paths: Sequence, quality: int = 95, subsampling: int = 420, workers: int = 8) -> List[bool]:
    """Encode on the GPU, write the bitstreams with `workers` threads; True per file written."""
    blobs = encode_batch(batch, quality, subsampling)

    def one(pb):
        p, b = pb
        if not b:
            return False
        try:
            p = Path(p)
            p.parent.mkdir(parents=True, exist_ok=True)
            with open(p, "wb") as f:
                f.write(b)
            return True
        except OSError:
            return False
    with ThreadPoolExecutor(max_workers=max(1, workers)) as io:
        return list(io.map(one, zip(paths, blobs)))
