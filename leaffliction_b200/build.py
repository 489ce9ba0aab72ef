"""Build libleafx.so in-tree with nvcc for sm_100a (no other arch, no fallback)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libleafx.so")
OUT_JPEG = os.path.join(HERE, "libleafx_jpeg.so")     # nvJPEG file boundary (include/leafx_jpeg.h), a separate object
SOURCES = ["lfx_api.cu", "lfx_color.cu", "lfx_augment.cu", "lfx_gauss.cu", "lfx_mask.cu", "lfx_roi.cu", "lfx_contour.cu", "lfx_front.cu", "lfx_core.cu", "lfx_gauss_tma.cu", "lfx_rng.cu", "lfx_params.cu", "lfx_resize.cu", "lfx_score.cu", "lfx_kmeans.cu", "lfx_draw.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC,-ffp-contract=off", "--use_fast_math=false"]


def _nvcc() -> str:
    nv = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nv):
        raise RuntimeError("nvcc not found: libleafx cannot be built (there is no CPU fallback)")
    return nv


def _stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f != "lfx_jpeg.cu"] + [os.path.join(HERE, "..", "include", "leafx.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_jpeg(force: bool = False) -> str:
    """libleafx_jpeg.so = csrc/lfx_jpeg.cu + libnvjpeg (CUDA toolkit)."""
    src = os.path.join(CSRC, "lfx_jpeg.cu")
    hdr = os.path.join(HERE, "..", "include", "leafx_jpeg.h")
    if not force and os.path.exists(OUT_JPEG) and os.path.getmtime(OUT_JPEG) >= max(os.path.getmtime(src), os.path.getmtime(hdr)):
        return OUT_JPEG
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
           src, "-o", OUT_JPEG, "-lnvjpeg", "-lcudart", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for lfx_jpeg.cu:\n{r.stdout}\n{r.stderr}")
    return OUT_JPEG


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return OUT
    nv = _nvcc()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]
    if verbose:
        flags += ["-Xptxas", "-v"]

    def one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nv, *flags, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(one, SOURCES))
    cmd = [nv, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT, *objs, "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_jpeg(force="--force" in sys.argv))
