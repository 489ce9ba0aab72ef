#!/usr/bin/env python3
"""Throughput of the nvJPEG file boundary (leaffliction_b200.jpegio) against the reference's host codec (Pillow, as
srcs/utils/image_utils.py:19-59 uses it) on this box's cores.  Bitstreams in host memory, pixels in HBM.

  python tools/bench_jpeg.py [--batch 4096] [--size 256] [--backend 0|1|2|3] > profiles/r02_jpeg.json
"""
import argparse
import io
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--backend", type=int, default=2)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    import torch
    from PIL import Image

    from leaffliction_b200 import jpegio, synth
    cores = len(os.sched_getaffinity(0))
    base = synth.leaf_batch(64, a.size, a.size)
    imgs = np.concatenate([base] * ((a.batch + 63) // 64))[:a.batch]

    def pil_enc(arr):
        buf = io.BytesIO()
        Image.fromarray(arr).save(buf, format="JPEG", quality=95)
        return buf.getvalue()

    def pil_dec(b):
        return np.asarray(Image.open(io.BytesIO(b)).convert("RGB"))
    nhost = min(a.batch, 1024)
    with ThreadPoolExecutor(max_workers=cores) as ex:
        t0 = time.perf_counter()
        blobs_h = list(ex.map(pil_enc, imgs[:nhost]))
        t_pil_enc = time.perf_counter() - t0
        t0 = time.perf_counter()
        list(ex.map(pil_dec, blobs_h))
        t_pil_dec = time.perf_counter() - t0
        blobs = blobs_h + list(ex.map(pil_enc, imgs[nhost:]))
    jpegio.init(0, a.backend)
    x = torch.from_numpy(imgs).cuda()
    out = torch.empty_like(x)
    res = {"batch": a.batch, "size": a.size, "host_cores": cores, "backend_requested": a.backend,
           "backend_used": int(jpegio.load().lfx_jpeg_backend()), "mean_stream_bytes": float(np.mean([len(b) for b in blobs]))}
    jpegio.decode_batch(blobs, a.size, a.size, out=out)          # warm-up (allocations inside nvJPEG)
    ts = []
    for _ in range(a.reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        jpegio.decode_batch(blobs, a.size, a.size, out=out)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    res["decode_images_per_s"] = a.batch / min(ts)
    jpegio.encode_batch(x[:256])
    ts = []
    for _ in range(a.reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        enc = jpegio.encode_batch(x)
        ts.append(time.perf_counter() - t0)
    res["encode_images_per_s"] = a.batch / min(ts)
    res["encode_mean_stream_bytes"] = float(np.mean([len(b) for b in enc]))
    res["pillow_decode_images_per_s"] = nhost / t_pil_dec
    res["pillow_encode_images_per_s"] = nhost / t_pil_enc
    res["pillow_note"] = f"{nhost} images, {cores} threads (Pillow releases the GIL inside the codec)"
    res["pcie_bytes_per_image"] = {"raw_rgb": a.size * a.size * 3, "jpeg_q95": res["mean_stream_bytes"]}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
