"""Drop-in for srcs/cli/Transformation.py: same flags (:568-608), same type aliases (`--types`, :46-60,393-408),
same output names `{stem}__T_{Type}.jpg` (:411-420), same skip/overwrite rule (:461).  The arithmetic runs in
libleafx (leaffliction_b200.transform / .filters).  Differences, all outside the numeric hot path (SURVEY.md
section 8): `Hist` writes its statistics as JSON next to where the matplotlib figure would go, `Landmarks` is
skipped with a warning, no mosaic is drawn; folder mode batches `make_mask` per image shape on the GPU instead of
forking an `mp.Pool` (`--workers` is accepted)."""
from __future__ import annotations

import argparse
import json
import logging
import os
import re
from dataclasses import dataclass
from pathlib import Path
from typing import List, Sequence, Tuple

import numpy as np

from leaffliction_b200 import transform as T

DEFAULT_CONFIG = "srcs/transform/config.yaml"                      # the reference's default (:587)
PACKAGED_CONFIG = Path(T.__file__).with_name("config.yaml")           # same keys and values, shipped with this package


@dataclass
class ProcessArgs:
    img_path: Path
    out_dir: Path
    types: Tuple[str, ...]
    cfg: T.TransformConfig
    skip_existing: bool = False
    overwrite: bool = False


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="Image transformation pipeline (CUDA).\n- Single image: Transformation.py path/to/image.jpg\n"
                                            "- Folder mode: Transformation.py -src DIR -dst OUTDIR [--workers N]")
    p.add_argument("image", nargs="?", help="Path to a single image for preview mode")
    p.add_argument("--out-dir", default=None, help="Output directory for single image preview")
    p.add_argument("-src", "--src", default=None, help="Source directory (folder mode)")
    p.add_argument("-dst", "--dst", default=None, help="Destination directory (folder mode)")
    p.add_argument("--types", default=",".join(T.DEFAULT_TYPES), help="Comma-separated transforms to run")
    p.add_argument("--config", default=DEFAULT_CONFIG, help="YAML config path (optional)")
    p.add_argument("--workers", type=int, default=0, help="JPEG I/O threads (0=auto); the GPU batches themselves ignore it")
    p.add_argument("--skip-existing", action="store_true", help="Skip images whose outputs already exist")
    p.add_argument("--overwrite", action="store_true", help="Overwrite existing outputs")
    p.add_argument("--preview", action="store_true", help="Force saving outputs (no GUI popups)")
    return p


def parse_args(argv=None) -> argparse.Namespace:
    return build_parser().parse_args(argv)


def _want(params: ProcessArgs, out: Path) -> bool:
    return params.overwrite or (not params.skip_existing or not out.exists())   # Transformation.py:461


def process_single_image(params: ProcessArgs, premade=None, rgb=None, sink=None) -> List[Path]:
    """Transformation.py:423-536.  `premade` = (mask, contour) from a batched make_mask (folder mode); `rgb` = the already
    decoded image; `sink`: a list that receives (path, image) instead of the file being written here (folder mode with
    LEAFX_GPU_JPEG=1 encodes them as one batch)."""
    if rgb is None:
        try:
            rgb = T.pil_read_rgb(params.img_path)
        except Exception as exc:
            logging.error("Failed to read %s (%s)", params.img_path, exc)
            return []
    pipe = T.TransformPipeline(params.cfg)
    names = T.output_names(params.img_path.stem)
    saved: List[Path] = []
    mask_img, contour, masked_rgb = None, None, rgb
    if any(t in params.types for t in ("Mask", "ROI", "Analyze", "Landmarks", "Brown", "Blur")):
        mask_img, contour = premade if premade is not None else pipe.make_mask(rgb)
        if mask_img is not None:
            masked_rgb = pipe.create_masked_rgb(rgb, mask_img)

    def save(kind, img):
        out = params.out_dir / names[kind]
        if _want(params, out) and img is not None:
            if sink is not None:
                sink.append((out, np.ascontiguousarray(img)))
            else:
                T.imwrite_rgb(out, img)
            saved.append(out)

    if "Mask" in params.types:
        save("Mask", T.apply_mask(rgb, mask_img, "black") if mask_img is not None else rgb)   # mask.py:585-607
    if "Blur" in params.types:
        save("Blur", pipe.blur(masked_rgb))
    if "ROI" in params.types:
        _, roi_vis, _ = pipe.roi(masked_rgb, contour)
        save("ROI", roi_vis if roi_vis is not None else masked_rgb)                          # :480-484 saves the rectangle view
    if "Analyze" in params.types:
        save("Analyze", pipe.analyze(masked_rgb, mask_img, contour))
    if "Landmarks" in params.types:
        logging.warning("Landmarks: out of scope for the CUDA hot path (SURVEY.md 8f), skipped for %s", params.img_path.name)
    if "Hist" in params.types:
        out = params.out_dir / (names["Hist"][:-4] + ".json")
        if _want(params, out):
            st = pipe.histogram_hsv(masked_rgb)
            out.write_text(json.dumps({"total_pixels": st["total_pixels"], "color_analysis": st["color_analysis"],
                                       "hue_ranges": st["hue_ranges"], "hsv_hist": np.asarray(st["hsv_hist"]).tolist()}))
            saved.append(out)
    if "Brown" in params.types:
        brown_img, _, _ = pipe.detect_brown_spots(masked_rgb, mask_img)
        save("Brown", brown_img)
    return saved


def _decode_chunk_nvjpeg(chunk, nthreads):
    """Sources of one chunk through nvJPEG, grouped by header size; anything it refuses falls back to Pillow."""
    from leaffliction_b200 import jpegio
    blobs = jpegio.read_files(chunk, nthreads)
    arrays = [None] * len(chunk)
    groups = {}
    for i, b in enumerate(blobs):
        try:
            h, w, _ = jpegio.probe(b) if b else (0, 0, 0)
        except Exception:
            h = w = 0
        if h and w:
            groups.setdefault((h, w), []).append(i)
    for (h, w), ids in groups.items():
        x, status = jpegio.decode_batch([blobs[i] for i in ids], h, w)
        host = x.cpu().numpy()
        for k, i in enumerate(ids):
            if status[k] == 0:
                arrays[i] = host[k]
    for i, a in enumerate(arrays):
        if a is None:
            try:
                arrays[i] = T.pil_read_rgb(chunk[i])
            except Exception:
                pass
    return arrays


def _encode_sink_nvjpeg(sink, nthreads):
    """(path, image) pairs -> files, one nvJPEG batch per image shape (grey images are replicated to RGB)."""
    import torch

    from leaffliction_b200 import jpegio
    groups = {}
    for k, (_, img) in enumerate(sink):
        if img.ndim == 2:
            sink[k] = (sink[k][0], np.repeat(img[:, :, None], 3, axis=2))
        groups.setdefault(sink[k][1].shape, []).append(k)
    for ids in groups.values():
        batch = torch.from_numpy(np.stack([sink[k][1] for k in ids])).cuda()
        ok = jpegio.encode_to_files(batch, [sink[k][0] for k in ids], 95, 420, nthreads)
        for k, good in zip(ids, ok):
            if not good:
                T.imwrite_rgb(sink[k][0], sink[k][1])


def run_folder(src: Path, dst: Path, types: Sequence[str], cfg, skip_existing: bool, overwrite: bool, batch: int = 256,
               workers: int = 0, gpu_jpeg: bool = False):
    """Folder mode (:664-699): images grouped by shape, make_mask batched per group on the GPU; `workers` threads
    decode the JPEGs (0 = auto = min(8, cpu // 2), the reference's rule :672-674).  `gpu_jpeg`: the codec runs on the GPU
    (nvJPEG) for sources and outputs."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    nthreads = workers if workers and workers > 0 else max(1, min(8, (os.cpu_count() or 2) // 2))

    def try_read(ip):
        try:
            return T.pil_read_rgb(ip)
        except Exception:
            return None
    imgs = list(T.iter_images_in_dir(src))
    if not imgs:
        logging.warning("No images found in %s", src)
        return 0
    logging.info("Found %d images in %s", len(imgs), src)
    total = 0
    for b0 in range(0, len(imgs), batch):
        chunk = imgs[b0:b0 + batch]
        if gpu_jpeg:
            arrays = _decode_chunk_nvjpeg(chunk, nthreads)
        else:
            with ThreadPoolExecutor(max_workers=nthreads) as io:
                arrays = list(io.map(try_read, chunk))
        sink = [] if gpu_jpeg else None
        premade = [None] * len(chunk)
        by_shape = {}
        for i, a in enumerate(arrays):
            if a is not None:
                by_shape.setdefault(a.shape, []).append(i)
        for ids in by_shape.values():
            try:
                masks, _, contours = T.make_mask_batch(np.stack([arrays[i] for i in ids]), cfg)
            except Exception as e:      # one shape group that cannot be masked must not end the run (the reference's
                logging.error("Failed to mask %d image(s) of shape %s - %s", len(ids), arrays[ids[0]].shape, e)   # worker logs and goes on)
                continue
            for k, i in enumerate(ids):
                premade[i] = (masks[k], contours[k])
        for ip, a, pm in zip(chunk, arrays, premade):
            if a is not None and pm is None:
                continue                # its group failed above
            try:
                total += len(process_single_image(ProcessArgs(ip, dst, tuple(types), cfg, skip_existing, overwrite), pm, rgb=a, sink=sink))
            except Exception as e:      # Transformation.py:700-705: a failing image is logged, the folder run continues
                logging.error("Failed to process %s - %s", ip, e)
        if sink:
            _encode_sink_nvjpeg(sink, nthreads)
    logging.info("Processed %d images, saved %d outputs", len(imgs), total)
    return total


def main(argv=None) -> None:
    args = parse_args(argv)
    logging.basicConfig(level=logging.INFO, format="%(asctime)s | %(levelname)s | %(message)s")
    types = T.build_types_filter(args.types)
    cfg_path = Path(args.config) if args.config else None
    if cfg_path is not None and not cfg_path.exists() and args.config == DEFAULT_CONFIG:
        cfg_path = PACKAGED_CONFIG                                      # not inside a reference checkout
    cfg = T.load_config(cfg_path)
    if args.image and not args.src and not args.dst:
        ip = Path(args.image)
        if not T.is_image(ip):
            logging.error("Not a valid image: %s", ip)
            return
        m = re.search(r"image \((\d+)\)", ip.stem)
        image_number = m.group(1) if m else ip.stem
        out_d = Path(args.out_dir) if args.out_dir else Path("artifacts") / "transformations" / image_number
        out_d.mkdir(parents=True, exist_ok=True)
        saved = process_single_image(ProcessArgs(ip, out_d, types, cfg, args.skip_existing, args.overwrite))
        print(f"Saved {len(saved)} outputs to {out_d}")
        for s in saved:
            print(f"  - {s}")
        return
    if args.src and args.dst:
        src, dst = Path(args.src), Path(args.dst)
        if not src.exists():
            logging.error("Source directory does not exist: %s", src)
            return
        dst.mkdir(parents=True, exist_ok=True)
        # LEAFX_GPU_JPEG=1: nvJPEG decode / encode in folder mode (an environment switch: the flag set stays the reference's)
        run_folder(src, dst, types, cfg, args.skip_existing, args.overwrite, workers=args.workers,
                   gpu_jpeg=os.environ.get("LEAFX_GPU_JPEG", "0") == "1")
        return
    logging.error("Must specify either single image or --src/--dst for folder mode")


if __name__ == "__main__":
    main()
