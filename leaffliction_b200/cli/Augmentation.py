"""Drop-in for srcs/cli/Augmentation.py: same positional argument and flags (:32-78), same modes, same
output names (`original_<name>`, `<transform>_<name>` :126,142) and exit codes (any failure -> exit 1, :103-113).
The six augmentations run on the GPU (leaffliction_b200.augment); dataset mode balances with
leaffliction_b200.balance.DatasetBalancer (one batched submission per rank instead of a process pool).
Under torchrun (WORLD_SIZE > 1) the augment tasks shard across ranks and the class histogram is merged by
one NCCL allreduce."""
from __future__ import annotations

import argparse
import logging
import os
import shutil
import sys
from pathlib import Path

logger = logging.getLogger("leaffliction_b200.cli.Augmentation")

SUPPORTED_IMAGE_EXTENSIONS = {".jpg", ".jpeg", ".png", ".bmp", ".tiff"}
DEFAULT_DATASET_OUTPUT = "artifacts/augmented_directory"
DEFAULT_SINGLE_OUTPUT = "artifacts/example"
DEFAULT_SEED = 42
TRANSFORMATIONS = ["flip", "rotate", "skew", "shear", "crop", "distortion"]


class AugmentationError(Exception):
    pass


class InputValidationError(AugmentationError):
    pass


class ProcessingError(AugmentationError):
    pass


def build_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(
        description="Apply augmentations to balance a dataset. Preferred usage: provide a dataset root (PLANT/CLASS/*.jpg).",
        formatter_class=argparse.RawDescriptionHelpFormatter)
    parser.add_argument("input_path", help="Path to dataset root directory (preferred) OR single image file.")
    parser.add_argument("-out", "--output", help="Output directory (default: artifacts/augmented_directory for datasets, "
                                                   "artifacts/example for single images)")
    parser.add_argument("-seed", "--seed", type=int, default=DEFAULT_SEED, help="Random seed for reproducible results")
    parser.add_argument("--workers", type=int, default=None, help="JPEG I/O threads (default min(8, cores)); the GPU batches themselves ignore it")
    return parser


def parse_args(argv=None):
    return build_parser().parse_args(argv)


def single_image_mode(args, image_path: Path):
    from leaffliction_b200.augment import ImageAugmenter
    output_dir = Path(args.output) if args.output else Path(DEFAULT_SINGLE_OUTPUT)
    output_dir.mkdir(parents=True, exist_ok=True)
    logger.info(f"Processing single image: {image_path}")
    original_output = output_dir / f"original_{image_path.name}"
    shutil.copy2(image_path, original_output)
    augmenter = ImageAugmenter(seed=args.seed)
    for transform in TRANSFORMATIONS:
        output_path = output_dir / f"{transform}_{image_path.name}"
        if getattr(augmenter, transform)(str(image_path), str(output_path)):
            logger.info(f"{transform.capitalize()} applied: {output_path}")
        else:
            raise ProcessingError(f"Failed to apply {transform} transformation")
    logger.info("Single image augmentation completed successfully")


def dataset_mode_dir(args, source_dir: Path):
    from leaffliction_b200.balance import DatasetBalancer, count_images
    target_dir = Path(args.output) if args.output else Path(DEFAULT_DATASET_OUTPUT)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    balancer = DatasetBalancer(source_dir=str(source_dir), target_dir=str(target_dir), seed=args.seed, workers=args.workers,
                               rank=rank, world=world, gpu_jpeg=os.environ.get("LEAFX_GPU_JPEG", "0") == "1")
    balancer.run()
    if rank == 0:
        rows = count_images(target_dir, None)
        logger.info("Total balanced images: %d", sum(n for _, _, n in rows))
        for plant, cls, n in rows:
            logger.info("  %s / %s: %d", plant, cls, n)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def main(argv=None):
    logging.basicConfig(level=logging.INFO, format="%(asctime)s | %(levelname)s | %(message)s")
    try:
        args = parse_args(argv)
        input_path = Path(args.input_path)
        if not input_path.exists():
            raise InputValidationError(f"Input path not found: {input_path}")
        if input_path.is_file() and input_path.suffix.lower() in SUPPORTED_IMAGE_EXTENSIONS:
            single_image_mode(args, input_path)
            return
        if input_path.is_dir():
            dataset_mode_dir(args, input_path)
            return
        raise InputValidationError("Unsupported input. Provide a dataset directory or an image file.")
    except InputValidationError as e:
        logger.error(f"Input validation error: {e}")
        sys.exit(1)
    except ProcessingError as e:
        logger.error(f"Processing error: {e}")
        sys.exit(1)
    except Exception as e:
        logger.error(f"Unexpected error: {e}")
        sys.exit(1)


if __name__ == "__main__":
    main()
