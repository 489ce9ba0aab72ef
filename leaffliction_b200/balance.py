"""Class-balancing host logic: drop-in for srcs/cli/Distribution.py:count_images,
srcs/preprocessing/dataset_components.py (DistributionAnalyzer / AugmentationPlanner) and
srcs/preprocessing/dataset_balancer.py (DatasetBalancer).

What changes against the reference is WHERE the work runs, not WHAT is computed:
  * the per-task process pool (dataset_balancer.py:137-141, one OS process per output image) becomes one
    batched GPU submission per rank (`augment.augment_arrays`), tasks sharded by index across ranks;
  * the serial class count (Distribution.py:41-49) becomes per-rank partial counts merged by ONE allreduce
    (SUM, int64) over [num_classes | 9*256 colour histogram] -- NCCL over NVLink on GPUs, gloo in CPU tests;
  * plan and task list are recomputed identically on every rank from the same seeded `random` stream, so
    ranks agree without further communication.
Plan arithmetic, task order, RNG consumption, file naming and error conventions follow the reference.
"""
from __future__ import annotations

import logging
import os
import random
import shutil
from collections import defaultdict
from dataclasses import dataclass
from pathlib import Path
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

logger = logging.getLogger(__name__)

IMAGE_EXTS = {".jpg"}                                                        # Distribution.py:9, dataset_components.py:19
TRANSFORMATIONS = ["flip", "rotate", "skew", "shear", "crop", "distortion"]  # dataset_components.py:94


# --------------------------------------------------------------------------- counting (Distribution.py:21-49)
def is_image(path: Path) -> bool:
    return path.is_file() and path.suffix.lower() in IMAGE_EXTS


def iter_images(root: Path, plants: Optional[Iterable[str]] = None):
    """Yield (plant, class, path) for each image under root/PLANT/CLASS/*.jpg (Distribution.py:26-38)."""
    plant_filter = set(plants) if plants else None
    for plant_dir in sorted(d for d in Path(root).iterdir() if d.is_dir()):
        if plant_filter and plant_dir.name not in plant_filter:
            continue
        for class_dir in sorted(d for d in plant_dir.iterdir() if d.is_dir()):
            for img in class_dir.iterdir():
                if is_image(img):
                    yield plant_dir.name, class_dir.name, img


def count_images(root, plants: Optional[Iterable[str]] = None) -> List[Tuple[str, str, int]]:
    """Distribution.py:41-49 -- sorted [(plant, class, n)]."""
    counts: Dict[Tuple[str, str], int] = {}
    for plant, cls, _ in iter_images(Path(root), plants):
        counts[(plant, cls)] = counts.get((plant, cls), 0) + 1
    return sorted(((p, c, n) for (p, c), n in counts.items()), key=lambda x: (x[0], x[1]))


def analyze_dir(root) -> Dict[str, Dict[str, int]]:
    """DistributionAnalyzer._analyze_dir (dataset_components.py:29-43): counts[plant][class]."""
    root = Path(root)
    if not root.exists():
        raise FileNotFoundError(f"Dataset directory not found: {root}")
    counts: Dict[str, Dict[str, int]] = defaultdict(lambda: defaultdict(int))
    for plant_dir in (d for d in root.iterdir() if d.is_dir()):
        for class_dir in (c for c in plant_dir.iterdir() if c.is_dir()):
            n = sum(1 for f in class_dir.iterdir() if is_image(f))
            if n > 0:
                counts[plant_dir.name][class_dir.name] += n
    return {p: dict(c) for p, c in counts.items()}


# --------------------------------------------------------------------------- plan (dataset_components.py:79-109)
def calculate_plan(counts: Dict[str, Dict[str, int]]) -> Dict[str, Dict[str, int]]:
    """AugmentationPlanner.calculate_plan: per-plant maximum -> per-class deficit (keyed by class name only,
    reference quirk B.13) -> deficit // 6 per transform, +1 for the first deficit % 6 transforms."""
    deficits: Dict[str, int] = {}
    for _plant, classes in counts.items():
        plant_max = max(classes.values())
        for class_name, count in classes.items():
            deficit = plant_max - count
            if deficit > 0:
                deficits[class_name] = deficit
    plan: Dict[str, Dict[str, int]] = {}
    for class_name, deficit in deficits.items():
        plan[class_name] = {}
        base, rem = deficit // 6, deficit % 6
        for i, t in enumerate(TRANSFORMATIONS):
            n = base + (1 if i < rem else 0)
            if n > 0:
                plan[class_name][t] = n
    return plan


# --------------------------------------------------------------------------- tasks (dataset_balancer.py:105-129)
@dataclass
class Task:
    source_img: str
    output_path: str
    transform_name: str
    class_name: str
    seed: int
    source_index: int = -1     # index into the class' image list (in-memory datasets)


def build_tasks(plan: Dict[str, Dict[str, int]], images_by_class: Dict[str, Sequence], rng=random) -> List[Task]:
    """The reference's task list, drawn from the GLOBAL `random` stream in the reference's order:
    per class, per transform, per copy: random.choice(source_images) then random.randint(0, 1000000)."""
    tasks: List[Task] = []
    for class_name, transforms in plan.items():
        if class_name not in images_by_class:
            logger.warning(f"No images found for class '{class_name}'")
            continue
        source_images = images_by_class[class_name]
        for transform_name, count in transforms.items():
            for i in range(count):
                source_img = rng.choice(source_images)
                src = Path(str(source_img))
                new_name = src.stem + f"_aug_{transform_name}_{i + 1}" + src.suffix
                tasks.append(Task(str(source_img), str(src.parent / new_name), transform_name, class_name,
                                  rng.randint(0, 1000000)))
    return tasks


def shard(n_items: int, rank: int, world: int) -> range:
    """Image / task index i belongs to rank i % world."""
    return range(rank, n_items, world)


# --------------------------------------------------------------------------- the one collective
def allreduce_histograms(class_counts: np.ndarray, color_hist: Optional[np.ndarray] = None, device=None):
    """Merge per-rank partial class counts [num_classes] (+ optional colour histogram [9,256]) with one
    SUM allreduce of a single int64 buffer.  Returns (class_counts, color_hist) as NumPy int64.
    Without an initialised process group (single rank) the inputs are returned unchanged."""
    import torch
    import torch.distributed as dist
    cc = np.asarray(class_counts, np.int64).ravel()
    ch = None if color_hist is None else np.asarray(color_hist, np.int64).ravel()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return cc, (None if ch is None else ch.reshape(9, 256))
    buf = np.concatenate([cc, ch]) if ch is not None else cc
    t = torch.from_numpy(buf.copy())
    if dist.get_backend() == "nccl":
        t = t.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    out = t.cpu().numpy()
    return out[: cc.size], (None if ch is None else out[cc.size:].reshape(9, 256))


def distributed_counts(root, rank: int, world: int, plants=None) -> List[Tuple[str, str, int]]:
    """count_images with the file list sharded across ranks and merged by allreduce; equals count_images(root)."""
    items = [(p, c) for p, c, _ in iter_images(Path(root), plants)]
    keys = sorted(set(items))
    index = {k: i for i, k in enumerate(keys)}
    part = np.zeros(len(keys), np.int64)
    for i in shard(len(items), rank, world):
        part[index[items[i]]] += 1
    total, _ = allreduce_histograms(part)
    return [(p, c, int(n)) for (p, c), n in zip(keys, total) if n > 0]


# --------------------------------------------------------------------------- DatasetBalancer (dataset_balancer.py:19-198)
class DatasetBalancer:
    """Same constructor and `run()` as the reference.  `workers` is accepted for compatibility (JPEG I/O threads);
    the augmentations themselves run as GPU batches of `batch` tasks on this rank's shard."""

    def __init__(self, manifest_path=None, source_dir="images", target_dir="augmented_directory", seed=42, workers=None,
                 rank: int = 0, world: int = 1, batch: int = 256, gpu_jpeg: bool = False):
        from . import augment
        self.manifest_path = Path(manifest_path) if manifest_path else None
        self.source_dir = Path(source_dir)
        self.target_dir = Path(target_dir)
        self.transformer = augment.ImageAugmenter(seed=seed)   # seeds the GLOBAL RNGs like the reference (:31)
        # JPEG decode / encode threads (Pillow releases the GIL there); the reference's default is one process per core
        self.workers = max(1, int(workers)) if workers else max(1, min(8, os.cpu_count() or 1))
        self.rank, self.world, self.batch = int(rank), int(world), int(batch)
        # gpu_jpeg: sources are decoded and outputs encoded by nvJPEG (leaffliction_b200.jpegio), pixels stay in HBM between
        # the two; the default keeps the reference's Pillow codec, so that outputs match it bit for bit before encoding
        self.gpu_jpeg = bool(gpu_jpeg)
        self.counts: Dict[str, Dict[str, int]] = {}
        self.plan: Dict[str, Dict[str, int]] = {}
        self.completed = 0
        self.failed = 0

    def analyze_distribution(self):
        self.counts = analyze_dir(self.source_dir)
        for plant, classes in sorted(self.counts.items()):
            logger.info(f"\n[{plant}]")
            for class_name, count in sorted(classes.items()):
                logger.info(f"  {class_name}: {count} images")
        return self.counts

    def calculate_plan(self):
        self.plan = calculate_plan(self.counts)
        if not self.plan:
            logger.info("Dataset already balanced - no augmentations needed")
        return self.plan

    def _prepare_target_directory(self):
        """rank 0 recreates target_dir as a copy of source_dir (dataset_balancer.py:61-68); every rank learns whether that
        worked through an allreduce, so a failure on rank 0 raises everywhere instead of leaving the others in a barrier."""
        err = None
        if self.rank == 0:
            try:
                if self.target_dir.exists():
                    shutil.rmtree(self.target_dir)
                if not self.source_dir.exists():
                    raise FileNotFoundError(f"Source directory not found: {self.source_dir}")
                shutil.copytree(self.source_dir, self.target_dir)
            except Exception as e:   # noqa: BLE001 -- re-raised below, after the other ranks have been told
                err = e
        failed, _ = allreduce_histograms(np.array([1 if err is not None else 0], np.int64))
        if err is not None:
            raise err
        if int(failed[0]):
            raise RuntimeError("dataset balancing: rank 0 could not prepare the target directory")

    def _barrier(self):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.barrier()

    def _get_images_by_class(self):
        """Class name -> source images inside target_dir (dataset_balancer.py:85-93).  The listing is taken from
        SOURCE_DIR (target_dir is a fresh copy of it, so the names are the same) and rebased: ranks that are already
        writing `*_aug_*.jpg` files into target_dir cannot change what a slower rank lists.  Unlike the reference
        (filesystem order, quirk B.14) the lists are sorted so that every rank draws the same task list from the same
        seeded stream."""
        images_by_class = defaultdict(list)
        for plant_dir in self.source_dir.iterdir():
            if plant_dir.is_dir():
                for class_dir in plant_dir.iterdir():
                    if class_dir.is_dir():
                        files = list(class_dir.glob("*.JPG")) + list(class_dir.glob("*.jpg"))
                        images_by_class[class_dir.name] = [self.target_dir / f.relative_to(self.source_dir) for f in files]
        return {k: sorted(v) for k, v in images_by_class.items()}

    def execute_balancing(self):
        from . import augment
        if not self.plan:
            logger.info("No augmentation plan - skipping execution")
            return
        self._prepare_target_directory()
        tasks = build_tasks(self.plan, self._get_images_by_class())
        mine = [tasks[i] for i in shard(len(tasks), self.rank, self.world)]
        logger.info(f"Starting GPU augmentation: {len(tasks)} images to generate, {len(mine)} on rank {self.rank}")
        from concurrent.futures import ThreadPoolExecutor

        def try_load(t):
            try:
                return augment._load_rgb(t.source_img)
            except Exception as e:      # reference convention: log, count as failed, never raise
                logger.error(f"Failed to process {t.source_img} - {e}")
                return None

        def try_save(pair):
            t, o = pair
            try:
                augment._save_rgb(o, t.output_path)
                return True
            except Exception as e:
                logger.error(f"Failed: {t.output_path} - {e}")
                return False

        with ThreadPoolExecutor(max_workers=self.workers) as io:
            for b0 in range(0, len(mine), self.batch):
                chunk = mine[b0:b0 + self.batch]
                if self.gpu_jpeg:
                    chunk = self._run_chunk_on_device(chunk)     # returns the tasks it could not take (other sizes, bad files)
                    if not chunk:
                        continue
                loaded = list(io.map(try_load, chunk))
                good = [t for t, a in zip(chunk, loaded) if a is not None]
                imgs = [a for a in loaded if a is not None]
                try:
                    outs = augment.augment_arrays(imgs, [t.transform_name for t in good], [t.seed for t in good])
                except Exception as e:
                    logger.error(f"Batch failed - {e}")
                    self.failed += len(chunk)
                    continue
                ok = list(io.map(try_save, zip(good, outs)))
                self.completed += sum(ok)
                self.failed += len(chunk) - sum(ok)
        done, _ = allreduce_histograms(np.array([self.completed, self.failed], np.int64))
        logger.info(f"Augmentation complete: {int(done[0])} images generated, {int(done[1])} failed")
        self._barrier()

    def _run_chunk_on_device(self, chunk):
        """nvJPEG file boundary: the chunk's distinct sources are decoded into one device batch, the tasks run on it in place
        (augment.augment_device: kernels read source src_index[i]), every output is encoded from device memory.  Tasks whose
        source has another size than the chunk's first image, or does not decode, are returned for the host-codec path."""
        import dataclasses

        from . import augment, jpegio
        srcs = sorted({str(t.source_img) for t in chunk})
        index = {p: i for i, p in enumerate(srcs)}
        try:
            blobs = jpegio.read_files(srcs, self.workers)
            first = next((b for b in blobs if b), None)
            if first is None:
                return chunk
            H, W, _ = jpegio.probe(first)
            x, status = jpegio.decode_batch(blobs, H, W)
        except Exception as e:
            logger.error(f"nvJPEG decode failed, host codec takes this batch - {e}")
            return chunk
        good = [t for t in chunk if status[index[str(t.source_img)]] == 0]
        rest = [t for t in chunk if status[index[str(t.source_img)]] != 0]
        if not good:
            return rest
        try:
            out = augment.augment_device(x, [dataclasses.replace(t, source_index=index[str(t.source_img)]) for t in good])
            written = 0
            for key, val in out.items():
                if key == "rotate":
                    ids, slab, hw = val
                    for k, i in enumerate(ids):
                        nh, nw = int(hw[k][0]), int(hw[k][1])
                        img = slab[k, :nh * nw * 3].view(1, nh, nw, 3)
                        written += sum(jpegio.encode_to_files(img, [good[i].output_path], 95, 420, 1))
                else:
                    ids, batch = val
                    written += sum(jpegio.encode_to_files(batch, [good[i].output_path for i in ids], 95, 420, self.workers))
            self.completed += written
            self.failed += len(good) - written
        except Exception as e:
            logger.error(f"Batch failed - {e}")
            self.failed += len(good)
        return rest

    def run(self):
        logger.info("=== Dataset Balancing System ===")
        try:
            self.analyze_distribution()
            self.calculate_plan()
            self.execute_balancing()
            logger.info("=== Balancing Complete ===")
        except Exception as e:
            logger.error(f"Dataset balancing failed - {e}")
            raise


# --------------------------------------------------------------------------- in-memory datasets (BASELINE config 3)
def synthetic_class_counts() -> Dict[str, Dict[str, int]]:
    """64 Ki images, 2 plants x 4 classes, imbalanced."""
    return {"PlantA": {"A_c0": 16384, "A_c1": 12288, "A_c2": 8192, "A_c3": 6144},
            "PlantB": {"B_c0": 9216, "B_c1": 6144, "B_c2": 3584, "B_c3": 3584}}


def tasks_for_labels(labels: np.ndarray, class_names: Sequence[str], plants: Dict[str, Sequence[str]], seed: int = 42):
    """Balance plan + task list for an in-memory dataset (labels[i] = class id of image i).
    Returns (plan, tasks) with Task.source_index = dataset index of the source image; identical on every rank."""
    labels = np.asarray(labels)
    per_class = np.bincount(labels, minlength=len(class_names))
    counts = {p: {c: int(per_class[class_names.index(c)]) for c in cls if per_class[class_names.index(c)] > 0} for p, cls in plants.items()}
    plan = calculate_plan(counts)
    rng = random.Random(seed)
    by_class = {c: np.nonzero(labels == class_names.index(c))[0] for c in class_names}
    tasks: List[Task] = []
    for class_name, transforms in plan.items():
        src = by_class[class_name]
        for t, n in transforms.items():
            for i in range(n):
                j = int(src[rng.randrange(len(src))])
                tasks.append(Task(f"{class_name}/{j}.jpg", f"{class_name}/{j}_aug_{t}_{i + 1}.jpg", t, class_name,
                                  rng.randint(0, 1000000), j))
    return plan, tasks


def task_arrays_for_labels(labels: np.ndarray, class_names: Sequence[str], plants: Dict[str, Sequence[str]], seed: int = 42):
    """tasks_for_labels without the Python objects: (plan, augment.TaskArrays) with the same draws, made by the native
    restatement of the interpreter's stream (lfx_draw_balance_tasks).  Identical on every rank."""
    import ctypes as C

    from . import _lib, augment
    labels = np.asarray(labels)
    per_class = np.bincount(labels, minlength=len(class_names))
    counts = {p: {c: int(per_class[class_names.index(c)]) for c in cls if per_class[class_names.index(c)] > 0} for p, cls in plants.items()}
    plan = calculate_plan(counts)
    groups = [(class_names.index(cn), augment.TRANSFORM_CODE[t], n) for cn, tr in plan.items() for t, n in tr.items()]
    gcount = np.array([g[2] for g in groups], np.int32)
    gsize = np.array([per_class[g[0]] for g in groups], np.int32)
    total = int(gcount.sum())
    local = np.zeros(total, np.int32)
    tseed = np.zeros(total, np.int32)
    P = C.c_void_p
    _lib.check(_lib.load().lfx_draw_balance_tasks(int(seed) & 0xFFFFFFFF, len(groups), gcount.ctypes.data_as(P), gsize.ctypes.data_as(P),
                                                  local.ctypes.data_as(P), tseed.ctypes.data_as(P)))
    order = np.argsort(labels, kind="stable")                    # dataset indices grouped by class, ascending inside a class
    starts = np.concatenate([[0], np.cumsum(per_class)])[:-1]
    gclass = np.repeat(np.array([g[0] for g in groups], np.int64), gcount)
    ta = object.__new__(augment.TaskArrays)
    ta.transform = np.repeat(np.array([g[1] for g in groups], np.int32), gcount)
    ta.seed = tseed.astype(np.int64)
    ta.source_index = order[starts[gclass] + local].astype(np.int64)
    return plan, ta
