"""CPU tests of the class-balancing host logic (SURVEY.md section 8 row a16 and row e): plan arithmetic against
the reference's AugmentationPlanner (golden fixture), counting, the task list's RNG order, index sharding and
the world_size-2 gloo allreduce that merges class / colour histograms."""
import os
import random
import socket
import sys

import numpy as np
import pytest

from leaffliction_b200 import balance

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "golden_v1.npz"))


def test_plan_matches_reference_planner():
    plan = balance.calculate_plan(balance.synthetic_class_counts())
    classes = [str(c) for c in G["plan/classes"]]
    assert sorted(plan) == classes
    for cls, row in zip(classes, G["plan/counts"]):
        assert [plan[cls].get(t, 0) for t in balance.TRANSFORMATIONS] == row.tolist()
    # 36,864 augment tasks in total
    assert sum(sum(v.values()) for v in plan.values()) == 36864


def test_plan_properties():
    rng = np.random.default_rng(0)
    for _ in range(50):
        counts = {f"P{p}": {f"P{p}_c{c}": int(rng.integers(1, 500)) for c in range(int(rng.integers(1, 6)))} for p in range(3)}
        plan = balance.calculate_plan(counts)
        for plant, classes in counts.items():
            mx = max(classes.values())
            for cls, n in classes.items():
                assert sum(plan.get(cls, {}).values()) == mx - n           # plan sums to the deficit
                per = list(plan.get(cls, {}).values())
                assert not per or max(per) - min(per) <= 1                   # spread evenly over the transforms


def _make_tree(tmp_path, spec):
    for plant, classes in spec.items():
        for cls, n in classes.items():
            d = tmp_path / plant / cls
            d.mkdir(parents=True)
            for i in range(n):
                (d / f"img_{i:03d}.JPG").write_bytes(b"x")
            (d / "notes.txt").write_text("ignored")                           # only .jpg counts (quirk B.10)
            (d / "scan.png").write_bytes(b"x")
    return tmp_path


def test_count_images_and_analyze_dir(tmp_path):
    spec = {"Apple": {"Apple_healthy": 7, "Apple_scab": 3}, "Grape": {"Grape_spot": 5}}
    root = _make_tree(tmp_path, spec)
    assert balance.count_images(root) == [("Apple", "Apple_healthy", 7), ("Apple", "Apple_scab", 3), ("Grape", "Grape_spot", 5)]
    assert balance.count_images(root, ["Grape"]) == [("Grape", "Grape_spot", 5)]
    assert balance.analyze_dir(root) == spec


@pytest.mark.needs_reference
def test_count_and_plan_against_reference(tmp_path):
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import ref_harness
    ns = ref_harness.load()
    import importlib
    dist_mod = importlib.import_module("srcs.cli.Distribution")
    spec = {"Apple": {"Apple_healthy": 9, "Apple_scab": 2, "Apple_rust": 4}, "Grape": {"Grape_spot": 5, "Grape_esca": 6}}
    root = _make_tree(tmp_path, spec)
    assert balance.count_images(root) == dist_mod.count_images(root, None)
    an = ns.components.DistributionAnalyzer(root)
    counts = {p: dict(c) for p, c in an.analyze().items()}
    assert balance.analyze_dir(root) == counts
    assert balance.calculate_plan(counts) == ns.components.AugmentationPlanner(counts).calculate_plan()


def test_task_list_follows_reference_rng_order():
    """dataset_balancer.py:105-129: per class, per transform, per copy: random.choice then random.randint."""
    plan = {"c1": {"flip": 2, "rotate": 1}, "c2": {"crop": 2}}
    imgs = {"c1": ["/d/c1/a.JPG", "/d/c1/b.JPG", "/d/c1/c.JPG"], "c2": ["/d/c2/x.jpg", "/d/c2/y.jpg"]}
    random.seed(42)
    tasks = balance.build_tasks(plan, imgs)
    random.seed(42)
    exp = []
    for cls, tr in plan.items():
        for t, n in tr.items():
            for i in range(n):
                src = random.choice(imgs[cls])
                exp.append((src, t, i + 1, random.randint(0, 1000000)))
    assert [(t.source_img, t.transform_name, int(t.output_path.rsplit("_", 1)[1].split(".")[0]), t.seed) for t in tasks] == exp
    assert tasks[0].output_path.endswith(f"_aug_flip_1.JPG")


def test_shard_partition():
    for n in (0, 1, 7, 64, 36864):
        for world in (1, 2, 8):
            seen = sorted(i for r in range(world) for i in balance.shard(n, r, world))
            assert seen == list(range(n))
            sizes = [len(balance.shard(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_tasks_for_labels_deterministic():
    counts = balance.synthetic_class_counts()
    names = [c for p in counts.values() for c in p]
    plants = {p: list(c) for p, c in counts.items()}
    labels = np.repeat(np.arange(len(names)), [max(1, n // 64) for p in counts.values() for n in p.values()])
    plan1, t1 = balance.tasks_for_labels(labels, names, plants, seed=42)
    plan2, t2 = balance.tasks_for_labels(labels, names, plants, seed=42)
    assert plan1 == plan2 and [(t.source_index, t.seed, t.transform_name) for t in t1] == [(t.source_index, t.seed, t.transform_name) for t in t2]
    for t in t1:
        assert names[labels[t.source_index]] == t.class_name


# ----------------------------------------------------------------------------- world_size 2, gloo
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, root, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        counts = balance.distributed_counts(root, rank, world)
        rng = np.random.default_rng(100 + rank)
        part_cls = rng.integers(0, 1000, 8)
        part_hist = rng.integers(0, 5000, (9, 256))
        cc, ch = balance.allreduce_histograms(part_cls, part_hist)
        q.put((rank, counts, cc.tolist(), int(ch.sum()), ch[3, :5].tolist()))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_histogram_allreduce(tmp_path):
    import torch.multiprocessing as mp
    spec = {"Apple": {"Apple_healthy": 11, "Apple_scab": 4}, "Grape": {"Grape_spot": 6}}
    root = str(_make_tree(tmp_path, spec))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, root, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    serial = balance.count_images(root)
    exp_cls = sum(np.random.default_rng(100 + r).integers(0, 1000, 8) for r in range(2))
    for rank, counts, cc, hsum, hrow in res:
        assert counts == serial                          # merged class histogram == serial count_images
        assert cc == exp_cls.tolist()
    assert res[0][3:] == res[1][3:]


def _prep_worker(rank, world, port, src, dst, q):
    """Both ranks call the balancer's directory preparation; rank 0's source directory does not exist."""
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b = balance.DatasetBalancer(source_dir=src, target_dir=dst, seed=42, rank=rank, world=world)
        try:
            b._prepare_target_directory()
            q.put((rank, "ok", sorted(k for k in b._get_images_by_class())))
        except Exception as e:   # noqa: BLE001
            q.put((rank, type(e).__name__, str(e)[:60]))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_prepare_failure_reaches_every_rank(tmp_path):
    """a rank-0 failure while preparing the target directory used to leave the other ranks in a barrier.
    Now every rank raises (rank 0 its own error, the others a RuntimeError) and nobody hangs; with a valid source both
    ranks list the same classes, from source_dir (immune to files other ranks are already writing into target_dir)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    for src, expect_ok in ((str(tmp_path / "missing"), False), (str(_make_tree(tmp_path / "tree", {"Apple": {"Apple_a": 3, "Apple_b": 1}})), True)):
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_prep_worker, args=(r, 2, port, src, str(tmp_path / ("out_ok" if expect_ok else "out_bad")), q)) for r in range(2)]
        for p in procs:
            p.start()
        res = sorted(q.get(timeout=120) for _ in procs)
        for p in procs:
            p.join(60)
            assert p.exitcode == 0
        if expect_ok:
            assert [r[1] for r in res] == ["ok", "ok"] and res[0][2] == res[1][2] == ["Apple_a", "Apple_b"]
        else:
            assert res[0][1] == "FileNotFoundError" and res[1][1] == "RuntimeError"
