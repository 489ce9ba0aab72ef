"""lfx_draw_augment_params (host-side, no GPU): the native restatement of CPython's `random` stream + PIL rotate
geometry must reproduce, bit for bit, what the interpreter draws for a fresh ImageAugmenter(seed)
(image_augmenter.py:16-18 and the draw order of each method; dataset_balancer.py:201-207)."""
import math
import random

import numpy as np
import pytest

from leaffliction_b200 import augment


def _python_params(name, seed, h, w):
    p = augment.draw_task_params(name, seed, h, w, want_noise=False)
    ip, dp = np.zeros(8, np.int32), np.zeros(8, np.float64)
    if name == "flip":
        ip[0] = 0 if p[0] else 1
    elif name == "rotate":
        m, nw, nh = augment.rotate_matrix(p[0], w, h)
        if isinstance(m, str):
            m, nw, nh = [1.0, 0.0, 0.0, 0.0, 1.0, 0.0], w, h
        ip[:] = augment.fixed_affine(m) + [nw, nh]
        dp[0] = p[0]
    elif name in ("skew", "shear"):
        dp[:] = p[0]
        ip[0] = int(p[1])
    elif name == "crop":
        ip[:4] = p
    else:
        ip[0] = int(h * w * p[1] // 100)
        dp[0] = p[1]
    return ip, dp


@pytest.mark.parametrize("h,w", [(256, 256), (96, 131), (1024, 768)])
def test_native_params_equal_interpreter(h, w):
    rng = random.Random(h * 7 + w)
    seeds = [rng.randint(1, 1_000_000) for _ in range(1500)] + [1, 2, 42, 7, 999983, 1_000_000, 2**31 - 1, 2**32 - 1]
    names = [augment.TRANSFORMATIONS[i % 6] for i in range(len(seeds))]
    ip, dp = augment.draw_params_batch(names, seeds, h, w, threads=1)
    for i, (n, s) in enumerate(zip(names, seeds)):
        eip, edp = _python_params(n, s, h, w)
        assert np.array_equal(ip[i], eip), (n, s, ip[i], eip)
        assert np.array_equal(dp[i].view(np.int64), edp.view(np.int64)), (n, s, dp[i], edp)     # same doubles, bit for bit


def test_native_params_threads_and_every_transform_per_seed():
    rng = random.Random(3)
    seeds = np.array([rng.randint(1, 1_000_000) for _ in range(4096)])
    for code in range(6):
        tr = np.full(len(seeds), code, np.int32)
        a = augment.draw_params_batch(tr, seeds, 256, 256, threads=1)
        b = augment.draw_params_batch(tr, seeds, 256, 256, threads=0)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    # the reference's own numbers for seed 42 on 256x256 (SURVEY 3.1 parameter trace, golden file)
    ip, dp = augment.draw_params_batch(list(augment.TRANSFORMATIONS), [42] * 6, 256, 256)
    random.seed(42)
    assert ip[0, 0] == (0 if random.choice([True, False]) else 1)
    random.seed(42)
    assert dp[1, 0] == random.uniform(-30, 30)
    random.seed(42)
    assert dp[5, 0] == random.uniform(0, 2) and math.isclose(dp[5, 0], 1.2789, abs_tol=1e-4)


def test_seed_zero_is_unseeded_like_the_reference():
    """ImageAugmenter(0) does not seed (image_augmenter.py:16 `if seed:`): the draw continues the global stream."""
    random.seed(1234)
    expect = random.uniform(-30, 30)
    random.seed(1234)
    ip, dp = augment.draw_params_batch(["rotate"], [0], 256, 256)
    assert dp[0, 0] == expect


def test_bad_transform_code_rejected():
    from leaffliction_b200._lib import LeafxError
    with pytest.raises(LeafxError):
        augment.draw_params_batch(np.array([9], np.int32), [5], 32, 32)


def test_native_task_list_equals_interpreter():
    """balance.task_arrays_for_labels (native stream) == balance.tasks_for_labels (random.Random) task for task."""
    from leaffliction_b200 import balance
    counts = balance.synthetic_class_counts()
    names = [c for p in counts.values() for c in p]
    plants = {p: list(c) for p, c in counts.items()}
    per = [max(1, n // 16) for p in counts.values() for n in p.values()]
    rng = np.random.default_rng(0)
    labels = rng.permutation(np.repeat(np.arange(len(names)), per))           # shuffled: class members are not contiguous
    for seed in (42, 7, 123456789):
        plan, tasks = balance.tasks_for_labels(labels, names, plants, seed=seed)
        plan2, ta = balance.task_arrays_for_labels(labels, names, plants, seed=seed)
        assert plan == plan2 and len(ta) == len(tasks)
        ref = augment.TaskArrays(tasks)
        assert np.array_equal(ta.transform, ref.transform)
        assert np.array_equal(ta.seed, ref.seed)
        assert np.array_equal(ta.source_index, ref.source_index)


def test_task_array_shards_partition_the_task_list():
    from leaffliction_b200 import balance
    counts = balance.synthetic_class_counts()
    names = [c for p in counts.values() for c in p]
    plants = {p: list(c) for p, c in counts.items()}
    labels = np.repeat(np.arange(len(names)), [max(1, n // 64) for p in counts.values() for n in p.values()])
    _, ta = balance.task_arrays_for_labels(labels, names, plants, seed=42)
    for world in (1, 2, 8):
        parts = [ta.shard(r, world) for r in range(world)]
        assert sum(len(p) for p in parts) == len(ta)
        merged_seed = np.empty(len(ta), np.int64)
        for r, p in enumerate(parts):
            merged_seed[r::world] = p.seed
            assert np.array_equal(p.transform, ta.transform[r::world]) and np.array_equal(p.source_index, ta.source_index[r::world])
        assert np.array_equal(merged_seed, ta.seed)


def test_native_params_property_random_shapes_and_seeds():
    """Property test: for arbitrary image shapes and seeds the native drawer equals the interpreter (crop boxes stay inside
    the image, rotate sizes cover the rotated corners)."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=150, deadline=None)
    @given(st.integers(8, 2048), st.integers(8, 2048), st.integers(1, 2**32 - 1), st.integers(0, 5))
    def check(h, w, seed, code):
        ip, dp = augment.draw_params_batch(np.array([code], np.int32), [seed], h, w, threads=1)
        eip, edp = _python_params(augment.TRANSFORMATIONS[code], seed, h, w)
        assert np.array_equal(ip[0], eip) and np.array_equal(dp[0].view(np.int64), edp.view(np.int64))
        if code == 4:
            left, top, nw, nh = ip[0, :4]
            assert 0 <= left and left + nw <= w and 0 <= top and top + nh <= h
        if code == 1:
            assert ip[0, 6] >= 1 and ip[0, 7] >= 1 and -30.0 <= dp[0, 0] <= 30.0
    check()


def test_params_from_stream_words_match_the_interpreter():
    """lfx_draw_augment_params_words (the host half of the device-seeded path): given the first outputs of each task's
    `random` stream -- produced here by the interpreter itself, on the GPU by lfx_seed_words -- it returns exactly what
    lfx_draw_augment_params returns; with too few words the task is re-seeded on the host (same results)."""
    import random

    from leaffliction_b200 import augment
    rng = np.random.default_rng(77)
    B = 1200
    tr = rng.integers(0, 6, B).astype(np.int32)
    seeds = rng.integers(1, 1000001, B).astype(np.int64)
    seeds[:4] = (1, 2**31, 2**32 - 1, 999983)
    ip0, dp0 = augment.draw_params_batch(tr, seeds, 256, 256)
    for nw in (16, 4):
        words = np.zeros((B, nw), np.uint32)
        for i, sd in enumerate(seeds):
            random.seed(int(sd))
            words[i] = [random.getrandbits(32) for _ in range(nw)]
        ip, dp = augment.draw_params_from_words(tr, seeds.astype(np.uint32), words, 256, 256)
        assert np.array_equal(ip, ip0) and np.array_equal(dp, dp0), nw
    ip1, dp1 = augment.draw_params_from_words(tr, seeds.astype(np.uint32), words, 96, 160)
    ip2, dp2 = augment.draw_params_batch(tr, seeds, 96, 160)
    assert np.array_equal(ip1, ip2) and np.array_equal(dp1, dp2)
