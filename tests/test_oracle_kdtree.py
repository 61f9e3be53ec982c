"""Pins of the oracle's exact k-d tree (SURVEY.md §8c pin iv) and of the fitness loop, whose
definition is the one piece of hot-path arithmetic held in the reference tree
[REF src/hdl_graph_slam/information_matrix_calculator.cpp:77-108]."""
import numpy as np


def brute_knn(points, queries, k):
    p = points[:, :3].astype(np.float32)
    idx = np.zeros((len(queries), k), np.int32)
    d2 = np.zeros((len(queries), k), np.float32)
    for i, q in enumerate(queries[:, :3].astype(np.float32)):
        d = p - q
        dist = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]).astype(np.float32) + d[:, 2] * d[:, 2]  # FLANN L2_Simple order
        order = np.lexsort((np.arange(len(p)), dist))[:k]
        idx[i], d2[i] = order, dist[order]
    return idx, d2


def test_knn_matches_brute_force(oracle):
    rng = np.random.default_rng(5)
    pts = np.ones((4000, 4), np.float32)
    pts[:, :3] = rng.uniform(-10, 10, (4000, 3)).astype(np.float32)
    pts[500:520] = pts[499]  # exact ties -> lowest index first
    q = np.ones((10000, 4), np.float32)
    q[:, :3] = rng.uniform(-12, 12, (10000, 3)).astype(np.float32)
    q[:100] = pts[450:550]
    for k in (1, 20):
        idx, d2 = oracle.knn(pts, q[:2000] if k == 20 else q, k)
        ridx, rd2 = brute_knn(pts, q[:2000] if k == 20 else q, k)
        assert np.array_equal(idx, ridx)
        assert np.array_equal(d2.view(np.uint32), rd2.view(np.uint32))


def test_knn_fewer_points_than_k(oracle):
    pts = np.array([[0, 0, 0, 1], [1, 0, 0, 1], [0, 2, 0, 1]], np.float32)
    idx, d2 = oracle.knn(pts, pts[:1], 3)
    assert idx.tolist() == [[0, 1, 2]] and d2.tolist() == [[0.0, 1.0, 4.0]]


def reference_fitness(target, source, T, max_range):
    """calc_fitness_score: transform cloud2, 1-NN in cloud1, mean of squared distances <= max_range,
    DBL_MAX when nothing qualifies [REF information_matrix_calculator.cpp:77-108]."""
    T = np.asarray(T, np.float32)
    q = np.ones_like(source)
    for r in range(3):
        q[:, r] = ((T[r, 0] * source[:, 0] + T[r, 1] * source[:, 1]) + T[r, 2] * source[:, 2]) + T[r, 3]
    _, d2 = brute_knn(target, q, 1)
    d2 = d2[:, 0].astype(np.float64)
    sel = d2 <= max_range
    return d2[sel].sum() / sel.sum() if sel.any() else np.finfo(np.float64).max


def test_fitness_score_follows_the_in_tree_definition(oracle):
    rng = np.random.default_rng(9)
    tgt = np.ones((3000, 4), np.float32)
    tgt[:, :3] = rng.uniform(-8, 8, (3000, 3)).astype(np.float32)
    src = tgt[::3].copy()
    src[:, :3] += rng.normal(0, 0.05, (len(src), 3)).astype(np.float32)
    reg = oracle.Registration(oracle.NDT, resolution=2.0)
    reg.setInputTarget(tgt)
    reg.setInputSource(src)
    # before any align final_transformation_ is the identity
    for max_range in (np.finfo(np.float64).max, 0.01, 1e-9):
        want = reference_fitness(tgt, src, np.eye(4), max_range)
        got = reg.getFitnessScore(max_range)
        assert got == want or abs(got - want) <= 1e-12 * want
    assert reg.getFitnessScore(-1.0) == np.finfo(np.float64).max
