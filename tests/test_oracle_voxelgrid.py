"""Pins of the oracle's pcl::VoxelGrid restatement (SURVEY.md §8c pin iii): an independent
dict-based numpy reference must give the identical voxel -> count map, output order and centroids,
including the NaN / is_dense, overflow and min_points_per_voxel paths
[REF apps/prefiltering_nodelet.cpp:59-63,249-260,286]."""
import numpy as np
import pytest


def dict_voxelgrid(pts, leaf, min_points=0, is_dense=False):
    """Straight restatement of A.1 with a Python dict (float32 arithmetic via numpy scalars)."""
    leaf = np.asarray((leaf,) * 3 if np.isscalar(leaf) else leaf, np.float32)
    inv = (np.float32(1.0) / leaf).astype(np.float32)
    xyz = pts[:, :3]
    finite = np.isfinite(xyz).all(axis=1) if not is_dense else np.ones(len(pts), bool)
    good = xyz[finite]
    mn, mx = good.min(axis=0), good.max(axis=0)
    min_b = np.floor(mn * inv).astype(np.int64)
    max_b = np.floor(mx * inv).astype(np.int64)
    div_b = max_b - min_b + 1
    mul = np.array([1, div_b[0], div_b[0] * div_b[1]], np.int64)
    cells = {}
    keys = np.full(len(pts), 0xFFFFFFFF, np.uint32)
    for i in np.nonzero(finite)[0]:
        ijk = np.floor(xyz[i] * inv).astype(np.int64) - min_b
        k = int((ijk * mul).sum())
        keys[i] = k
        cells.setdefault(k, []).append(i)
    out, ids, counts = [], [], []
    for k in sorted(cells):
        idx = cells[k]
        if len(idx) < min_points:
            continue
        acc = np.zeros(3, np.float32)
        for i in idx:  # ascending input order, float accumulation
            acc = (acc + xyz[i]).astype(np.float32)
        out.append(acc / np.float32(len(idx)))
        ids.append(k)
        counts.append(len(idx))
    return np.array(out, np.float32).reshape(-1, 3), np.array(ids, np.uint32), np.array(counts, np.uint32), keys, min_b, div_b


@pytest.mark.parametrize("seed,leaf,min_points,is_dense", [(0, 0.25, 0, False), (1, 0.1, 0, True), (2, (0.5, 0.25, 1.0), 2, False), (3, 1.0, 3, False)])
def test_voxelgrid_matches_dict_reference(oracle, seed, leaf, min_points, is_dense):
    rng = np.random.default_rng(seed)
    n = 3000
    pts = np.ones((n, 4), np.float32)
    pts[:, :3] = rng.normal(0, 3, (n, 3)).astype(np.float32)
    pts[100:140] = pts[99]  # duplicates
    if not is_dense:
        pts[::11, 2] = np.nan
        pts[5::97, 0] = np.inf
    got = oracle.voxelgrid(pts, leaf, min_points_per_voxel=min_points, is_dense=is_dense)
    out, ids, counts, keys, min_b, div_b = dict_voxelgrid(pts, leaf, min_points, is_dense)
    assert not got["overflow"]
    assert np.array_equal(got["min_b"], min_b) and np.array_equal(got["div_b"], div_b)
    assert np.array_equal(got["key"], keys)
    assert np.array_equal(got["voxel_id"], ids)
    assert np.array_equal(got["count"], counts)
    assert np.array_equal(got["out"][:, :3].view(np.uint32), out.view(np.uint32))
    assert np.all(got["out"][:, 3] == 1.0)
    assert counts.sum() <= np.isfinite(pts[:, :3]).all(axis=1).sum()


def test_voxelgrid_output_sorted_and_idempotent_counts(oracle, scans):
    r = oracle.voxelgrid(scans["raw0"], 0.1)
    assert np.all(np.diff(r["voxel_id"].astype(np.int64)) > 0), "output is ordered by ascending linear voxel index"
    assert int(r["count"].sum()) == len(scans["raw0"])
    # every centroid lies inside its own voxel -> filtering the output again keeps one point per voxel
    r2 = oracle.voxelgrid(r["out"], 0.1)
    assert len(r2["out"]) == len(r["out"])


def test_voxelgrid_empty_single_and_overflow(oracle):
    e = oracle.voxelgrid(np.zeros((0, 4), np.float32), 0.1)
    assert len(e["out"]) == 0
    one = np.array([[1.5, -2.5, 0.25, 1.0]], np.float32)
    r = oracle.voxelgrid(one, 0.1)
    assert len(r["out"]) == 1 and np.array_equal(r["out"], one) and r["count"][0] == 1
    # dx*dy*dz > INT32_MAX: "Leaf size is too small for the input dataset" -> output = input copy
    far = np.ones((4, 4), np.float32)
    far[:, :3] = [[0, 0, 0], [4000, 0, 0], [0, 4000, 0], [0, 0, 4000]]
    r = oracle.voxelgrid(far, 0.001, is_dense=True)
    assert r["overflow"] and np.array_equal(r["out"], far)
    # all-NaN cloud with is_dense = false: nothing survives
    nan = np.full((8, 4), np.nan, np.float32)
    assert len(oracle.voxelgrid(nan, 0.1, is_dense=False)["out"]) == 0
