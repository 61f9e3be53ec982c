"""Oracle restatement of MapCloudGenerator::generate [REF src/hdl_graph_slam/map_cloud_generator.cpp:13-49] (SURVEY.md §8f rank 4,
second half): checked through properties that do not depend on the restatement's own arithmetic.  The GPU counterpart is next round's."""
import numpy as np

from helpers import bits_equal


def pose(yaw, t):
    c, s = np.cos(yaw), np.sin(yaw)
    T = np.eye(4, dtype=np.float32)
    T[:3, :3] = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], np.float32)
    T[:3, 3] = t
    return T


def keyframes(rng, n_kf=5, n_pts=3000):
    clouds, poses = [], []
    for k in range(n_kf):
        c = np.ones((n_pts, 4), np.float32)
        c[:, :3] = rng.normal(0, [6, 6, 1.0], (n_pts, 3)).astype(np.float32)
        clouds.append(c)
        poses.append(pose(0.3 * k, [4.0 * k, -1.5 * k, 0.1 * k]))
    return clouds, poses


def morton(keys, depth):
    code = np.zeros(len(keys), dtype=object)
    for b in range(depth - 1, -1, -1):
        for a in range(3):
            code = code * 2 + ((keys[:, a] >> b) & 1)
    return code


def test_unfiltered_map_is_the_transformed_concatenation(oracle):
    clouds, poses = keyframes(np.random.default_rng(31))
    got = oracle.map_cloud(clouds, poses, 0.0)
    want = np.concatenate([(c[:, :1] * T[:, 0] + c[:, 1:2] * T[:, 1]).astype(np.float32) + c[:, 2:3] * T[:, 2] + T[:, 3] for c, T in zip(clouds, poses)]).astype(np.float32)
    assert bits_equal(got, want) and (got[:, 3] == 1).all()
    assert bits_equal(oracle.map_cloud(clouds, poses, -1.0), got)


def test_voxel_centres_properties(oracle):
    rng = np.random.default_rng(32)
    clouds, poses = keyframes(rng)
    clouds[2][7, 0] = np.nan  # skipped by addPointsFromInputCloud
    world = oracle.map_cloud(clouds, poses, 0.0)
    for res in (0.05, 0.25, 1.0):
        centres, info = oracle.map_cloud(clouds, poses, res, details=True)
        fin = world[np.isfinite(world[:, :3]).all(axis=1), :3].astype(np.float64)
        # the box is a cube of 2^depth cells that holds every point, anchored one cell below the first point
        side = res * 2 ** info["depth"]
        assert (fin >= info["min"]).all() and (fin < info["min"] + side).all()
        assert np.allclose((world[0, :3] - info["min"]) / res % 1.0, 0.0, atol=1e-4) or np.allclose((world[0, :3] - info["min"]) / res % 1.0, 1.0, atol=1e-4)
        # one centre per occupied cell, each centre in the middle of its cell, every point inside the cell of some centre
        keys = np.floor((fin - info["min"]) / res).astype(np.int64)
        uniq = np.unique(keys, axis=0)
        assert abs(len(centres) - len(uniq)) <= 2  # see the boundary note below
        ckeys = np.floor((centres[:, :3].astype(np.float64) - info["min"]) / res).astype(np.int64)
        assert len(np.unique(ckeys, axis=0)) == len(centres)
        assert np.abs((ckeys + 0.5) * res + info["min"] - centres[:, :3]).max() < 1e-5 * max(1.0, np.abs(centres[:, :3]).max())
        # upstream keys a point with the box origin OF THE MOMENT and later origins differ from it by whole cells only in exact
        # arithmetic: a point within rounding of a cell face can land one cell over from the key recomputed with the final origin
        a, b = set(map(tuple, ckeys)), set(map(tuple, uniq))
        odd = a ^ b
        assert len(odd) <= 4
        for cell in odd:
            other = b if cell in a else a
            assert any(max(abs(cell[0] - o[0]), abs(cell[1] - o[1]), abs(cell[2] - o[2])) == 1 for o in other if abs(cell[0] - o[0]) <= 1)
        # getOccupiedVoxelCenters walks depth-first, child index = x << 2 | y << 1 | z: strictly ascending Morton code
        code = morton(ckeys, info["depth"])
        assert all(code[i] < code[i + 1] for i in range(len(code) - 1))
    # the lattice hangs on the FIRST point only: shuffling the others leaves the centres unchanged (up to the boundary note)
    cat = np.concatenate([c for c in clouds])
    eye = [np.eye(4, dtype=np.float32)]
    base = oracle.map_cloud([cat], eye, 1.0)
    perm = np.concatenate([[0], 1 + rng.permutation(len(cat) - 1)])
    shuffled = oracle.map_cloud([cat[perm]], eye, 1.0)
    assert abs(len(base) - len(shuffled)) <= 2 and len(set(map(tuple, base[:, :3])) ^ set(map(tuple, shuffled[:, :3]))) <= 4
    # a single point: one centre, the point itself sits in the middle of the upper cell of a two-cell box
    one = np.array([[1.25, -3.5, 0.75, 1.0]], np.float32)
    c1, i1 = oracle.map_cloud([one], [np.eye(4, dtype=np.float32)], 0.5, details=True)
    assert i1["depth"] == 1 and np.allclose(i1["min"], one[0, :3] - 0.5) and len(c1) == 1 and np.allclose(c1[0, :3], one[0, :3] + 0.0, atol=0.25 + 1e-6)
    assert len(oracle.map_cloud([np.zeros((0, 4), np.float32)], [np.eye(4, dtype=np.float32)], 0.5)) == 0
