"""MapCloudGenerator::generate on the device (b200reg_map_cloud, csrc/map_cloud.cuh) against the CPU restatement of pcl's
octree (oracle.map_cloud) [REF src/hdl_graph_slam/map_cloud_generator.cpp:13-49]: transformed concatenation bit for bit, and
the occupied voxel centres in the octree's own order bit for bit — including the order-dependent growth of pcl's bounding box."""
import numpy as np
import pytest

from helpers import bits_equal
from test_oracle_map_cloud import keyframes, pose

pytestmark = pytest.mark.gpu


def same_cloud(a, b):
    """Bit-identical, NaN-aware: a non-finite input point turns into NaNs under the pose (inf * 0), and the payload / sign
    of a generated NaN is the one thing x86 and the GPU do not agree on."""
    a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    if a.shape != b.shape:
        return False
    na, nb = np.isnan(a), np.isnan(b)
    return bool(np.array_equal(na, nb) and np.array_equal(a.view(np.uint32)[~na], b.view(np.uint32)[~nb]))


def snapshots(clouds, poses):
    from delta_graph_slam_b200.map_cloud_generator import KeyFrameSnapshot
    return [KeyFrameSnapshot(p, c) for c, p in zip(clouds, poses)]


def test_unfiltered_and_voxel_centres_match_the_oracle_bit_for_bit(oracle):
    from delta_graph_slam_b200.map_cloud_generator import MapCloudGenerator
    gen = MapCloudGenerator()
    rng = np.random.default_rng(41)
    clouds, poses = keyframes(rng, n_kf=7, n_pts=4000)
    clouds[2][7, 0] = np.nan  # skipped by addPointsFromInputCloud
    clouds[5][100:110, 2] = np.inf
    kfs = snapshots(clouds, poses)
    for res in (0.0, -1.0):
        assert same_cloud(gen.generate(kfs, res), oracle.map_cloud(clouds, poses, 0.0))
    for res in (0.05, 0.1, 0.25, 1.0, 3.0):
        got, info = gen.generate(kfs, res, details=True)
        want, winfo = oracle.map_cloud(clouds, poses, res, details=True)
        assert info["depth"] == winfo["depth"] and np.array_equal(info["min"], winfo["min"]), res
        assert bits_equal(got, want), res


def test_point_order_decides_the_box_as_in_pcl(oracle):
    """The same points in another keyframe order give another octree origin (pcl grows the box as the points arrive): the
    device path must follow the oracle through that, not merely produce SOME voxelisation."""
    from delta_graph_slam_b200.map_cloud_generator import MapCloudGenerator
    gen = MapCloudGenerator()
    rng = np.random.default_rng(43)
    clouds, poses = keyframes(rng, n_kf=6, n_pts=2500)
    origins = set()
    for perm in ([0, 1, 2, 3, 4, 5], [5, 4, 3, 2, 1, 0], [3, 0, 5, 1, 4, 2]):
        c, p = [clouds[i] for i in perm], [poses[i] for i in perm]
        got, info = gen.generate(snapshots(c, p), 0.2, details=True)
        want, winfo = oracle.map_cloud(c, p, 0.2, details=True)
        assert bits_equal(got, want) and np.array_equal(info["min"], winfo["min"])
        origins.add(tuple(info["min"]))
    assert len(origins) >= 2


def test_full_size_keyframes_and_the_cached_form(oracle):
    """Down-sampled HDL-64 keyframes along the synthetic trajectory (what the graph holds), resolution 0.05 m as the
    reference's map_cloud_resolution; the cached form reads the clouds from the loop detector's keyframe cache."""
    import delta_graph_slam_b200 as eng
    from delta_graph_slam_b200.map_cloud_generator import MapCloudGenerator
    P0 = oracle.synth_traj(0)
    clouds, poses = [], []
    for k in range(0, 40, 8):
        P = oracle.synth_traj(k)
        clouds.append(oracle.voxelgrid(oracle.synth_scan(P, noise_seed=1000 + k), 0.1)["out"])
        poses.append(np.linalg.inv(P0) @ P)
    reg = eng.Registration()
    gen = MapCloudGenerator(registration=reg)
    got = gen.generate(snapshots(clouds, poses), 0.05)
    want = oracle.map_cloud(clouds, [np.asarray(p, np.float64).astype(np.float32) for p in poses], 0.05)
    assert len(got) > 100_000 and bits_equal(got, want)
    for i, c in enumerate(clouds):
        reg.cloudPut(100 + i, c)
    assert bits_equal(gen.generate_cached([100 + i for i in range(len(clouds))], poses, 0.05), want)


def test_edge_cases(oracle):
    import delta_graph_slam_b200 as eng
    from delta_graph_slam_b200.map_cloud_generator import MapCloudGenerator
    gen = MapCloudGenerator()
    assert gen.generate([], 0.1) is None
    one = np.array([[1.5, -2.25, 0.75, 1.0]], np.float32)
    empty = np.zeros((0, 4), np.float32)
    I = np.eye(4)
    for clouds in ([one], [empty, one, empty], [one, one], [np.full((5, 4), np.nan, np.float32)], [empty]):
        poses = [I] * len(clouds)
        for res in (0.0, 0.1):
            assert same_cloud(gen.generate(snapshots(clouds, poses), res), oracle.map_cloud(clouds, poses, res)), (len(clouds), res)
    # a map too wide for 21 octree levels is refused, never voxelised wrongly
    far = np.array([[0, 0, 0, 1], [3.0e6, 0, 0, 1]], np.float32)
    with pytest.raises(eng.B200RegError):
        gen.generate(snapshots([far], [I]), 0.01)
    with pytest.raises(eng.B200RegError):
        gen.generate_cached([424242], [I], 0.1)
