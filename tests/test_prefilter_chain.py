"""The stages either side of VoxelGrid in PrefilteringNodelet::cloud_callback
[REF apps/prefiltering_nodelet.cpp:150-153 chain; :275-291 distance_filter; :88-96,262-273 RadiusOutlierRemoval; :77-87 StatisticalOutlierRemoval]."""
import os

import numpy as np
import pytest

from helpers import bits_equal

DEVNULL = open(os.devnull, "w")
# the reference's launch file: launch/delta_graph_slam.launch:31-42
LAUNCH = dict(downsample_method="VOXELGRID", downsample_resolution=0.1, use_distance_filter=True, distance_near_thresh=0.1, distance_far_thresh=100.0,
              outlier_removal_method="RADIUS", radius_radius=0.5, radius_min_neighbors=2)


def np_distance_filter(c, near, far):
    sq = (c[:, 0] * c[:, 0] + c[:, 1] * c[:, 1]).astype(np.float32) + c[:, 2] * c[:, 2]
    d = np.sqrt(sq.astype(np.float32)).astype(np.float64)
    with np.errstate(invalid="ignore"):
        return c[(d > near) & (d < far)]


def np_radius_outlier_removal(c, radius, min_neighbors):
    p = c[:, :3].astype(np.float32)
    keep = np.zeros(len(c), bool)
    r2 = np.float32(radius * radius)
    for i in range(len(c)):
        with np.errstate(invalid="ignore"):
            d = p - p[i]
            d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]).astype(np.float32) + d[:, 2] * d[:, 2]
            keep[i] = np.count_nonzero(d2 < r2) > min_neighbors
    return c[keep]


def cloud_with_outliers(oracle, rng):
    raw = oracle.synth_scan(oracle.synth_traj(2), noise_seed=1002)[::5].copy()
    far = np.ones((40, 4), np.float32)
    far[:, :3] = rng.uniform(-150, 150, (40, 3)).astype(np.float32)
    bad = np.ones((3, 4), np.float32)
    bad[0, 0], bad[1, 1], bad[2, 2] = np.nan, np.inf, -np.inf
    return np.concatenate([raw[:100], far[:20], bad, raw[100:], far[20:]]).astype(np.float32)


def test_oracle_distance_filter_and_outlier_removal_follow_their_definitions(oracle):
    rng = np.random.default_rng(3)
    c = cloud_with_outliers(oracle, rng)
    for near, far in ((1.0, 100.0), (0.1, 100.0), (5.0, 20.0)):
        assert bits_equal(oracle.distance_filter(c, near, far), np_distance_filter(c, near, far))
    small = oracle.voxelgrid(oracle.distance_filter(c, 1.0, 60.0), 0.4)["out"][:3000]
    small = np.concatenate([small[:500], np.array([[np.nan, 0, 0, 1], [0, np.inf, 0, 1]], np.float32), small[500:]])  # never kept, never counted
    for radius, mn in ((0.5, 2), (0.8, 2), (1.3, 5)):
        assert bits_equal(oracle.radius_outlier_removal(small, radius, mn), np_radius_outlier_removal(small, radius, mn))
    assert len(oracle.radius_outlier_removal(np.zeros((0, 4), np.float32), 0.5, 2)) == 0


def np_statistical_outlier_removal(c, mean_k, stddev_mul):
    """pcl::StatisticalOutlierRemoval::applyFilterIndices written out with brute-force neighbour searches."""
    p = c[:, :3].astype(np.float32)
    finite = np.isfinite(p).all(axis=1)
    tree = p[finite]
    dist = np.zeros(len(c), np.float32)
    valid = 0
    for i in range(len(c)):
        if not finite[i] or len(tree) < mean_k + 1:
            continue
        d = tree - p[i]
        d2 = np.sort((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]).astype(np.float32) + d[:, 2] * d[:, 2])[: mean_k + 1]
        acc = 0.0
        for v in np.sqrt(d2[1:]).astype(np.float32):
            acc += float(v)
        dist[i] = np.float32(acc / mean_k)
        valid += 1
    s = sq = 0.0
    for v in dist:
        s += float(v)
        sq += float(np.float32(v * v))
    with np.errstate(invalid="ignore", divide="ignore"):
        mean = np.float64(s) / np.float64(valid)
        var = (np.float64(sq) - np.float64(s) * np.float64(s) / np.float64(valid)) / (np.float64(valid) - 1.0)
        thr = mean + stddev_mul * np.sqrt(var)
        return c[~(dist.astype(np.float64) > thr)], dist, thr


def small_cloud(oracle, rng, n=1500):
    c = oracle.voxelgrid(oracle.distance_filter(cloud_with_outliers(oracle, rng), 1.0, 60.0), 0.4)["out"][:n]
    return np.concatenate([c[:500], np.array([[np.nan, 0, 0, 1], [0, np.inf, 0, 1]], np.float32), c[500:]])


def line_cloud(n=64):
    """Equally spaced points: every mean neighbour distance is the same float, the variance is exactly 0."""
    c = np.ones((n, 4), np.float32)
    c[:, 0] = np.arange(n, dtype=np.float32) * np.float32(0.5)
    c[:, 1:3] = 0
    return c


def test_oracle_statistical_outlier_removal_follows_its_definition(oracle):
    c = small_cloud(oracle, np.random.default_rng(8))
    for mean_k, mul in ((20, 1.0), (10, 0.5), (31, 2.0), (1, 0.0)):
        got, det = oracle.statistical_outlier_removal(c, mean_k, mul, details=True)
        want, dist, thr = np_statistical_outlier_removal(c, mean_k, mul)
        assert bits_equal(det["distances"], dist) and det["threshold"] == thr
        assert bits_equal(got, want) and 0 < len(got) < len(c)
        assert np.isnan(got[:, :3]).any() and np.isinf(got[:, :3]).any()  # uncounted points stand at distance 0 and stay (upstream behaviour)
    # fewer finite points than mean_k + 1: no search succeeds, the cut is NaN, everything stays
    tiny = c[495:507]
    got, det = oracle.statistical_outlier_removal(tiny, 20, 1.0, details=True)
    assert bits_equal(got, tiny) and np.isnan(det["threshold"]) and not det["distances"].any()
    assert len(oracle.statistical_outlier_removal(np.zeros((0, 4), np.float32), 20, 1.0)) == 0
    # equally spaced points with one neighbour each: all distances equal, variance exactly 0, nothing removed
    got, det = oracle.statistical_outlier_removal(line_cloud(), 1, 1.0, details=True)
    assert len(got) == 64 and det["stddev"] == 0.0 and det["threshold"] == 0.5


def test_prefilter_mirror_reads_the_reference_parameters(monkeypatch):
    """initialize_params [REF apps/prefiltering_nodelet.cpp:55-102] on recorders instead of engine handles: parameter names,
    the nodelet's defaults for the chosen methods, and its console lines."""
    import io
    import delta_graph_slam_b200.odometry as odo
    calls = []

    class Rec:
        def __init__(self, name):
            self._name = name
            self._reg = object()

        def __getattr__(self, attr):
            return lambda *a, **k: calls.append((self._name, attr) + a)

    monkeypatch.setattr(odo, "VoxelGrid", lambda device=0: Rec("vg"))
    monkeypatch.setattr(odo, "StatisticalOutlierRemoval", lambda device=0, registration=None: Rec("sor"))
    monkeypatch.setattr(odo, "RadiusOutlierRemoval", lambda device=0, registration=None: Rec("ror"))
    out = io.StringIO()
    odo.Prefilter(dict(outlier_removal_method="STATISTICAL", use_distance_filter=True), out=out)
    assert out.getvalue().splitlines() == ["downsample: VOXELGRID 0.1", "outlier_removal: STATISTICAL 20 - 1"]
    assert calls == [("vg", "setLeafSize", 0.1, 0.1, 0.1), ("sor", "setMeanK", 20), ("sor", "setStddevMulThresh", 1.0), ("vg", "setDistanceFilter", True, 1.0, 100.0)]
    calls.clear()
    out = io.StringIO()
    odo.Prefilter(dict(LAUNCH, statistical_mean_k=30), out=out)  # the launch file: RADIUS 0.5 - 2, gate 0.1 .. 100
    assert out.getvalue().splitlines() == ["downsample: VOXELGRID 0.1", "outlier_removal: RADIUS 0.5 - 2"]
    assert calls == [("vg", "setLeafSize", 0.1, 0.1, 0.1), ("ror", "setRadiusSearch", 0.5), ("ror", "setMinNeighborsInRadius", 2), ("vg", "setDistanceFilter", True, 0.1, 100.0)]
    calls.clear()
    out = io.StringIO()
    odo.Prefilter(dict(downsample_method="VOXELGRID", downsample_resolution=0.25, outlier_removal_method="STATISTICAL", statistical_mean_k=12, statistical_stddev=2.5), out=out)
    assert out.getvalue().splitlines() == ["downsample: VOXELGRID 0.25", "outlier_removal: STATISTICAL 12 - 2.5"]
    assert ("sor", "setMeanK", 12) in calls and ("sor", "setStddevMulThresh", 2.5) in calls
    with pytest.raises(NotImplementedError):
        odo.Prefilter(dict(downsample_method="APPROX_VOXELGRID"), out=io.StringIO())
    # no parameters at all: the nodelet's own defaults [REF apps/prefiltering_nodelet.cpp:56-57,77-80,100-102]
    calls.clear()
    out = io.StringIO()
    pre = odo.Prefilter({}, out=out)
    assert out.getvalue().splitlines() == ["downsample: VOXELGRID 0.1", "outlier_removal: STATISTICAL 20 - 1"]
    assert calls == [("vg", "setLeafSize", 0.1, 0.1, 0.1), ("sor", "setMeanK", 20), ("sor", "setStddevMulThresh", 1.0), ("vg", "setDistanceFilter", True, 1.0, 100.0)]
    # use_distance_filter = false changes nothing: the reference reads the flag (:100) and gates every scan anyway (:150)
    calls.clear()
    odo.Prefilter(dict(use_distance_filter=False, outlier_removal_method="NONE"), out=io.StringIO())
    assert calls == [("vg", "setLeafSize", 0.1, 0.1, 0.1), ("vg", "setDistanceFilter", True, 1.0, 100.0)]
    # the mirror's own opt-out
    calls.clear()
    pre = odo.Prefilter(dict(b200_skip_distance_filter=True, outlier_removal_method="NONE"), out=io.StringIO())
    assert calls == [("vg", "setLeafSize", 0.1, 0.1, 0.1)] and not pre.distance_filter_on
    # downsample NONE: the gate becomes a call of its own on the outlier filter's handle
    calls.clear()
    out = io.StringIO()
    pre = odo.Prefilter(dict(downsample_method="NONE", outlier_removal_method="RADIUS"), out=out)
    assert out.getvalue().splitlines() == ["downsample: NONE", "outlier_removal: RADIUS 0.8 - 2"]
    assert pre.filter is None and pre.distance_filter_on and pre._reg is pre.outlier_removal_filter._reg


def test_prefilter_mirror_base_link_step_on_recorders(monkeypatch):
    """The base_link step of cloud_callback [REF apps/prefiltering_nodelet.cpp:123-148] on recorders: the x / y translation
    is zeroed before the matrix reaches the handle of the chain's FIRST stage, the remaining translation is the lidar
    position, None switches the transform off, and a chain with no first stage to attach it to says so."""
    import io
    import delta_graph_slam_b200.odometry as odo
    calls = []

    class Handle:
        def setInputTransform(self, m):
            calls.append(None if m is None else np.array(m, np.float64).copy())

    class Rec:
        def __init__(self):
            self._reg = Handle()

        def __getattr__(self, attr):
            return lambda *a, **k: None

    monkeypatch.setattr(odo, "VoxelGrid", lambda device=0: Rec())
    monkeypatch.setattr(odo, "StatisticalOutlierRemoval", lambda device=0, registration=None: Rec())
    monkeypatch.setattr(odo, "RadiusOutlierRemoval", lambda device=0, registration=None: Rec())
    m = base_link_transform(None)
    pre = odo.Prefilter(LAUNCH, out=io.StringIO())
    lidar = pre.setBaseLinkTransform(m)
    assert np.array_equal(lidar, [0.0, 0.0, 1.9]) and np.array_equal(pre.lidar_position, lidar)
    want = m.copy()
    want[0, 3] = want[1, 3] = 0.0
    assert len(calls) == 1 and np.array_equal(calls[0], want) and m[0, 3] == 0.8, "the caller's matrix is not modified"
    assert np.array_equal(pre.setBaseLinkTransform(None), [0.0, 0.0, 0.0]) and calls[-1] is None
    # down-sampling NONE: the stand-alone gate's handle (the outlier filter's) is the first stage
    calls.clear()
    pre = odo.Prefilter(dict(downsample_method="NONE", outlier_removal_method="RADIUS"), out=io.StringIO())
    pre.setBaseLinkTransform(m)
    assert len(calls) == 1 and np.array_equal(calls[0], want)
    # neither a VoxelGrid nor the gate: nothing to attach the transform to
    pre = odo.Prefilter(dict(downsample_method="NONE", outlier_removal_method="RADIUS", b200_skip_distance_filter=True), out=io.StringIO())
    with pytest.raises(ValueError):
        pre.setBaseLinkTransform(m)
    assert np.array_equal(pre.setBaseLinkTransform(None), [0.0, 0.0, 0.0])


def base_link_transform(rng):
    """a sensor -> base_link transform as tf would hand it over: tilted, yawed, mounted 1.9 m up and off-centre"""
    a, b, c = 0.03, -0.02, 0.7
    rx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
    rz = np.array([[np.cos(c), -np.sin(c), 0], [np.sin(c), np.cos(c), 0], [0, 0, 1]])
    m = np.eye(4)
    m[:3, :3] = rz @ ry @ rx
    m[:3, 3] = [0.8, -0.15, 1.9]
    return m


def test_oracle_base_link_transform_is_the_double_expression_rounded_once(oracle):
    """pcl::transformPointCloud<PointT, double> [REF apps/prefiltering_nodelet.cpp:137-147]: per coordinate
    m(r,0) x + m(r,1) y + m(r,2) z + m(r,3) in double, left to right, then one rounding to float; w = 1; the
    non-finite points of a non-dense cloud stay untouched."""
    rng = np.random.default_rng(12)
    c = cloud_with_outliers(oracle, rng)
    c[:, 3] = 7.0  # whatever the sensor driver left in the padding
    m = base_link_transform(rng)
    got = oracle.transform_cloud_d(c, m, is_dense=False)
    fin = np.isfinite(c[:, :3]).all(axis=1)
    x, y, z = (c[fin, k].astype(np.float64) for k in range(3))
    for r in range(3):
        want = (((m[r, 0] * x) + (m[r, 1] * y)) + (m[r, 2] * z)) + m[r, 3]
        assert np.array_equal(got[fin, r], want.astype(np.float32))
    assert np.all(got[fin, 3] == 1.0)
    assert bits_equal(got[~fin], c[~fin]) and (~fin).sum() == 3


@pytest.mark.gpu
def test_base_link_transform_in_front_of_the_chain_matches_the_oracle(oracle):
    """cloud_callback with base_link_frame set [REF apps/prefiltering_nodelet.cpp:123-153]: transform (x / y translation
    zeroed) -> distance_filter -> VoxelGrid -> outlier removal, every stage bit-exact; lidar_position is what the height
    filter gets; None returns to the sensor frame."""
    import delta_graph_slam_b200 as eng
    rng = np.random.default_rng(13)
    c = cloud_with_outliers(oracle, rng)
    m = base_link_transform(rng)
    m0 = m.copy()
    m0[0, 3] = m0[1, 3] = 0.0
    moved = oracle.transform_cloud_d(c, m0, is_dense=False)
    assert not bits_equal(moved, c)
    # launch-file chain
    pre = eng.Prefilter(LAUNCH, out=DEVNULL)
    lidar = pre.setBaseLinkTransform(m)
    assert np.array_equal(lidar, [0.0, 0.0, 1.9])
    want = oracle.radius_outlier_removal(oracle.voxelgrid(oracle.distance_filter(moved, 0.1, 100.0), 0.1, is_dense=False)["out"], 0.5, 2)
    assert bits_equal(pre.filter3d(c), want)
    # the same from a device-resident scan
    import torch
    t = torch.from_numpy(c).cuda()
    bufs = [torch.empty_like(t) for _ in range(2)]
    f = pre.filter3d(eng.DeviceCloud(t.data_ptr(), len(t), t), out=eng.DeviceCloud(bufs[0].data_ptr(), len(t), bufs[0]), out2=eng.DeviceCloud(bufs[1].data_ptr(), len(t), bufs[1]))
    assert f.n == len(want) and bits_equal(bufs[1][: f.n].cpu().numpy(), want)
    # a chain without a down-sampler: the stand-alone gate is the first stage and takes the transform
    pre2 = eng.Prefilter(dict(downsample_method="NONE", outlier_removal_method="NONE", distance_near_thresh=5.0, distance_far_thresh=20.0), out=DEVNULL)
    pre2.setBaseLinkTransform(m)
    assert bits_equal(pre2.filter3d(c), oracle.distance_filter(moved, 5.0, 20.0))
    # the VoxelGrid alone (no gate), and off again
    vg = eng.VoxelGrid()
    vg.setLeafSize(0.1, 0.1, 0.1)
    vg._reg.setInputTransform(m0)
    vg.setInputCloud(c, is_dense=False)
    assert bits_equal(vg.filter(), oracle.voxelgrid(moved, 0.1, is_dense=False)["out"])
    vg._reg.setInputTransform(None)
    assert bits_equal(vg.filter(), oracle.voxelgrid(c, 0.1, is_dense=False)["out"])
    pre.setBaseLinkTransform(None)
    assert bits_equal(pre.filter3d(c), oracle.radius_outlier_removal(oracle.voxelgrid(oracle.distance_filter(c, 0.1, 100.0), 0.1, is_dense=False)["out"], 0.5, 2))


@pytest.mark.gpu
def test_fused_distance_filter_matches_filter_then_voxelgrid(oracle):
    import delta_graph_slam_b200 as eng
    rng = np.random.default_rng(4)
    c = cloud_with_outliers(oracle, rng)
    for near, far in ((1.0, 100.0), (0.1, 100.0), (5.0, 20.0)):
        want = oracle.voxelgrid(oracle.distance_filter(c, near, far), 0.1, is_dense=False)["out"]
        vg = eng.VoxelGrid()
        vg.setLeafSize(0.1, 0.1, 0.1)
        vg.setDistanceFilter(True, near, far)
        vg.setInputCloud(c, is_dense=False)
        assert bits_equal(vg.filter(), want)
    # gate that lets nothing through -> empty cloud; gate off again -> plain VoxelGrid
    vg.setDistanceFilter(True, 500.0, 600.0)
    assert len(vg.filter()) == 0
    vg.setDistanceFilter(False)
    assert bits_equal(vg.filter(), oracle.voxelgrid(c, 0.1, is_dense=False)["out"])


@pytest.mark.gpu
def test_leaf_too_small_with_the_gate_returns_the_gated_input(oracle):
    import delta_graph_slam_b200 as eng
    rng = np.random.default_rng(5)
    c = cloud_with_outliers(oracle, rng)
    vg = eng.VoxelGrid()
    vg.setLeafSize(1e-4, 1e-4, 1e-4)  # extents / leaf overflow int32: PCL warns and copies its input
    vg.setDistanceFilter(True, 1.0, 100.0)
    vg.setInputCloud(c, is_dense=False)
    gated = oracle.distance_filter(c, 1.0, 100.0)
    ref = oracle.voxelgrid(gated, 1e-4, is_dense=False)
    assert ref["overflow"] and bits_equal(ref["out"], gated)
    assert bits_equal(vg.filter(), gated)


@pytest.mark.gpu
@pytest.mark.parametrize("radius,min_neighbors", [(0.5, 2), (0.8, 2), (1.3, 5)])
def test_radius_outlier_removal_matches_the_oracle(oracle, radius, min_neighbors):
    import torch
    import delta_graph_slam_b200 as eng
    rng = np.random.default_rng(6)
    c = oracle.voxelgrid(oracle.distance_filter(cloud_with_outliers(oracle, rng), 0.1, 100.0), 0.1)["out"]
    c = np.concatenate([c, np.array([[np.nan, 0, 0, 1]], np.float32)])  # non-finite points are dropped
    want = oracle.radius_outlier_removal(c, radius, min_neighbors)
    ror = eng.RadiusOutlierRemoval()
    ror.setRadiusSearch(radius)
    ror.setMinNeighborsInRadius(min_neighbors)
    ror.setInputCloud(c)
    assert bits_equal(ror.filter(), want)
    # device-resident and page-locked (zero-copy) outputs
    d_in = torch.from_numpy(c).cuda()
    d_out = torch.zeros_like(d_in)
    ror.setInputCloud(eng.DeviceCloud(d_in.data_ptr(), len(c), d_in))
    got = ror.filter(out=eng.DeviceCloud(d_out.data_ptr(), len(c), d_out))
    assert got.n == len(want) and bits_equal(d_out[: got.n].cpu().numpy(), want)
    h_out = torch.zeros((len(c), 4), dtype=torch.float32, pin_memory=True)
    ror.setInputCloud(c)
    ror.filter_begin(h_out.numpy())
    assert bits_equal(ror.filter_end(), want) and not h_out.numpy()[len(want):].any()


@pytest.mark.gpu
@pytest.mark.parametrize("mean_k,mul", [(20, 1.0), (10, 0.5), (31, 2.0), (1, 0.0)])
def test_statistical_outlier_removal_matches_the_oracle(oracle, mean_k, mul):
    import torch
    import delta_graph_slam_b200 as eng
    rng = np.random.default_rng(9)
    c = oracle.voxelgrid(oracle.distance_filter(cloud_with_outliers(oracle, rng), 0.1, 100.0), 0.1)["out"]
    c = np.concatenate([c[:1000], np.array([[np.nan, 0, 0, 1], [0, -np.inf, 0, 1]], np.float32), c[1000:]])  # uncounted, kept (upstream behaviour)
    want, det = oracle.statistical_outlier_removal(c, mean_k, mul, details=True)
    sor = eng.StatisticalOutlierRemoval()
    sor.setMeanK(mean_k)
    sor.setStddevMulThresh(mul)
    sor.setInputCloud(c)
    assert bits_equal(sor.filter(), want) and 0 < len(want) < len(c)
    st = sor.last_stats(len(c))
    assert bits_equal(st["distances"], det["distances"]) and st["valid"] == len(c) - 2
    # the index-order sums are only taken when a distance sits within rounding of the cut; the tree sums agree to ~n * 2^-53
    for key in ("mean", "stddev", "threshold"):
        assert st[key] == det[key] if st["exact_pass"] else abs(st[key] - det[key]) <= 1e-11 * abs(det[key]), key
    # device-resident and page-locked (zero-copy) outputs
    d_in = torch.from_numpy(c).cuda()
    d_out = torch.zeros_like(d_in)
    sor.setInputCloud(eng.DeviceCloud(d_in.data_ptr(), len(c), d_in))
    got = sor.filter(out=eng.DeviceCloud(d_out.data_ptr(), len(c), d_out))
    assert got.n == len(want) and bits_equal(d_out[: got.n].cpu().numpy(), want)
    h_out = torch.zeros((len(c), 4), dtype=torch.float32, pin_memory=True)
    sor.setInputCloud(c)
    sor.filter_begin(h_out.numpy())
    assert bits_equal(sor.filter_end(), want) and not h_out.numpy()[len(want):].any()


@pytest.mark.gpu
def test_statistical_outlier_removal_edge_cases(oracle):
    import delta_graph_slam_b200 as eng
    sor = eng.StatisticalOutlierRemoval()
    sor.setMeanK(20)
    sor.setStddevMulThresh(1.0)
    # fewer finite points than mean_k + 1: nothing is counted, the cut is NaN, everything stays — non-finite points too
    tiny = small_cloud(oracle, np.random.default_rng(8))[495:507]
    sor.setInputCloud(tiny)
    assert bits_equal(sor.filter(), tiny) and bits_equal(oracle.statistical_outlier_removal(tiny, 20, 1.0), tiny)
    st = sor.last_stats(len(tiny))
    assert st["valid"] == 0 and np.isnan(st["threshold"]) and not st["distances"].any()
    sor.setInputCloud(np.zeros((0, 4), np.float32))
    assert len(sor.filter()) == 0
    # isolated points (their 21st neighbour is beyond the ring search: block-per-query scan) among a dense patch
    rng = np.random.default_rng(10)
    dense = np.ones((4000, 4), np.float32)
    dense[:, :3] = rng.normal(0, 1.5, (4000, 3)).astype(np.float32)
    lonely = np.ones((25, 4), np.float32)
    lonely[:, :3] = rng.uniform(-150, 150, (25, 3)).astype(np.float32)
    c = np.concatenate([dense[:2000], lonely, dense[2000:]])
    want, det = oracle.statistical_outlier_removal(c, 20, 1.0, details=True)
    sor.setInputCloud(c)
    assert bits_equal(sor.filter(), want)
    assert bits_equal(sor.last_stats(len(c))["distances"], det["distances"])
    # exactly zero variance: no rounding bound exists, the index-order pass decides (and agrees to the bit)
    sor.setMeanK(1)
    line = line_cloud()
    sor.setInputCloud(line)
    want, det = oracle.statistical_outlier_removal(line, 1, 1.0, details=True)
    assert bits_equal(sor.filter(), want) and len(want) == 64
    st = sor.last_stats(64)
    assert st["exact_pass"] and (st["mean"], st["stddev"], st["threshold"]) == (det["mean"], det["stddev"], det["threshold"])
    # mean_k beyond the warp-wide neighbour list is refused loudly
    sor.setMeanK(32)
    with pytest.raises(eng.B200RegError):
        sor.filter()
    # a cloud too wide for the 0.5 m search lattice (> 2^31 cells) is refused by both filters, not filtered wrongly
    wide = np.concatenate([dense[:100], np.array([[1000, 1000, 1000, 1]], np.float32)])
    sor.setMeanK(20)
    sor.setInputCloud(wide)
    with pytest.raises(eng.B200RegError, match="lattice"):
        sor.filter()
    ror = eng.RadiusOutlierRemoval(registration=sor._reg)
    ror.setInputCloud(wide)
    with pytest.raises(eng.B200RegError, match="lattice"):
        ror.filter()
    sor.setInputCloud(c)  # and the handle carries on
    assert bits_equal(sor.filter(), oracle.statistical_outlier_removal(c, 20, 1.0))


@pytest.mark.gpu
def test_distance_filter_alone_and_the_chain_without_a_voxelgrid(oracle):
    """downsample_method NONE [REF apps/prefiltering_nodelet.cpp:70-75]: distance_filter is a call of its own
    (b200reg_distance_filter), host and device-resident, bit-exact and order-preserving; then the nodelet's default
    STATISTICAL filter on the gated cloud."""
    import torch
    import delta_graph_slam_b200 as eng
    rng = np.random.default_rng(5)
    c = cloud_with_outliers(oracle, rng)
    reg = eng.Registration()
    for near, far in ((1.0, 100.0), (0.1, 100.0), (5.0, 20.0), (200.0, 300.0)):
        want = oracle.distance_filter(c, near, far)
        assert bits_equal(reg.distance_filter(c, near, far), want)
        t = torch.from_numpy(c).cuda()
        o = torch.empty_like(t)
        f = reg.distance_filter(eng.DeviceCloud(t.data_ptr(), len(t), t), near, far, out=eng.DeviceCloud(o.data_ptr(), len(o), o))
        assert f.n == len(want) and bits_equal(o[: f.n].cpu().numpy(), want)
    assert len(reg.distance_filter(np.zeros((0, 4), np.float32))) == 0
    pinned = torch.empty((len(c), 4), dtype=torch.float32, pin_memory=True).numpy()
    assert bits_equal(np.array(reg.distance_filter(c, 1.0, 100.0, out=pinned)), oracle.distance_filter(c, 1.0, 100.0))
    small = c[::3].copy()
    pre = eng.Prefilter(dict(downsample_method="NONE"), out=DEVNULL)  # defaults: gate 1 .. 100, STATISTICAL 20 - 1
    assert bits_equal(pre.filter3d(small), oracle.statistical_outlier_removal(oracle.distance_filter(small, 1.0, 100.0), 20, 1.0))
    pre = eng.Prefilter(dict(downsample_method="NONE", outlier_removal_method="NONE", distance_near_thresh=5.0, distance_far_thresh=20.0), out=DEVNULL)
    assert bits_equal(pre.filter3d(small), oracle.distance_filter(small, 5.0, 20.0))
    # the reference's defaults end to end: gate fused into the VoxelGrid, then STATISTICAL
    pre = eng.Prefilter({}, out=DEVNULL)
    want = oracle.statistical_outlier_removal(oracle.voxelgrid(oracle.distance_filter(c, 1.0, 100.0), 0.1, is_dense=False)["out"], 20, 1.0)
    assert bits_equal(pre.filter3d(c), want)


@pytest.mark.gpu
def test_nodelet_default_chain_statistical(oracle, capsys):
    """The nodelet's own defaults [REF apps/prefiltering_nodelet.cpp:77-80]: STATISTICAL 20 - 1, through filter3d and the three-stage front end."""
    import torch
    import delta_graph_slam_b200 as eng
    from test_gpu_frontend import ODOM
    import sys
    params = dict(LAUNCH, outlier_removal_method="STATISTICAL")
    eng.Prefilter(params, out=sys.stdout)
    assert "outlier_removal: STATISTICAL 20 - 1" in capsys.readouterr().out
    clouds = [oracle.synth_scan(oracle.synth_traj(k), noise_seed=1000 + k) for k in range(4)]
    want = [oracle.statistical_outlier_removal(oracle.voxelgrid(oracle.distance_filter(c, 0.1, 100.0), 0.1, is_dense=False)["out"], 20, 1.0) for c in clouds]
    cap = max(len(c) for c in clouds)
    h_a = torch.empty((3, cap, 4), dtype=torch.float32, pin_memory=True).numpy()
    h_b = torch.empty((3, cap, 4), dtype=torch.float32, pin_memory=True).numpy()
    bufs, rbufs = [h_a[j] for j in range(3)], [h_b[j] for j in range(3)]
    pre = eng.Prefilter(params, out=DEVNULL)
    for k, c in enumerate(clouds):
        assert bits_equal(np.array(pre.filter3d(c, out=bufs[k % 3], out2=rbufs[k % 3])), want[k]), k

    def make():
        p, o = eng.Prefilter(params, out=DEVNULL), eng.ScanMatchingOdometry(ODOM, out=DEVNULL)
        p.setSmBudget(40)
        o.registration.setSmBudget(108)
        return p, o
    p1, o1 = make()
    seq = [o1.matching(0.1 * k, p1.filter3d(c, out=bufs[k % 3], out2=rbufs[k % 3])) for k, c in enumerate(clouds)]
    p2, o2 = make()
    seen = []
    got = eng.FrontEnd(p2, o2, bufs, filter_sms=40, ror_bufs=rbufs).run(clouds, on_frame=lambda k, f: seen.append(np.array(f).copy()))
    assert all(np.array_equal(a, b) for a, b in zip(got, seq))
    assert all(bits_equal(a, b) for a, b in zip(seen, want))


@pytest.mark.gpu
@pytest.mark.parametrize("where", ["host", "device"])
def test_launch_file_chain_and_three_stage_front_end(oracle, where):
    import torch
    import delta_graph_slam_b200 as eng
    from test_gpu_frontend import ODOM
    clouds = [oracle.synth_scan(oracle.synth_traj(k), noise_seed=1000 + k) for k in range(5)]
    want = [oracle.radius_outlier_removal(oracle.voxelgrid(oracle.distance_filter(c, 0.1, 100.0), 0.1, is_dense=False)["out"], 0.5, 2) for c in clouds]
    cap = max(len(c) for c in clouds)
    if where == "device":
        d_in = [torch.from_numpy(c).cuda() for c in clouds]
        inputs = [eng.DeviceCloud(t.data_ptr(), len(t), t) for t in d_in]
        d_a, d_b = torch.empty((3, cap, 4), dtype=torch.float32, device="cuda"), torch.empty((3, cap, 4), dtype=torch.float32, device="cuda")
        bufs = [eng.DeviceCloud(d_a[j].data_ptr(), cap, d_a) for j in range(3)]
        rbufs = [eng.DeviceCloud(d_b[j].data_ptr(), cap, d_b) for j in range(3)]
        to_np = lambda f: torch.empty((f.n, 4), dtype=torch.float32).copy_(d_b.view(-1, 4)[(f.ptr - d_b.data_ptr()) // 16:(f.ptr - d_b.data_ptr()) // 16 + f.n]).numpy()
    else:
        inputs = clouds
        h_a = torch.empty((3, cap, 4), dtype=torch.float32, pin_memory=True).numpy()
        h_b = torch.empty((3, cap, 4), dtype=torch.float32, pin_memory=True).numpy()
        bufs, rbufs = [h_a[j] for j in range(3)], [h_b[j] for j in range(3)]
        to_np = lambda f: np.array(f)
    pre = eng.Prefilter(LAUNCH, out=DEVNULL)
    for k, c in enumerate(inputs):
        f = pre.filter3d(c, out=bufs[k % 3], out2=rbufs[k % 3])
        assert bits_equal(to_np(f), want[k]), k

    def make():
        p, o = eng.Prefilter(LAUNCH, out=DEVNULL), eng.ScanMatchingOdometry(ODOM, out=DEVNULL)
        p.setSmBudget(40)
        o.registration.setSmBudget(108)
        return p, o
    p1, o1 = make()
    seq = [o1.matching(0.1 * k, p1.filter3d(c, out=bufs[k % 3], out2=rbufs[k % 3])) for k, c in enumerate(inputs)]
    p2, o2 = make()
    seen = []
    got = eng.FrontEnd(p2, o2, bufs, filter_sms=40, ror_bufs=rbufs).run(inputs, on_frame=lambda k, f: seen.append(to_np(f).copy()))
    assert all(np.array_equal(a, b) for a, b in zip(got, seq))
    assert all(bits_equal(a, b) for a, b in zip(seen, want))
