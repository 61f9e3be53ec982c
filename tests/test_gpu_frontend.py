"""The filter as begin / end halves, SM budgets and the pipelined front end (eng.FrontEnd) — the
reference's prefiltering_nodelet -> /filtered_points -> scan_matching_odometry_nodelet chain
[REF launch/delta_graph_slam.launch:26,46; apps/prefiltering_nodelet.cpp:48,51; apps/scan_matching_odometry_nodelet.cpp:53]."""
import os

import numpy as np
import pytest

from helpers import bits_equal

pytestmark = pytest.mark.gpu
DEVNULL = open(os.devnull, "w")
ODOM = dict(keyframe_delta_trans=1.0, keyframe_delta_angle=1.0, keyframe_delta_time=10000.0, downsample_method="NONE", registration_method="NDT_OMP", reg_resolution=1.0,
            reg_nn_search_method="DIRECT7", reg_transformation_epsilon=0.01, reg_maximum_iterations=64)
# a plain VoxelGrid: the nodelet's outlier filter and distance gate have tests of their own (test_prefilter_chain.py)
PRE = dict(downsample_method="VOXELGRID", downsample_resolution=0.1, outlier_removal_method="NONE", b200_skip_distance_filter=True)


@pytest.fixture(scope="module")
def scans():
    from oracle import oracle_py as O
    return O, [O.synth_scan(O.synth_traj(k), noise_seed=1000 + k) for k in range(6)]


def test_filter_begin_end_equals_filter_and_oracle(scans):
    import delta_graph_slam_b200 as eng
    O, clouds = scans
    vg = eng.VoxelGrid()
    vg.setLeafSize(0.1, 0.1, 0.1)
    for c in clouds[:3]:
        ref = O.voxelgrid(c, 0.1, is_dense=False)["out"]
        vg.setInputCloud(c, is_dense=False)
        sync = vg.filter()
        out = np.zeros((len(c), 4), np.float32)  # pageable: staged in _end
        vg.filter_begin(out)
        got = vg.filter_end()
        assert bits_equal(sync, ref) and bits_equal(got, ref)


def test_filter_zero_copy_into_pinned_output(scans):
    import torch
    import delta_graph_slam_b200 as eng
    O, clouds = scans
    c = clouds[0]
    ref = O.voxelgrid(c, 0.1, is_dense=False)["out"]
    h_in = torch.empty((len(c), 4), dtype=torch.float32, pin_memory=True)
    h_in.numpy()[:] = c
    h_out = torch.zeros((len(c), 4), dtype=torch.float32, pin_memory=True)
    vg = eng.VoxelGrid()
    vg.setLeafSize(0.1, 0.1, 0.1)
    vg.setInputCloud(h_in.numpy(), is_dense=False)
    vg.filter_begin(h_out.numpy())
    got = vg.filter_end()
    assert bits_equal(got, ref)
    assert not h_out.numpy()[len(ref):].any(), "the kernel wrote past the filtered cloud"
    # capacity smaller than the result: error, nothing written past the capacity
    small = torch.zeros((1000, 4), dtype=torch.float32, pin_memory=True)
    guard = small.numpy()
    with pytest.raises(ValueError):
        vg.filter_begin(guard)  # the Python mirror insists on room for the whole input, like pcl::Filter's output cloud
    L = eng._lib.load()
    import ctypes as C
    leaf = (C.c_float * 3)(0.1, 0.1, 0.1)
    rc = L.b200reg_voxelgrid_filter_begin(vg._reg._h, h_in.numpy().ctypes.data, len(c), 16, leaf, 0, 0, guard.ctypes.data, 1000)
    assert rc == 0
    n = C.c_size_t()
    assert L.b200reg_voxelgrid_filter_end(vg._reg._h, C.byref(n)) == eng._lib.E_CAPACITY
    assert n.value == len(ref)
    assert bits_equal(guard, ref[:1000])


def test_second_begin_and_stray_end_are_state_errors(scans):
    import delta_graph_slam_b200 as eng
    _, clouds = scans
    vg = eng.VoxelGrid()
    vg.setLeafSize(0.1, 0.1, 0.1)
    vg.setInputCloud(clouds[0], is_dense=False)
    out = np.zeros((len(clouds[0]), 4), np.float32)
    vg.filter_begin(out)
    with pytest.raises(eng.B200RegError) as e:
        vg.filter_begin(out)
    assert e.value.code == eng._lib.E_STATE
    vg.filter_end()
    import ctypes as C
    n = C.c_size_t()
    assert eng._lib.load().b200reg_voxelgrid_filter_end(vg._reg._h, C.byref(n)) == eng._lib.E_STATE


@pytest.mark.parametrize("budget", [1, 7, 40, 148])
def test_filter_is_bit_exact_under_any_sm_budget(scans, budget):
    import delta_graph_slam_b200 as eng
    O, clouds = scans
    ref = O.voxelgrid(clouds[1], 0.1, is_dense=False)["out"]
    vg = eng.VoxelGrid()
    vg.setSmBudget(budget)
    vg.setLeafSize(0.1, 0.1, 0.1)
    vg.setInputCloud(clouds[1], is_dense=False)
    assert bits_equal(vg.filter(), ref)


def test_registration_under_sm_budget_matches_oracle(scans):
    import delta_graph_slam_b200 as eng
    from helpers import transform_delta
    O, clouds = scans
    t = O.voxelgrid(clouds[0], 0.1)["out"]
    s = O.voxelgrid(clouds[1], 0.1)["out"]
    ref = O.Registration(O.NDT, resolution=1.0, nn_search=O.DIRECT7, trans_eps=0.01, max_iter=64)
    ref.setInputTarget(t); ref.setInputSource(s); ref.align(np.eye(4, dtype=np.float32))
    for budget in (108, 37):
        ndt = eng.select_registration_method(dict(registration_method="NDT_OMP", reg_resolution=1.0, reg_nn_search_method="DIRECT7"), out=DEVNULL)
        ndt.setSmBudget(budget)
        ndt.setInputTarget(t); ndt.setInputSource(s); ndt.align(np.eye(4, dtype=np.float32))
        dt, dr = transform_delta(ndt.getFinalTransformation(), ref.getFinalTransformation())
        assert dt < 1e-4 and dr < 1e-4  # tolerance of BASELINE.json's north_star
        assert ndt.getFinalNumIteration() == ref.getFinalNumIteration()
        f1, f0 = ndt.getFitnessScore(), ref.getFitnessScore()
        assert abs(f1 - f0) <= 1e-5 * abs(f0)


@pytest.mark.parametrize("where", ["host", "device"])
def test_pipelined_front_end_equals_the_plain_loop(scans, where):
    import torch
    import delta_graph_slam_b200 as eng
    _, clouds = scans
    cap = max(len(c) for c in clouds)
    if where == "device":
        d_in = [torch.from_numpy(c).cuda() for c in clouds]
        inputs = [eng.DeviceCloud(t.data_ptr(), len(t), t) for t in d_in]
        d_out = torch.empty((3, cap, 4), dtype=torch.float32, device="cuda")
        bufs = [eng.DeviceCloud(d_out[j].data_ptr(), cap, d_out) for j in range(3)]
    else:
        inputs = clouds
        h_out = torch.empty((3, cap, 4), dtype=torch.float32, pin_memory=True).numpy()
        bufs = [h_out[j] for j in range(3)]

    def make():
        pre = eng.Prefilter(PRE, out=DEVNULL)
        odo = eng.ScanMatchingOdometry(ODOM, out=DEVNULL)
        pre.filter.setSmBudget(40)
        odo.registration.setSmBudget(108)
        return pre, odo
    pre, odo = make()
    seq = [odo.matching(0.1 * k, pre.downsample(c, out=bufs[k % 3])) for k, c in enumerate(inputs)]
    pre2, odo2 = make()
    fe = eng.FrontEnd(pre2, odo2, bufs, filter_sms=40)
    got = fe.run(inputs)
    assert len(got) == len(seq) and all(np.array_equal(a, b) for a, b in zip(got, seq))
    assert odo2.num_keyframes == odo.num_keyframes


def test_prepared_promotion_changes_no_pose(scans):
    """b200reg_prepare_promotion: the next keyframe's NDT grid built on a side stream during its own
    registration; wrong guesses (prepared, not promoted) and unprepared promotions included."""
    import delta_graph_slam_b200 as eng
    O, _ = scans
    clouds = [O.voxelgrid(O.synth_scan(O.synth_traj(k), noise_seed=1000 + k), 0.1)["out"] for k in range(9)]

    def run(prepare, hint_always=False):
        odo = eng.ScanMatchingOdometry(dict(ODOM, prepare_promotion=prepare), out=DEVNULL)
        odo.registration.setSmBudget(108)
        odo.registration.setSideBudget(16)
        poses = []
        for k, c in enumerate(clouds):
            if hint_always and k > 0:
                odo._last_step = 10.0  # force a hint on every frame: most of them will not come true
            poses.append(odo.matching(0.1 * k, c))
        return poses, odo
    base, o0 = run(False)
    got, o1 = run(True)
    forced, o2 = run(True, hint_always=True)
    assert o0.num_keyframes >= 3 and o1.num_keyframes == o0.num_keyframes == o2.num_keyframes
    assert 0 < o1.promotions_prepared < len(clouds) and o2.promotions_prepared == len(clouds) - 1
    assert all(np.array_equal(a, b) for a, b in zip(base, got))
    assert all(np.array_equal(a, b) for a, b in zip(base, forced))
    # the hint is a state error without a source, and a no-op on a FAST_GICP handle
    ndt = eng.select_registration_method(dict(registration_method="NDT_OMP"), out=DEVNULL)
    with pytest.raises(eng.B200RegError) as e:
        ndt.preparePromotion()
    assert e.value.code == eng._lib.E_STATE
    eng.select_registration_method(dict(registration_method="FAST_GICP"), out=DEVNULL).preparePromotion()


def test_odometry_downsamples_device_scans_with_its_own_voxelgrid(scans):
    """downsample_method VOXELGRID in the odometry nodelet [REF apps/scan_matching_odometry_nodelet.cpp:85-89,155-165]
    on device-resident scans: the filter's outputs rotate through caller-owned device buffers; without them the
    mirror says so instead of crashing."""
    import torch
    import delta_graph_slam_b200 as eng
    _, clouds = scans
    params = dict(ODOM, downsample_method="VOXELGRID", downsample_resolution=0.2)
    host = eng.ScanMatchingOdometry(params, out=DEVNULL)
    want = [host.matching(0.1 * k, c) for k, c in enumerate(clouds[:4])]
    cap = max(len(c) for c in clouds)
    d_in = [torch.from_numpy(c).cuda() for c in clouds[:4]]
    d_out = torch.empty((3, cap, 4), dtype=torch.float32, device="cuda")
    bufs = [eng.DeviceCloud(d_out[j].data_ptr(), cap, d_out) for j in range(3)]
    dev = eng.ScanMatchingOdometry(params, out=DEVNULL, downsample_bufs=bufs)
    got = [dev.matching(0.1 * k, eng.DeviceCloud(t.data_ptr(), len(t), t)) for k, t in enumerate(d_in)]
    assert all(np.array_equal(a, b) for a, b in zip(got, want))
    bare = eng.ScanMatchingOdometry(params, out=DEVNULL)
    with pytest.raises(ValueError):
        bare.matching(0.0, eng.DeviceCloud(d_in[0].data_ptr(), len(d_in[0]), d_in[0]))
    vg = eng.VoxelGrid()
    vg.setLeafSize(0.2, 0.2, 0.2)
    vg.setInputCloud(eng.DeviceCloud(d_in[0].data_ptr(), len(d_in[0]), d_in[0]))
    with pytest.raises(ValueError):
        vg.filter()
