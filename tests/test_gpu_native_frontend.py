"""b200reg_odometry_* / b200reg_frontend_* (csrc/b200reg_odometry.cu): the two front-end nodelets as host C++ above the
engine, against the Python mirror of the same state machine (delta_graph_slam_b200.odometry) that the oracle parity tests
drive — ScanMatchingOdometryNodelet::matching [REF apps/scan_matching_odometry_nodelet.cpp:173-270] and the prefilter
pipeline [REF launch/delta_graph_slam.launch:26,46]."""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import transform_delta

pytestmark = pytest.mark.gpu
DEVNULL = open(os.devnull, "w")
PARAMS = dict(keyframe_delta_trans=1.0, keyframe_delta_angle=1.0, keyframe_delta_time=10000.0, registration_method="NDT_OMP", reg_resolution=1.0, reg_nn_search_method="DIRECT7",
              reg_transformation_epsilon=0.01, reg_maximum_iterations=64, downsample_method="VOXELGRID", downsample_resolution=0.1, distance_near_thresh=0.1, distance_far_thresh=100.0,
              outlier_removal_method="NONE")
FRAMES = 14


@pytest.fixture(scope="module")
def scans():
    from oracle import oracle_py as O
    return [O.synth_scan(O.synth_traj(k), noise_seed=1000 + k) for k in range(FRAMES)]


@pytest.fixture(scope="module")
def python_mirror(scans):
    """The Python FrontEnd on the same SM budgets: poses and per-frame registration records."""
    import torch
    import delta_graph_slam_b200 as eng
    cap = max(len(c) for c in scans)
    h_out = torch.empty((3, cap, 4), dtype=torch.float32, pin_memory=True).numpy()
    pre = eng.Prefilter(PARAMS, out=DEVNULL)
    odo = eng.ScanMatchingOdometry(dict(PARAMS, downsample_method="NONE"), out=DEVNULL)
    fe = eng.FrontEnd(pre, odo, [h_out[j] for j in range(3)], filter_sms=40)
    recs = []
    poses = fe.run(scans, on_frame=lambda k, f: recs.append((odo.registration.getResult() if k else None, len(f), odo.num_keyframes)))
    return poses, recs


def check(poses, python_mirror, results=None, n_filtered=None):
    want_poses, recs = python_mirror
    for k in range(FRAMES):
        dt, dr = transform_delta(poses[k], want_poses[k])
        assert dt < 2e-6 and dr < 2e-6, f"frame {k}: accumulated pose {dt:.1e} m / {dr:.1e} rad from the Python mirror's (float 4x4 products in another order)"
        if k and results is not None:
            r = recs[k][0]
            assert np.array_equal(np.array(results[k]["transformation"]).reshape(4, 4).T, r["transformation"]), f"frame {k}: registration result bit-identical"
            assert results[k]["iterations"] == r["iterations"] and results[k]["evaluations"] == r["evaluations"] and results[k]["hits"] == r["hits"]
        if n_filtered is not None:
            assert int(n_filtered[k]) == recs[k][1]


def test_native_front_end_device_sequence_equals_the_python_mirror(scans, python_mirror):
    import torch
    import delta_graph_slam_b200 as eng
    d_in = [torch.from_numpy(c).cuda() for c in scans]
    clouds = [eng.DeviceCloud(t.data_ptr(), len(t), t) for t in d_in]
    fe = eng.NativeFrontEnd(PARAMS, filter_sms=40)
    poses, res, nf = fe.run_device(clouds)
    check(poses, python_mirror, res, nf)
    assert fe.num_keyframes() == python_mirror[1][-1][2] >= 4
    poses2, res2, _ = fe.run_device(clouds)   # a second pass over the sequence restarts cleanly and repeats itself bit for bit
    assert np.array_equal(poses, poses2) and np.array_equal(res.view(np.uint8), res2.view(np.uint8))


@pytest.mark.parametrize("form", ["two_nodelets_pinned", "two_nodelets_pageable", "fused"])
def test_native_front_end_host_scans(scans, python_mirror, form):
    import torch
    import delta_graph_slam_b200 as eng
    cap = max(len(c) for c in scans)
    fe = eng.NativeFrontEnd(PARAMS, filter_sms=40)
    aligned = np.zeros((cap, 4), np.float32)
    if form == "fused":
        bufs = None
    elif form == "two_nodelets_pinned":
        keep = torch.empty((3, cap, 4), dtype=torch.float32, pin_memory=True)
        bufs = [keep.numpy()[j] for j in range(3)]
    else:
        bufs = [np.zeros((cap, 4), np.float32) for _ in range(3)]
    poses = fe.run_host(scans, filtered_bufs=bufs, aligned_out=aligned)
    check(poses, python_mirror)
    # a second sequence through the same object starts from a clean state (first scan -> first keyframe): same poses again
    assert np.array_equal(fe.run_host(scans, filtered_bufs=bufs, aligned_out=aligned), poses)
    if bufs is not None:  # the last scan's filtered cloud reached the caller, and `aligned` is the final transform applied to it
        from oracle import oracle_py as O
        k = FRAMES - 1
        want = O.voxelgrid(O.distance_filter(scans[k], 0.1, 100.0), 0.1, is_dense=False)["out"]
        got = bufs[k % 3][: len(want)]
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
        # the last frame was aligned against the current keyframe: aligned = T * filtered
        st = fe._L.b200reg_frontend_odometry(fe._h)
        from delta_graph_slam_b200 import _lib
        last = _lib.Result()
        fe._L.b200reg_odometry_get_state(st, None, None, None, None, None, C.byref(last))
        T = np.array(last.transformation[:], np.float32).reshape(4, 4).T
        ref = (want[:, :3] @ T[:3, :3].T + T[:3, 3]).astype(np.float32)
        assert np.max(np.abs(aligned[: len(want), :3] - ref)) < 1e-4


def test_odometry_object_alone_follows_the_python_mirror(scans):
    """b200reg_odometry_matching on prefiltered host clouds, one call per scan (the odometry nodelet on its own)."""
    import delta_graph_slam_b200 as eng
    from delta_graph_slam_b200 import _lib
    from oracle import oracle_py as O
    L = _lib.load()
    filtered = [O.voxelgrid(c, 0.1, is_dense=False)["out"] for c in scans[:8]]
    odo_py = eng.ScanMatchingOdometry(dict(PARAMS, downsample_method="NONE"), out=DEVNULL)
    want = [odo_py.matching(0.1 * k, f) for k, f in enumerate(filtered)]
    reg = eng.select_registration_method(PARAMS, out=DEVNULL)
    cfg = _lib.OdometryConfig()
    L.b200reg_odometry_default_config(C.byref(cfg))
    assert (cfg.keyframe_delta_trans, cfg.keyframe_delta_angle, cfg.keyframe_delta_time, cfg.transform_thresholding) == (0.25, 0.15, 1.0, 0)
    cfg.keyframe_delta_trans, cfg.keyframe_delta_angle, cfg.keyframe_delta_time = 1.0, 1.0, 10000.0
    h = C.c_void_p()
    assert L.b200reg_odometry_create(reg._h, C.byref(cfg), C.byref(h)) == 0
    try:
        for k, f in enumerate(filtered):
            odom = np.zeros(16, np.float32)
            assert L.b200reg_odometry_matching(h, 0.1 * k, f.ctypes.data, len(f), 16, None, None, odom.ctypes.data) == 0
            dt, dr = transform_delta(odom.reshape(4, 4).T, want[k])
            assert dt < 2e-6 and dr < 2e-6, k
        n = C.c_int()
        L.b200reg_odometry_get_state(h, C.byref(n), None, None, None, None, None)
        assert n.value == odo_py.num_keyframes
        # an empty source is a state error inside align: the frame is ignored, the pose repeats (PCL logs and returns)
        odom = np.zeros(16, np.float32)
        assert L.b200reg_odometry_matching(h, 1.0, None, 0, 16, None, None, odom.ctypes.data) == 0
        assert L.b200reg_odometry_reset(h) == 0
        assert L.b200reg_odometry_matching(h, 0.0, filtered[0].ctypes.data, len(filtered[0]), 16, None, None, odom.ctypes.data) == 0
        assert np.array_equal(odom.reshape(4, 4), np.eye(4, dtype=np.float32))
    finally:
        L.b200reg_odometry_destroy(h)
