"""Host logic of the front-end mirrors, no GPU: the keyframe state machine of
ScanMatchingOdometryNodelet::matching [REF apps/scan_matching_odometry_nodelet.cpp:173-270] on a scripted
registration object, and the call order of the pipelined FrontEnd on scripted filters."""
import math
import os

import numpy as np
import pytest

from delta_graph_slam_b200.odometry import FrontEnd, ScanMatchingOdometry, quaternion_w

DEVNULL = open(os.devnull, "w")


def pose(x=0.0, y=0.0, yaw=0.0):
    T = np.eye(4, dtype=np.float32)
    T[0, 0], T[0, 1], T[1, 0], T[1, 1] = math.cos(yaw), -math.sin(yaw), math.sin(yaw), math.cos(yaw)
    T[0, 3], T[1, 3] = x, y
    return T


class ScriptedRegistration:
    """Returns the scripted transforms in order; records every call like a PCL registration would see them."""

    def __init__(self, transforms, converged=None):
        self.transforms = list(transforms)
        self.converged = list(converged) if converged is not None else [True] * len(self.transforms)
        self.calls = []
        self.k = -1
        self.prepared = 0

    def setInputTarget(self, cloud):
        self.calls.append(("target", id(cloud)))

    def setInputSource(self, cloud):
        self.calls.append(("source", id(cloud)))

    def align(self, guess):
        self.k += 1
        self.calls.append(("align", np.array(guess, np.float32)))

    def hasConverged(self):
        return self.converged[self.k]

    def getFinalTransformation(self):
        return self.transforms[self.k]

    def preparePromotion(self):
        self.prepared += 1
        self.calls.append(("prepare",))


def make(transforms, converged=None, **params):
    p = dict(keyframe_delta_trans=1.0, keyframe_delta_angle=1.0, keyframe_delta_time=10000.0, downsample_method="NONE")
    p.update(params)
    reg = ScriptedRegistration(transforms, converged)
    return ScanMatchingOdometry(p, registration=reg, out=DEVNULL), reg


def clouds(n):
    return [np.full((4, 4), float(k), np.float32) for k in range(n)]


def test_quaternion_w_matches_the_half_angle():
    for yaw in (0.0, 0.3, 1.5, 3.0, -2.5):
        assert abs(abs(float(quaternion_w(pose(yaw=yaw)[:3, :3]))) - abs(math.cos(yaw / 2))) < 1e-6


def test_first_frame_sets_the_keyframe_and_returns_identity():
    odo, reg = make([])
    c = clouds(1)
    assert np.array_equal(odo.matching(0.0, c[0]), np.eye(4, dtype=np.float32))
    assert [x[0] for x in reg.calls] == ["target"] and odo.num_keyframes == 1


def test_keyframe_switch_follows_the_translation_threshold():
    steps = [pose(x=0.4), pose(x=0.8), pose(x=1.2), pose(x=0.5)]  # relative to the current keyframe
    odo, reg = make(steps)
    c = clouds(5)
    out = [odo.matching(0.1 * k, c[k]) for k in range(5)]
    assert np.allclose(out[1][:3, 3], [0.4, 0, 0]) and np.allclose(out[2][:3, 3], [0.8, 0, 0])
    # frame 3 moved 1.2 m > keyframe_delta_trans: it becomes the keyframe (promoted as target), the guess restarts at identity
    assert np.allclose(out[3][:3, 3], [1.2, 0, 0]) and odo.num_keyframes == 2
    kinds = [x[0] for x in reg.calls]
    assert kinds == ["target", "source", "align", "source", "align", "source", "align", "target", "source", "align"]
    assert reg.calls[7][1] == reg.calls[5][1]          # setInputTarget(keyframe) with the cloud just used as source
    assert np.array_equal(reg.calls[9][1], np.eye(4))  # guess after the switch
    assert np.allclose(reg.calls[6][1], steps[1])      # guess = previous result while the keyframe stays
    assert np.allclose(out[4][:3, 3], [1.7, 0, 0])     # keyframe pose 1.2 + 0.5


def test_rotation_and_time_thresholds_also_switch():
    # delta_angle = acos(Quaternionf(R).w()) is HALF the rotation angle [REF apps/scan_matching_odometry_nodelet.cpp:250]
    c = clouds(2)
    odo, _ = make([pose(yaw=1.8)], keyframe_delta_angle=1.0)
    odo.matching(0.0, c[0]); odo.matching(0.1, c[1])
    assert odo.num_keyframes == 1
    odo, _ = make([pose(yaw=2.4)], keyframe_delta_angle=1.0)
    odo.matching(0.0, c[0]); odo.matching(0.1, c[1])
    assert odo.num_keyframes == 2
    odo, _ = make([pose(x=0.1)], keyframe_delta_time=1.0)
    odo.matching(0.0, c[0]); odo.matching(5.0, c[1])
    assert odo.num_keyframes == 2


def test_a_frame_that_did_not_converge_leaves_the_state_alone():
    odo, reg = make([pose(x=0.3), pose(x=9.0), pose(x=0.6)], converged=[True, False, True])
    c = clouds(4)
    out = [odo.matching(0.1 * k, c[k]) for k in range(4)]
    assert np.allclose(out[2], out[1])                   # "scan matching has not converged!! ignore this frame"
    assert np.allclose(reg.calls[-1][1], pose(x=0.3))    # the next guess is still the last accepted result
    assert np.allclose(out[3][:3, 3], [0.6, 0, 0]) and odo.num_keyframes == 1


def test_transform_thresholding_rejects_jumps():
    odo, _ = make([pose(x=0.3), pose(x=2.5), pose(x=0.5)], transform_thresholding=True, max_acceptable_trans=1.0, max_acceptable_angle=1.0)
    c = clouds(4)
    out = [odo.matching(0.1 * k, c[k]) for k in range(4)]
    assert np.allclose(out[2], out[1]) and odo.num_keyframes == 1
    assert np.allclose(out[3][:3, 3], [0.5, 0, 0])


def test_promotion_hint_fires_when_the_next_step_would_cross_the_threshold():
    steps = [pose(x=0.5), pose(x=1.05), pose(x=0.5), pose(x=0.98), pose(x=1.4)]
    odo, reg = make(steps, prepare_promotion=True)
    c = clouds(6)
    for k in range(6):
        odo.matching(0.1 * k, c[k])
    # hints: frame 2 (0.5 + 0.5 > 0.95), frame 4 (0.5 + 0.55 > 0.95), frame 5 (0.98 + 0.48 > 0.95); not frames 1 and 3
    assert reg.prepared == 3 and odo.promotions_prepared == 3
    kinds = [x[0] for x in reg.calls]
    assert all(kinds[i + 1] == "align" for i, k in enumerate(kinds) if k == "prepare")
    assert odo.num_keyframes == 3


class ScriptedPrefilter:
    """Filter stages that record when they are started and collected."""

    def __init__(self, log, with_ror):
        self.log = log
        self.filter = object()
        self.outlier_removal_filter = object() if with_ror else None
        self._vg = self._ror = None

    def setSmBudget(self, n):
        self.log.append(("budget", n))

    def downsample_begin(self, cloud, out):
        self.log.append(("vg_begin", int(cloud[0, 0])))
        self._vg = cloud

    def downsample_end(self):
        self.log.append(("vg_end", int(self._vg[0, 0])))
        return self._vg

    def outlier_removal_begin(self, cloud, out):
        self.log.append(("ror_begin", int(cloud[0, 0])))
        self._ror = cloud

    def outlier_removal_end(self):
        self.log.append(("ror_end", int(self._ror[0, 0])))
        return self._ror


class LoggingOdometry:
    def __init__(self, log):
        self.log = log
        self.registration = self
        self.prepare_promotion = False

    def setSmBudget(self, n):
        self.log.append(("reg_budget", n))

    def matching(self, stamp, cloud):
        self.log.append(("match", int(cloud[0, 0])))
        return int(cloud[0, 0])


def test_front_end_keeps_the_next_scan_in_flight():
    log = []
    fe = FrontEnd(ScriptedPrefilter(log, False), LoggingOdometry(log), [None] * 3, filter_sms=40)
    assert fe.run(clouds(3)) == [0, 1, 2]
    assert log[:2] == [("budget", 40), ("reg_budget", 108)]
    assert log[2:] == [("vg_begin", 0), ("vg_end", 0), ("vg_begin", 1), ("match", 0), ("vg_end", 1), ("vg_begin", 2), ("match", 1), ("vg_end", 2), ("match", 2)]
    with pytest.raises(ValueError):
        FrontEnd(ScriptedPrefilter([], False), LoggingOdometry([]), [None] * 2)


def test_three_stage_front_end_runs_two_scans_ahead():
    log = []
    fe = FrontEnd(ScriptedPrefilter(log, True), LoggingOdometry(log), [None] * 3, filter_sms=40, ror_bufs=[None] * 3)
    assert fe.run(clouds(4)) == [0, 1, 2, 3]
    seq = log[2:]
    # every scan goes filter -> outlier removal -> matching, in scan order per stage
    for stage in ("vg_begin", "vg_end", "ror_begin", "ror_end", "match"):
        assert [k for s, k in seq if s == stage] == [0, 1, 2, 3]
    pos = {e: i for i, e in enumerate(seq)}
    for k in range(4):
        assert pos[("vg_begin", k)] < pos[("vg_end", k)] < pos[("ror_begin", k)] < pos[("ror_end", k)] < pos[("match", k)]
    # while scan k is matched, the outlier removal of k+1 and the filter of k+2 have been started
    for k in range(2):
        assert pos[("ror_begin", k + 1)] < pos[("match", k)] and pos[("vg_begin", k + 2)] < pos[("match", k)]
    with pytest.raises(ValueError):
        FrontEnd(ScriptedPrefilter([], True), LoggingOdometry([]), [None] * 3)
