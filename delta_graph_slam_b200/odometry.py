"""Host-side mirror of ScanMatchingOdometryNodelet::matching and PrefilteringNodelet::downsample.

The frame-to-keyframe state machine of [REF apps/scan_matching_odometry_nodelet.cpp:173-270] and
its parameters [REF :65-106], with the registration and the down-sampling filters replaced by the
b200reg engine.  This is host logic only (O(1) per frame); every point-cloud operation is a C-ABI
call.  ROS plumbing (tf, msf / robot-odometry guesses, publishers) is out of scope: `msf_delta`
can be passed in by the caller and defaults to identity, which is what the reference uses when
`enable_imu_frontend` and `enable_robot_odometry_init_guess` are off (every launch file).
"""
import sys

import numpy as np

from .registration import DeviceCloud, RadiusOutlierRemoval, StatisticalOutlierRemoval, VoxelGrid, select_registration_method


def quaternion_w(R):
    """w of Eigen::Quaternionf(R) (Eigen's rotation-matrix -> quaternion conversion), float32."""
    R = np.asarray(R, np.float32)
    t = np.float32(R[0, 0] + R[1, 1] + R[2, 2])
    if t > 0:
        return np.float32(0.5) * np.sqrt(t + np.float32(1.0))
    i = 0
    if R[1, 1] > R[0, 0]:
        i = 1
    if R[2, 2] > R[i, i]:
        i = 2
    j, k = (i + 1) % 3, (i + 2) % 3
    t = np.sqrt(R[i, i] - R[j, j] - R[k, k] + np.float32(1.0))
    return (R[k, j] - R[j, k]) * (np.float32(0.5) / t)


class Prefilter:
    """PrefilteringNodelet's 3-D chain [REF apps/prefiltering_nodelet.cpp:150-153]: distance_filter (:275-291)
    -> downsample (:249-260, filter chosen :55-75) -> outlier_removal (:262-273, chosen :77-98), with the
    reference's parameter names AND defaults: down-sampling VOXELGRID 0.1, outlier removal STATISTICAL 20 / 1.0,
    distance gate 1.0 .. 100.0.  The gate runs on EVERY scan: the reference reads `use_distance_filter` (:100) and
    never tests it (cloud_callback calls distance_filter unconditionally, :150), so the parameter is accepted and
    ignored here as well.  The one opt-out is the mirror's own key `b200_skip_distance_filter` (not a reference
    parameter; for callers whose input has already been gated).  With a VoxelGrid the gate is fused into the filter's
    key pass; without one (downsample_method NONE) it is a call of its own (b200reg_distance_filter)."""

    def __init__(self, params=None, device=0, out=sys.stdout):
        p = dict(params or {})
        method = p.get("downsample_method", "VOXELGRID")
        res = p.get("downsample_resolution", 0.1)
        self.filter = None
        self._reg = None
        if method == "VOXELGRID":
            print(f"downsample: VOXELGRID {res:g}", file=out)
            self.filter = VoxelGrid(device=device)
            self.filter.setLeafSize(res, res, res)
            self._reg = self.filter._reg
        elif method == "APPROX_VOXELGRID":
            raise NotImplementedError("APPROX_VOXELGRID stays on the reference's pcl::ApproximateVoxelGrid (not on the B200 path)")
        else:
            if method != "NONE":
                print(f"warning: unknown downsampling type ({method})", file=sys.stderr)
                print("       : use passthrough filter", file=sys.stderr)
            print("downsample: NONE", file=out)
        orm = p.get("outlier_removal_method", "STATISTICAL")
        self.outlier_removal_filter = None
        if orm == "STATISTICAL":
            mean_k = p.get("statistical_mean_k", 20)
            stddev_mul_thresh = p.get("statistical_stddev", 1.0)
            print(f"outlier_removal: STATISTICAL {mean_k} - {stddev_mul_thresh:g}", file=out)
            self.outlier_removal_filter = StatisticalOutlierRemoval(device=device, registration=self._reg)
            self.outlier_removal_filter.setMeanK(mean_k)
            self.outlier_removal_filter.setStddevMulThresh(stddev_mul_thresh)
        elif orm == "RADIUS":
            radius = p.get("radius_radius", 0.8)
            min_neighbors = p.get("radius_min_neighbors", 2)
            print(f"outlier_removal: RADIUS {radius:g} - {min_neighbors}", file=out)
            # same handle (and stream) as the VoxelGrid: the stages of one scan run in order
            self.outlier_removal_filter = RadiusOutlierRemoval(device=device, registration=self._reg)
            self.outlier_removal_filter.setRadiusSearch(radius)
            self.outlier_removal_filter.setMinNeighborsInRadius(min_neighbors)
        else:
            print("outlier_removal: NONE", file=out)
        if self._reg is None and self.outlier_removal_filter is not None:
            self._reg = self.outlier_removal_filter._reg
        self.use_distance_filter = bool(p.get("use_distance_filter", True))  # read and never tested, as upstream
        self.distance_near_thresh = float(p.get("distance_near_thresh", 1.0))
        self.distance_far_thresh = float(p.get("distance_far_thresh", 100.0))
        self.distance_filter_on = not bool(p.get("b200_skip_distance_filter", False))
        if self.distance_filter_on:
            if self.filter is not None:
                self.filter.setDistanceFilter(True, self.distance_near_thresh, self.distance_far_thresh)
            elif self._reg is None:
                from .registration import Registration
                self._reg = Registration(device=device)  # a filter-only handle for the stand-alone gate

    def setBaseLinkTransform(self, transform):
        """The base_link step of cloud_callback [REF apps/prefiltering_nodelet.cpp:123-148].  `transform`: the 4x4
        sensor -> base_link transform the nodelet looks up from tf when `base_link_frame` is set (None: parameter empty,
        the reference's default — scans stay in the sensor frame).  As upstream the x / y translation is zeroed ("lidar
        scans should be centered in base_link"), the remaining translation is the lidar position handed to the height
        filter, and the cloud is transformed with the Matrix4d (double arithmetic, one rounding to float) in front of
        distance_filter — on the device, by the handle of the chain's first stage.  Returns the lidar position."""
        first = self.filter._reg if self.filter is not None else (self._reg if self.distance_filter_on else None)
        if transform is None:
            if first is not None:
                first.setInputTransform(None)
            self.lidar_position = np.zeros(3)
            return self.lidar_position
        if first is None:
            raise ValueError("base_link transform: this chain has neither a VoxelGrid nor the distance gate to attach it to (b200_skip_distance_filter with downsample_method NONE)")
        m = np.array(transform, np.float64).reshape(4, 4)
        m[0, 3] = 0.0
        m[1, 3] = 0.0
        self.lidar_position = m[:3, 3].copy()
        first.setInputTransform(m)
        return self.lidar_position

    def setSmBudget(self, n_sm):
        if self.filter is not None:
            self.filter.setSmBudget(n_sm)
        elif self.outlier_removal_filter is not None:
            self.outlier_removal_filter._reg.setSmBudget(n_sm)

    def distance_filter(self, cloud, out=None):
        """distance_filter as a stage of its own; with a VoxelGrid it is part of downsample() instead."""
        if not self.distance_filter_on or self.filter is not None:
            return cloud
        return self._reg.distance_filter(cloud, self.distance_near_thresh, self.distance_far_thresh, out=out)

    def downsample(self, cloud, out=None):
        if self.filter is None:
            return cloud
        # distance_filter hands over is_dense = false [REF apps/prefiltering_nodelet.cpp:286]
        self.filter.setInputCloud(cloud, is_dense=False)
        return self.filter.filter(out=out)

    def outlier_removal(self, cloud, out=None):
        if self.outlier_removal_filter is None:
            return cloud
        self.outlier_removal_filter.setInputCloud(cloud)
        return self.outlier_removal_filter.filter(out=out)

    def filter2d(self, filtered3d, lidar_z, out=None):
        """filtered2D of cloud_callback [REF apps/prefiltering_nodelet.cpp:155-158]: height_filtering (z > lidar z) ->
        normal_filtering (k = 10, |n_z| < 0.2) -> flatten, one engine call on the prefilter's handle."""
        if self._reg is None:
            from .registration import Registration
            self._reg = Registration(device=0)
        return self._reg.flat_filter(filtered3d, lidar_z, out=out)

    def filter3d(self, cloud, out=None, out2=None, out0=None):
        """filtered3D of cloud_callback: distance_filter -> downsample -> outlier_removal (`out0`: output of the
        stand-alone gate when there is no VoxelGrid)."""
        return self.outlier_removal(self.downsample(self.distance_filter(cloud, out=out0), out=out), out=out2)

    def downsample_begin(self, cloud, out):
        """downsample() split in two: enqueue the filter of `cloud` into the caller-owned `out` ..."""
        if self.filter is None:
            self._passthrough = self.distance_filter(cloud, out=out if self.distance_filter_on else None)
            return
        self.filter.setInputCloud(cloud, is_dense=False)
        self.filter.filter_begin(out)

    def downsample_end(self):
        """... and collect the filtered cloud."""
        if self.filter is None:
            return self._passthrough
        return self.filter.filter_end()

    def outlier_removal_begin(self, cloud, out):
        self.outlier_removal_filter.setInputCloud(cloud)
        self.outlier_removal_filter.filter_begin(out)

    def outlier_removal_end(self):
        return self.outlier_removal_filter.filter_end()


class FrontEnd:
    """prefiltering_nodelet -> /filtered_points -> scan_matching_odometry_nodelet as the pipeline it is in
    the reference: the two nodelets are loaded into one nodelet manager and joined by a topic
    [REF launch/delta_graph_slam.launch:26,46; apps/prefiltering_nodelet.cpp:48,51;
    apps/scan_matching_odometry_nodelet.cpp:53], so scan k+1 is being down-sampled while scan k is matched.

    Here the filter of scan k+1 is enqueued on the prefilter handle's stream before matching(k) starts and
    collected after it returns.  `out_bufs`: at least three caller-owned output clouds used in rotation
    (the odometry holds the keyframe's cloud by reference while the next two scans are filtered).
    `filter_sms`: SMs left to the filter handle; the registration handle's persistent kernel takes the
    rest, so both are resident together.  The poses are those of the sequential loop
    (`for cloud: matching(stamp, downsample(cloud))`) run with the same SM budgets."""

    def __init__(self, prefilter, odometry, out_bufs, filter_sms=52, total_sms=148, ror_bufs=None, side_sms=16):
        if len(out_bufs) < 3:
            raise ValueError("the pipelined front end needs three output clouds in rotation")
        self.prefilter, self.odometry, self.out_bufs = prefilter, odometry, list(out_bufs)
        self.ror_bufs = list(ror_bufs) if ror_bufs is not None else None
        if prefilter.outlier_removal_filter is not None and (self.ror_bufs is None or len(self.ror_bufs) < 3):
            raise ValueError("a prefilter with outlier removal needs three more output clouds (ror_bufs)")
        if filter_sms:
            # persistent kernels that may be resident together: registration, filter (+ outlier removal's
            # sort, same stream) and, with prepared promotions, the side build of the next target
            side = side_sms if getattr(odometry, "prepare_promotion", False) and hasattr(odometry.registration, "setSideBudget") else 0
            prefilter.setSmBudget(filter_sms - side)
            odometry.registration.setSmBudget(total_sms - filter_sms)
            if side:
                odometry.registration.setSideBudget(side)

    def run(self, clouds, stamps=None, on_frame=None):
        pre, odo, bufs = self.prefilter, self.odometry, self.out_bufs
        poses = []
        n = len(clouds)
        if n == 0:
            return poses
        stamp = (lambda k: 0.1 * k) if stamps is None else (lambda k: stamps[k])
        if pre.outlier_removal_filter is None:
            pre.downsample_begin(clouds[0], bufs[0])
            for k in range(n):
                filtered = pre.downsample_end()
                if k + 1 < n:
                    pre.downsample_begin(clouds[k + 1], bufs[(k + 1) % len(bufs)])
                poses.append(odo.matching(stamp(k), filtered))
                if on_frame is not None:
                    on_frame(k, filtered)
            return poses
        # three stages (filter, outlier removal, matching), two scans ahead: while scan k is matched the
        # outlier removal of scan k+1 and then the filter of scan k+2 run on the prefilter handle's stream
        rb = self.ror_bufs
        pre.downsample_begin(clouds[0], bufs[0])
        pre.outlier_removal_begin(pre.downsample_end(), rb[0])
        if n > 1:
            pre.downsample_begin(clouds[1], bufs[1 % len(bufs)])
        cur = pre.outlier_removal_end()
        for k in range(n):
            if k + 1 < n:
                pre.outlier_removal_begin(pre.downsample_end(), rb[(k + 1) % len(rb)])
                if k + 2 < n:
                    pre.downsample_begin(clouds[k + 2], bufs[(k + 2) % len(bufs)])
            poses.append(odo.matching(stamp(k), cur))
            if on_frame is not None:
                on_frame(k, cur)
            if k + 1 < n:
                cur = pre.outlier_removal_end()
        return poses


class ScanMatchingOdometry:
    """matching(stamp, cloud) -> odom (4x4 float32), state as in the nodelet."""

    def __init__(self, params=None, device=0, out=sys.stdout, registration=None, downsample_filter=None, downsample_bufs=None):
        """`downsample_bufs`: for device-resident input with a VoxelGrid down-sampler, at least three caller-owned
        DeviceClouds used in rotation as the filter's outputs (the keyframe's cloud stays referenced while the next
        scans are filtered); host input needs none."""
        p = dict(params or {})
        self.downsample_bufs = list(downsample_bufs) if downsample_bufs is not None else None
        self._ds_next = 0
        self.keyframe_delta_trans = p.get("keyframe_delta_trans", 0.25)
        self.keyframe_delta_angle = p.get("keyframe_delta_angle", 0.15)
        self.keyframe_delta_time = p.get("keyframe_delta_time", 1.0)
        self.transform_thresholding = p.get("transform_thresholding", False)
        self.max_acceptable_trans = p.get("max_acceptable_trans", 1.0)
        self.max_acceptable_angle = p.get("max_acceptable_angle", 1.0)
        method = p.get("downsample_method", "VOXELGRID")
        res = p.get("downsample_resolution", 0.1)
        self.downsample_filter = None
        if method == "VOXELGRID":
            print(f"downsample: VOXELGRID {res:g}", file=out)
            self.downsample_filter = VoxelGrid(device=device)
            self.downsample_filter.setLeafSize(res, res, res)
        elif method == "APPROX_VOXELGRID":
            raise NotImplementedError("APPROX_VOXELGRID stays on the reference's pcl::ApproximateVoxelGrid (not on the B200 path)")
        else:
            if method != "NONE":
                print(f"warning: unknown downsampling type ({method})", file=sys.stderr)
                print("       : use passthrough filter", file=sys.stderr)
            print("downsample: NONE", file=out)
        if downsample_filter is not None:
            self.downsample_filter = downsample_filter
        self.registration = registration if registration is not None else select_registration_method(p, device=device, out=out)
        self.keyframe = None
        self.keyframe_pose = np.eye(4, dtype=np.float32)
        self.keyframe_stamp = 0.0
        self.prev_trans = np.eye(4, dtype=np.float32)
        self.prev_time = None
        self.num_keyframes = 0
        self.last_converged = True
        # scheduling hint for the engine (not part of the reference's logic, never changes a result): when the
        # motion since the keyframe says this scan will probably become the next keyframe, the registration
        # is asked to build the scan's target structures while it is being aligned (preparePromotion)
        self.prepare_promotion = bool(p.get("prepare_promotion", False))
        self._last_step = 0.0
        self.promotions_prepared = 0
        # pcl::Registration::align(*aligned, guess) [REF apps/scan_matching_odometry_nodelet.cpp:217-218] always fills the
        # aligned cloud.  The mirror asks the engine for it when the caller provides the cloud to fill (`aligned_out`, an
        # (M, 4) float32 array) — the bench's end-to-end legs do; parity tests that only read the transform leave it out.
        self.aligned_out = None
        self.aligned = None

    def downsample(self, cloud):
        if self.downsample_filter is None:
            # pcl::PassThrough without a filter field copies the input into a NEW cloud object.  The
            # engine copies the cloud to its own device buffer on setInputSource, which gives the same
            # isolation from the caller's buffer, so the host-side copy is elided (a new view object
            # keeps the "keyframe is filtered" identity test of matching() working).
            if isinstance(cloud, DeviceCloud):
                return DeviceCloud(cloud.ptr, cloud.n, cloud.owner)
            return np.asarray(cloud, dtype=np.float32).view()
        self.downsample_filter.setInputCloud(cloud, is_dense=True)
        if isinstance(cloud, DeviceCloud):
            if not self.downsample_bufs or len(self.downsample_bufs) < 3:
                raise ValueError("ScanMatchingOdometry: device-resident scans with downsample_method VOXELGRID need downsample_bufs=[three DeviceClouds] (or downsample_method NONE)")
            buf = self.downsample_bufs[self._ds_next % len(self.downsample_bufs)]
            self._ds_next += 1
            return self.downsample_filter.filter(out=buf)
        return self.downsample_filter.filter()

    def matching(self, stamp, cloud, msf_delta=None):
        reg = self.registration
        if self.keyframe is None:
            self.prev_time = None
            self.prev_trans = np.eye(4, dtype=np.float32)
            self.keyframe_pose = np.eye(4, dtype=np.float32)
            self.keyframe_stamp = stamp
            self.keyframe = self.downsample(cloud)
            reg.setInputTarget(self.keyframe)
            self.num_keyframes = 1
            return np.eye(4, dtype=np.float32)

        filtered = self.downsample(cloud)
        reg.setInputSource(filtered)
        guess = self.prev_trans if msf_delta is None else (self.prev_trans @ np.asarray(msf_delta, np.float32))
        dist_before = float(np.linalg.norm(self.prev_trans[:3, 3]))
        if self.prepare_promotion and hasattr(reg, "preparePromotion") and dist_before + self._last_step > 0.95 * self.keyframe_delta_trans:
            reg.preparePromotion()
            self.promotions_prepared += 1
        if self.aligned_out is not None:
            self.aligned = reg.align(guess, aligned_out=self.aligned_out)
        else:
            reg.align(guess)
        self.last_converged = reg.hasConverged()
        if not self.last_converged:
            # "scan matching has not converged!! ignore this frame": state untouched
            return self.keyframe_pose @ self.prev_trans

        trans = reg.getFinalTransformation()
        odom = self.keyframe_pose @ trans
        if self.transform_thresholding:
            delta = np.linalg.inv(self.prev_trans) @ trans
            dx = float(np.linalg.norm(delta[:3, 3]))
            da = float(np.arccos(np.clip(quaternion_w(delta[:3, :3]), -1.0, 1.0)))
            if dx > self.max_acceptable_trans or da > self.max_acceptable_angle:
                return self.keyframe_pose @ self.prev_trans

        self.prev_time = stamp
        self.prev_trans = trans
        delta_trans = float(np.linalg.norm(trans[:3, 3]))
        self._last_step = max(delta_trans - dist_before, 0.0)
        delta_angle = float(np.arccos(np.clip(quaternion_w(trans[:3, :3]), -1.0, 1.0)))
        delta_time = stamp - self.keyframe_stamp
        if delta_trans > self.keyframe_delta_trans or delta_angle > self.keyframe_delta_angle or delta_time > self.keyframe_delta_time:
            self.keyframe = filtered
            # keyframe = filtered; registration->setInputTarget(keyframe): the cloud just aligned as the source.  The
            # engine changes its role on the device; any other registration object gets the plain call.
            if hasattr(reg, "promoteSourceToTarget"):
                reg.promoteSourceToTarget()
            else:
                reg.setInputTarget(self.keyframe)
            self.keyframe_pose = odom
            self.keyframe_stamp = stamp
            self.prev_time = stamp
            self.prev_trans = np.eye(4, dtype=np.float32)
            self.num_keyframes += 1
        return odom


class NativeFrontEnd:
    """b200reg_frontend_* (include/b200reg.h, csrc/b200reg_odometry.cu): the prefiltering nodelet and the scan-matching
    nodelet as host C++ above the engine — the same state machine as Prefilter + ScanMatchingOdometry + FrontEnd in this
    module, without an interpreter between the kernels (the reference's nodelets are compiled C++).  `params` takes the
    two nodelets' parameter names (this module's classes document them)."""

    def __init__(self, params=None, device=0, filter_sms=52, prepare_promotion=0, side_sms=16):
        import ctypes as C
        from . import _lib
        p = dict(params or {})
        L = _lib.load()
        cfg = _lib.FrontEndConfig()
        L.b200reg_frontend_default_config(C.byref(cfg))
        cfg.device = device
        method = p.get("registration_method", "NDT_OMP")
        if method == "FAST_GICP":
            L.b200reg_default_config(_lib.METHOD_GICP, C.byref(cfg.registration))
            cfg.registration.max_correspondence_distance = float(p.get("reg_max_correspondence_distance", 2.5))
            cfg.registration.correspondence_randomness = int(p.get("reg_correspondence_randomness", 20))
        elif "NDT" in method and "OMP" in method:
            L.b200reg_default_config(_lib.METHOD_NDT, C.byref(cfg.registration))
            cfg.registration.resolution = float(p.get("reg_resolution", 0.5))
            cfg.registration.nn_search = {"KDTREE": _lib.KDTREE, "DIRECT1": _lib.DIRECT1}.get(p.get("reg_nn_search_method", "DIRECT7"), _lib.DIRECT7)
        else:
            raise NotImplementedError(f"registration_method={method}: the engine replaces NDT_OMP and FAST_GICP")
        cfg.registration.transformation_epsilon = float(p.get("reg_transformation_epsilon", 0.01))
        cfg.registration.maximum_iterations = int(p.get("reg_maximum_iterations", 64))
        cfg.odometry.keyframe_delta_trans = float(p.get("keyframe_delta_trans", 0.25))
        cfg.odometry.keyframe_delta_angle = float(p.get("keyframe_delta_angle", 0.15))
        cfg.odometry.keyframe_delta_time = float(p.get("keyframe_delta_time", 1.0))
        cfg.odometry.transform_thresholding = int(bool(p.get("transform_thresholding", False)))
        cfg.odometry.max_acceptable_trans = float(p.get("max_acceptable_trans", 1.0))
        cfg.odometry.max_acceptable_angle = float(p.get("max_acceptable_angle", 1.0))
        if p.get("downsample_method", "VOXELGRID") != "VOXELGRID":
            raise NotImplementedError("the native front end runs the prefilter's VoxelGrid (downsample_method VOXELGRID)")
        cfg.downsample_resolution = float(p.get("downsample_resolution", 0.1))
        cfg.use_distance_filter = 0 if p.get("b200_skip_distance_filter", False) else 1
        cfg.distance_near_thresh = float(p.get("distance_near_thresh", 1.0))
        cfg.distance_far_thresh = float(p.get("distance_far_thresh", 100.0))
        cfg.filter_sms = int(filter_sms)
        cfg.prepare_promotion = int(prepare_promotion)
        cfg.side_sms = int(side_sms)
        self._cfg = cfg
        self._h = C.c_void_p()
        rc = L.b200reg_frontend_create(C.byref(cfg), C.byref(self._h))
        if rc != _lib.OK:
            self._h = None
            raise _lib.B200RegError(rc, "b200reg_frontend_create failed (no usable sm_100 CUDA device?)")
        self._L, self._C = L, C

    def close(self):
        if getattr(self, "_h", None):
            self._L.b200reg_frontend_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            from . import _lib
            raise _lib.B200RegError(rc, self._L.b200reg_frontend_last_error(self._h).decode())

    def registration_handle(self):
        return self._L.b200reg_frontend_registration(self._h)

    def stream(self):
        p = self._C.c_void_p()
        self._L.b200reg_get_stream(self.registration_handle(), self._C.byref(p))
        return p.value

    def set_timing(self, on=True):
        self._L.b200reg_set_timing(self.registration_handle(), int(on))

    def counters(self):
        C = self._C
        a, b, c = C.c_longlong(), C.c_longlong(), C.c_double()
        self._L.b200reg_get_counters(self.registration_handle(), C.byref(a), C.byref(b), C.byref(c))
        return dict(launches_total=a.value, timed_aligns=b.value, align_kernel_ms=c.value)

    def num_keyframes(self):
        n = self._C.c_int()
        self._L.b200reg_odometry_get_state(self._L.b200reg_frontend_odometry(self._h), self._C.byref(n), None, None, None, None, None)
        return n.value

    def timing(self):
        """Host wall clock per phase since the last call (b200reg_frontend_get_timing), microseconds per step."""
        v = np.zeros(9)
        self._ck(self._L.b200reg_frontend_get_timing(self._h, v.ctypes.data))
        n = max(v[0], 1.0)
        return dict(steps=int(v[0]), filter_wait_us=v[1] / n, begin_next_us=v[2] / n, set_source_us=v[3] / n, align_us=v[4] / n, promote_us_per_step=v[5] / n,
                    promote_us_each=v[5] / max(v[6], 1.0), promotions=int(v[6]), step_us=v[7] / n, hints=int(v[8]))

    def run_device(self, clouds, want_results=True):
        """A whole sequence of DeviceClouds: (poses (F, 4, 4) float32, result records or None, filtered sizes)."""
        from . import _lib
        C = self._C
        F = len(clouds)
        ptrs = (C.c_void_p * F)(*[c.ptr for c in clouds])
        ns = (C.c_size_t * F)(*[c.n for c in clouds])
        odom = np.zeros((F, 16), np.float32)
        res = np.zeros(F, _lib.RESULT_DTYPE) if want_results else None
        nf = np.zeros(F, np.uint64)
        kf = C.c_int()
        self._ck(self._L.b200reg_frontend_run_device(self._h, ptrs, ns, None, F, odom.ctypes.data, res.ctypes.data if want_results else None, nf.ctypes.data, C.byref(kf)))
        return odom.reshape(F, 4, 4).transpose(0, 2, 1).copy(), res, nf

    def reset(self):
        """Back to the state of a freshly created front end (b200reg_frontend_reset): no keyframe, nothing in flight."""
        self._ck(self._L.b200reg_frontend_reset(self._h))

    def run_host(self, clouds, filtered_bufs=None, aligned_out=None, stamps=None):
        """Host scans one by one through begin / step.  filtered_bufs: three (M, 4) float32 host clouds in rotation
        (the two-nodelet form: filtered cloud to the host, uploaded again for the registration) or None (fused)."""
        C = self._C
        F = len(clouds)
        poses = np.zeros((F, 16), np.float32)
        nf = C.c_size_t()
        stamp = (lambda k: 0.1 * k) if stamps is None else (lambda k: stamps[k])

        def fb(k):
            if filtered_bufs is None:
                return None, 0
            b = filtered_bufs[k % len(filtered_bufs)]
            return b.ctypes.data, len(b)
        if F == 0:
            return poses.reshape(0, 4, 4)
        self.reset()  # a new sequence: first scan -> first keyframe (run_device resets inside the library)
        c0 = clouds[0]
        p0, cap0 = fb(0)
        self._ck(self._L.b200reg_frontend_begin(self._h, stamp(0), c0.ctypes.data, len(c0), 16, p0, cap0))
        al = aligned_out.ctypes.data if aligned_out is not None else None
        for k in range(F):
            if k + 1 < F:
                c1 = clouds[k + 1]
                p1, cap1 = fb(k + 1)
                self._ck(self._L.b200reg_frontend_step(self._h, stamp(k + 1), c1.ctypes.data, len(c1), 16, p1, cap1, C.byref(nf), al, poses[k].ctypes.data))
            else:
                self._ck(self._L.b200reg_frontend_step(self._h, 0.0, None, 0, 16, None, 0, C.byref(nf), al, poses[k].ctypes.data))
        return poses.reshape(F, 4, 4).transpose(0, 2, 1).copy()
