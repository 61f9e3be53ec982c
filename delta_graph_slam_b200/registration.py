"""Host-side mirror of the reference's registration interface over the b200reg C ABI.

`select_registration_method(params)` follows [REF src/hdl_graph_slam/registrations.cpp:22-124]:
same parameter names (`registration_method`, `reg_*`), same defaults, same banners, same
warning for unknown strings (which, like plain "NDT", select the reference's single-thread pcl NDT: not replaced, raises).  The objects it returns expose the
pcl::Registration calls the reference makes: setInputTarget, setInputSource,
align, hasConverged, getFinalTransformation, getFitnessScore — backed by the CUDA engine only.
"""
import ctypes as C
import sys

import numpy as np

from . import _lib
from ._lib import B200RegError, Config, Result

DBL_MAX = float(np.finfo(np.float64).max)


class DeviceCloud:
    """A cloud that already lives on the GPU: raw device pointer to N float4 records.  `owner` keeps
    whatever allocated the memory (e.g. a torch tensor) alive.  Accepted wherever a host cloud is."""

    def __init__(self, ptr, n, owner=None):
        self.ptr = int(ptr)
        self.n = int(n)
        self.owner = owner

    def __len__(self):
        return self.n


class Registration:
    """pcl::Registration<PointXYZ, PointXYZ> call surface on one b200reg handle."""

    method = _lib.METHOD_NONE
    reg_name = "Registration"

    def __init__(self, device=0, **overrides):
        L = _lib.load()
        cfg = Config()
        L.b200reg_default_config(self.method, C.byref(cfg))
        cfg.device = device
        for k, v in overrides.items():
            setattr(cfg, k, v)
        self._cfg = cfg
        self._h = C.c_void_p()
        rc = L.b200reg_create(C.byref(cfg), C.byref(self._h))
        if rc != _lib.OK:
            self._h = None
            raise B200RegError(rc, "b200reg_create failed (no usable sm_100 CUDA device?)")
        self._n_src = 0
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            _lib.load().b200reg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        _lib.check(self._h, rc)

    # ---- setters used by the factory
    def setNumThreads(self, n):  # reg_num_threads: the CUDA grid replaces OpenMP
        self._cfg.num_threads = int(n)

    def setTransformationEpsilon(self, eps):
        self._ck(_lib.load().b200reg_set_transformation_epsilon(self._h, float(eps)))

    def setMaximumIterations(self, n):
        self._ck(_lib.load().b200reg_set_maximum_iterations(self._h, int(n)))

    # ---- data
    def setInputTarget(self, cloud):
        """pcl::Registration::setInputTarget: the cloud is copied to the device at the call (host arrays may be
        reused afterwards).  A caller that wants the source it has just aligned to become the target without
        another upload says so explicitly with promoteSourceToTarget()."""
        if isinstance(cloud, DeviceCloud):
            return self.setInputTargetDevice(cloud.ptr, cloud.n)
        c = _lib.as_cloud(cloud)
        if len(c) == 0:  # PCL_ERROR + return: the previous target stays
            print("[b200reg::setInputTarget] Invalid or empty point cloud dataset given!", file=sys.stderr)
            return
        self._ck(_lib.load().b200reg_set_target(self._h, c.ctypes.data, len(c), 16))

    def setInputSource(self, cloud):
        if isinstance(cloud, DeviceCloud):
            self.setInputSourceDevice(cloud.ptr, cloud.n)
            return
        c = _lib.as_cloud(cloud)
        self._ck(_lib.load().b200reg_set_source(self._h, c.ctypes.data if len(c) else None, len(c), 16))
        self._n_src = len(c)

    def setInputTargetDevice(self, ptr, n):
        self._ck(_lib.load().b200reg_set_target_device(self._h, ptr, n))

    def setInputSourceDevice(self, ptr, n):
        self._ck(_lib.load().b200reg_set_source_device(self._h, ptr, n))
        self._n_src = n

    def promoteSourceToTarget(self):
        """keyframe = filtered; registration->setInputTarget(keyframe) [REF apps/scan_matching_odometry_nodelet.cpp:253-254]
        for the cloud that is the current source: it changes role on the device (b200reg_promote_source_to_target)."""
        self._ck(_lib.load().b200reg_promote_source_to_target(self._h))

    def preparePromotion(self):
        """Scheduling hint before align: the source will probably become the target (b200reg_prepare_promotion)."""
        self._ck(_lib.load().b200reg_prepare_promotion(self._h))

    def setSideBudget(self, n_sm):
        self._ck(_lib.load().b200reg_set_side_budget(self._h, int(n_sm)))

    # ---- run
    def align(self, guess=None, want_aligned=False, aligned_out=None):
        """registration->align(*aligned, guess).  Returns the aligned cloud when asked for: a fresh array
        (want_aligned) or the first n_source rows of the caller's `aligned_out` ((M, 4) float32, M >= n_source;
        page-locked memory is written by DMA without staging)."""
        g = _lib.colmajor(np.eye(4) if guess is None else guess)
        out = None
        if aligned_out is not None:
            if aligned_out.dtype != np.float32 or aligned_out.ndim != 2 or aligned_out.shape[1] != 4 or not aligned_out.flags.c_contiguous or len(aligned_out) < self._n_src:
                raise ValueError("aligned_out must be a C-contiguous (M, 4) float32 array with M >= the source size")
            out = aligned_out[: self._n_src]
        elif want_aligned:
            out = np.zeros((self._n_src, 4), np.float32)
        rc = _lib.load().b200reg_align(self._h, g.ctypes.data, out.ctypes.data if out is not None and self._n_src else None)
        if rc == _lib.E_STATE:  # PCL logs and returns with converged_ == false
            print(f"[b200reg::align] {_lib.load().b200reg_last_error(self._h).decode()}", file=sys.stderr)
            return out
        self._ck(rc)
        return out

    # ---- results
    def hasConverged(self):
        v = C.c_int()
        self._ck(_lib.load().b200reg_has_converged(self._h, C.byref(v)))
        return bool(v.value)

    def getFinalTransformation(self):
        T = np.zeros(16, np.float32)
        self._ck(_lib.load().b200reg_get_final_transformation(self._h, T.ctypes.data))
        return _lib.from_colmajor(T)

    def getFinalNumIteration(self):
        v = C.c_int()
        self._ck(_lib.load().b200reg_get_num_iterations(self._h, C.byref(v)))
        return v.value

    def getFitnessScore(self, max_range=DBL_MAX):
        v = C.c_double()
        self._ck(_lib.load().b200reg_get_fitness_score(self._h, float(max_range), C.byref(v)))
        return v.value

    def calcFitnessScore(self, relpose, max_range=DBL_MAX):
        """InformationMatrixCalculator::calc_fitness_score(target, source, relpose, max_range)."""
        T = _lib.colmajor(relpose)
        v = C.c_double()
        self._ck(_lib.load().b200reg_calc_fitness_score(self._h, T.ctypes.data, float(max_range), C.byref(v)))
        return v.value

    def getInlierFraction(self, max_correspondence_dist=0.5):
        v = C.c_double()
        self._ck(_lib.load().b200reg_get_inlier_fraction(self._h, float(max_correspondence_dist), C.byref(v)))
        return v.value

    def getResult(self):
        r = Result()
        self._ck(_lib.load().b200reg_get_result(self._h, C.byref(r)))
        return dict(transformation=_lib.from_colmajor(np.array(r.transformation[:], np.float32)), fitness=r.fitness, score=r.score, converged=bool(r.converged),
                    iterations=r.iterations, evaluations=r.evaluations, passes=r.passes, hits=r.hits)

    # ---- loop-closure batches (LoopDetector::matching, many candidates in one call)
    def cloudPut(self, cloud_id, cloud):
        """Cache a keyframe cloud on the device under `cloud_id` (host array or DeviceCloud)."""
        if isinstance(cloud, DeviceCloud):
            self._ck(_lib.load().b200reg_cloud_put_device(self._h, int(cloud_id), cloud.ptr, cloud.n))
            return
        c = _lib.as_cloud(cloud)
        self._ck(_lib.load().b200reg_cloud_put(self._h, int(cloud_id), c.ctypes.data if len(c) else None, len(c), 16))

    def cloudSync(self):
        """Wait until every cloud put so far has left the caller's (page-locked) memory."""
        self._ck(_lib.load().b200reg_cloud_sync(self._h))

    def setInputSourceCached(self, cloud_id):
        """setInputSource with a cloud of the keyframe cache (b200reg_set_source_cached): no upload, and on a FAST_GICP handle
        the keyframe's covariances are computed once in its life instead of once per pair."""
        self._ck(_lib.load().b200reg_set_source_cached(self._h, int(cloud_id)))
        self._src_ref = None

    def setInputTargetCached(self, cloud_id):
        """setInputTarget with a cloud of the keyframe cache (b200reg_set_target_cached)."""
        self._ck(_lib.load().b200reg_set_target_cached(self._h, int(cloud_id)))

    def cloudDrop(self, cloud_id):
        self._ck(_lib.load().b200reg_cloud_drop(self._h, int(cloud_id)))

    def cloudClear(self):
        self._ck(_lib.load().b200reg_cloud_clear(self._h))

    def cloudCount(self):
        n = C.c_size_t()
        self._ck(_lib.load().b200reg_cloud_count(self._h, C.byref(n)))
        return n.value

    def alignBatch(self, pairs, with_fitness=True, fitness_max_range=DBL_MAX):
        """pairs: iterable of (target_id, source_id, guess 4x4) or a PAIR_DTYPE array.  Returns a
        RESULT_DTYPE structured array, one record per pair (transformation column-major)."""
        if isinstance(pairs, np.ndarray) and pairs.dtype == _lib.PAIR_DTYPE:
            arr = np.ascontiguousarray(pairs)
        else:
            pairs = list(pairs)
            arr = np.zeros(len(pairs), _lib.PAIR_DTYPE)
            for i, (t, s, g) in enumerate(pairs):
                arr[i] = (int(t), int(s), _lib.colmajor(np.eye(4) if g is None else g))
        out = np.zeros(len(arr), _lib.RESULT_DTYPE)
        if len(arr):
            self._ck(_lib.load().b200reg_align_batch(self._h, arr.ctypes.data, len(arr), int(with_fitness), float(fitness_max_range), out.ctypes.data))
        return out

    def calcFitnessBatch(self, pairs, max_range=DBL_MAX):
        """calc_fitness_score(cloud1 = target_id, cloud2 = source_id, relpose) for every pair of cached
        keyframes; pairs as in alignBatch with the relative pose in the `guess` slot.  Returns float64[n]."""
        if isinstance(pairs, np.ndarray) and pairs.dtype == _lib.PAIR_DTYPE:
            arr = np.ascontiguousarray(pairs)
        else:
            pairs = list(pairs)
            arr = np.zeros(len(pairs), _lib.PAIR_DTYPE)
            for i, (t, s, g) in enumerate(pairs):
                arr[i] = (int(t), int(s), _lib.colmajor(np.eye(4) if g is None else g))
        out = np.zeros(len(arr), np.float64)
        if len(arr):
            self._ck(_lib.load().b200reg_calc_fitness_batch(self._h, arr.ctypes.data, len(arr), float(max_range), out.ctypes.data))
        return out

    def batchTiming(self):
        a, b = C.c_double(), C.c_double()
        self._ck(_lib.load().b200reg_get_batch_timing(self._h, C.byref(a), C.byref(b)))
        return dict(align_kernel_ms=a.value, fitness_ms=b.value)

    def setTiming(self, on=True):
        self._ck(_lib.load().b200reg_set_timing(self._h, int(on)))

    def counters(self):
        a, b, c = C.c_longlong(), C.c_longlong(), C.c_double()
        self._ck(_lib.load().b200reg_get_counters(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(launches_total=a.value, timed_aligns=b.value, align_kernel_ms=c.value)

    def setProfile(self, on=True):
        """Developer cycle counters of the align kernels (a separate kernel instantiation; off by default)."""
        self._ck(_lib.load().b200reg_set_profile(self._h, int(on)))

    def profile(self):
        v = np.zeros(16, np.int64)
        self._ck(_lib.load().b200reg_get_profile(self._h, v.ctypes.data))
        return dict(zip(("pass", "reduce", "barrier", "total", "step", "n", "stage", "step_solve_mt", "step_trig", "step_tables",
                         "st_totals", "st_interval", "st_trial", "st_newton_end", "st_solve", "st_newton_begin"), v.tolist()))

    def trace(self, cap=255):
        """Per-pass More-Thuente quantities of the last profiled NDT align (b200reg_get_trace), (n, 13) float64."""
        out = np.zeros((cap, 13))
        n = C.c_size_t()
        self._ck(_lib.load().b200reg_get_trace(self._h, out.ctypes.data, cap, C.byref(n)))
        return out[: min(n.value, cap)]

    def nn_stats(self):
        v = np.zeros(3, np.int64)
        self._ck(_lib.load().b200reg_get_nn_stats(self._h, v.ctypes.data))
        return dict(queries=int(v[0]), far_pass=int(v[1]), brute_pass=int(v[2]))

    def stream(self):
        p = C.c_void_p()
        self._ck(_lib.load().b200reg_get_stream(self._h, C.byref(p)))
        return p.value

    # ---- pcl::VoxelGrid on the same device / stream
    def voxelgrid_filter(self, cloud, leaf, min_points_per_voxel=0, is_dense=False, out=None):
        """pcl::Filter::filter(output).  With `out` (a caller-owned (M, 4) float32 array, M >= len(cloud);
        page-locked memory is written by DMA without staging) the result is the view out[:n]."""
        c = _lib.as_cloud(cloud)
        leaf3 = (C.c_float * 3)(*((leaf,) * 3 if np.isscalar(leaf) else leaf))
        own = out is None
        if own:
            out = np.empty((max(len(c), 1), 4), np.float32)
        elif out.dtype != np.float32 or out.ndim != 2 or out.shape[1] != 4 or not out.flags.c_contiguous or len(out) < len(c):
            raise ValueError("out must be a C-contiguous (M, 4) float32 array with M >= len(cloud)")
        n_out = C.c_size_t()
        self._ck(_lib.load().b200reg_voxelgrid_filter(self._h, c.ctypes.data if len(c) else None, len(c), 16, leaf3, min_points_per_voxel, int(is_dense), out.ctypes.data,
                                                      len(out), C.byref(n_out)))
        return out[: n_out.value].copy() if own else out[: n_out.value]

    def voxelgrid_filter_device(self, cloud, leaf, out, min_points_per_voxel=0, is_dense=False):
        """Device-resident filter: `cloud` and `out` are DeviceClouds (out.n = capacity >= cloud.n)."""
        leaf3 = (C.c_float * 3)(*((leaf,) * 3 if np.isscalar(leaf) else leaf))
        n_out = C.c_size_t()
        if out.n < cloud.n:
            raise ValueError("output buffer smaller than the input cloud")
        self._ck(_lib.load().b200reg_voxelgrid_filter_device(self._h, cloud.ptr, cloud.n, leaf3, min_points_per_voxel, int(is_dense), out.ptr, C.byref(n_out)))
        return DeviceCloud(out.ptr, n_out.value, out.owner)

    def voxelgrid_filter_begin(self, cloud, leaf, out, min_points_per_voxel=0, is_dense=False):
        """First half of the filter: enqueue and return (b200reg_voxelgrid_filter_begin / _device_begin).
        `cloud` / `out` are both host arrays or both DeviceClouds; they stay untouched until the end call."""
        leaf3 = (C.c_float * 3)(*((leaf,) * 3 if np.isscalar(leaf) else leaf))
        if isinstance(cloud, DeviceCloud):
            if out.n < cloud.n:
                raise ValueError("output buffer smaller than the input cloud")
            self._ck(_lib.load().b200reg_voxelgrid_filter_device_begin(self._h, cloud.ptr, cloud.n, leaf3, min_points_per_voxel, int(is_dense), out.ptr))
            self._vg_pending = (cloud, out)
            return
        c = _lib.as_cloud(cloud)
        if out.dtype != np.float32 or out.ndim != 2 or out.shape[1] != 4 or not out.flags.c_contiguous or len(out) < len(c):
            raise ValueError("out must be a C-contiguous (M, 4) float32 array with M >= len(cloud)")
        self._ck(_lib.load().b200reg_voxelgrid_filter_begin(self._h, c.ctypes.data if len(c) else None, len(c), 16, leaf3, min_points_per_voxel, int(is_dense), out.ctypes.data, len(out)))
        self._vg_pending = (c, out)

    def voxelgrid_filter_end(self):
        """Second half: wait for the filter in flight; returns the filtered cloud (a view of `out`)."""
        _, out = self._vg_pending
        self._vg_pending = None
        n_out = C.c_size_t()
        self._ck(_lib.load().b200reg_voxelgrid_filter_end(self._h, C.byref(n_out)))
        if isinstance(out, DeviceCloud):
            return DeviceCloud(out.ptr, n_out.value, out.owner)
        return out[: n_out.value]

    def setDistanceFilter(self, use, near_thresh=1.0, far_thresh=100.0):
        """distance_filter of the prefiltering nodelet fused into this handle's VoxelGrid calls (b200reg_set_distance_filter)."""
        self._ck(_lib.load().b200reg_set_distance_filter(self._h, int(bool(use)), float(near_thresh), float(far_thresh)))

    def setInputTransform(self, matrix4x4):
        """pcl::transformPointCloud(src, transformed, Matrix4d) in front of this handle's VoxelGrid / distance_filter calls
        (b200reg_set_input_transform) [REF apps/prefiltering_nodelet.cpp:137-147]; None switches it off."""
        if matrix4x4 is None:
            self._ck(_lib.load().b200reg_set_input_transform(self._h, None))
            return
        m = np.ascontiguousarray(np.asarray(matrix4x4, np.float64).reshape(4, 4).T)  # column-major, as Eigen stores it
        self._ck(_lib.load().b200reg_set_input_transform(self._h, m.ctypes.data_as(C.POINTER(C.c_double))))

    def distance_filter(self, cloud, near_thresh=1.0, far_thresh=100.0, out=None):
        """PrefilteringNodelet::distance_filter as a call of its own (b200reg_distance_filter): for a prefilter without a VoxelGrid."""
        n_out = C.c_size_t()
        if isinstance(cloud, DeviceCloud):
            if out is None or out.n < cloud.n:
                raise ValueError("a device cloud needs a device output buffer at least as large as the input")
            self._ck(_lib.load().b200reg_distance_filter_device(self._h, cloud.ptr, cloud.n, float(near_thresh), float(far_thresh), out.ptr, C.byref(n_out)))
            return DeviceCloud(out.ptr, n_out.value, out.owner)
        c = _lib.as_cloud(cloud)
        own = out is None
        if own:
            out = np.empty((max(len(c), 1), 4), np.float32)
        elif out.dtype != np.float32 or out.ndim != 2 or out.shape[1] != 4 or not out.flags.c_contiguous or len(out) < len(c):
            raise ValueError("out must be a C-contiguous (M, 4) float32 array with M >= len(cloud)")
        self._ck(_lib.load().b200reg_distance_filter(self._h, c.ctypes.data if len(c) else None, len(c), 16, float(near_thresh), float(far_thresh), out.ctypes.data, len(out), C.byref(n_out)))
        return out[: n_out.value].copy() if own else out[: n_out.value]

    def radius_outlier_removal_begin(self, cloud, radius, min_neighbors, out):
        """pcl::RadiusOutlierRemoval, first half (enqueue).  `cloud` / `out` both host arrays or both DeviceClouds."""
        if isinstance(cloud, DeviceCloud):
            if out.n < cloud.n:
                raise ValueError("output buffer smaller than the input cloud")
            self._ck(_lib.load().b200reg_radius_outlier_removal_device_begin(self._h, cloud.ptr, cloud.n, float(radius), int(min_neighbors), out.ptr))
            self._ror_pending = (cloud, out)
            return
        c = _lib.as_cloud(cloud)
        if out.dtype != np.float32 or out.ndim != 2 or out.shape[1] != 4 or not out.flags.c_contiguous or len(out) < len(c):
            raise ValueError("out must be a C-contiguous (M, 4) float32 array with M >= len(cloud)")
        self._ck(_lib.load().b200reg_radius_outlier_removal_begin(self._h, c.ctypes.data if len(c) else None, len(c), 16, float(radius), int(min_neighbors), out.ctypes.data, len(out)))
        self._ror_pending = (c, out)

    def radius_outlier_removal_end(self):
        _, out = self._ror_pending
        self._ror_pending = None
        n_out = C.c_size_t()
        self._ck(_lib.load().b200reg_radius_outlier_removal_end(self._h, C.byref(n_out)))
        if isinstance(out, DeviceCloud):
            return DeviceCloud(out.ptr, n_out.value, out.owner)
        return out[: n_out.value]

    def radius_outlier_removal(self, cloud, radius, min_neighbors, out=None):
        own = out is None
        if own:
            if isinstance(cloud, DeviceCloud):
                raise ValueError("a device cloud needs a device output buffer")
            out = np.empty((max(len(cloud), 1), 4), np.float32)
        self.radius_outlier_removal_begin(cloud, radius, min_neighbors, out)
        res = self.radius_outlier_removal_end()
        return res.copy() if own else res

    def statistical_outlier_removal_begin(self, cloud, mean_k, stddev_mul, out):
        """pcl::StatisticalOutlierRemoval, first half (enqueue).  `cloud` / `out` both host arrays or both DeviceClouds."""
        if isinstance(cloud, DeviceCloud):
            if out.n < cloud.n:
                raise ValueError("output buffer smaller than the input cloud")
            self._ck(_lib.load().b200reg_statistical_outlier_removal_device_begin(self._h, cloud.ptr, cloud.n, int(mean_k), float(stddev_mul), out.ptr))
            self._ror_pending = (cloud, out)
            return
        c = _lib.as_cloud(cloud)
        if out.dtype != np.float32 or out.ndim != 2 or out.shape[1] != 4 or not out.flags.c_contiguous or len(out) < len(c):
            raise ValueError("out must be a C-contiguous (M, 4) float32 array with M >= len(cloud)")
        self._ck(_lib.load().b200reg_statistical_outlier_removal_begin(self._h, c.ctypes.data if len(c) else None, len(c), 16, int(mean_k), float(stddev_mul), out.ctypes.data, len(out)))
        self._ror_pending = (c, out)

    def statistical_outlier_removal_end(self):
        return self.radius_outlier_removal_end()  # one in-flight slot and one _end for both outlier filters

    def statistical_outlier_removal(self, cloud, mean_k, stddev_mul, out=None):
        own = out is None
        if own:
            if isinstance(cloud, DeviceCloud):
                raise ValueError("a device cloud needs a device output buffer")
            out = np.empty((max(len(cloud), 1), 4), np.float32)
        self.statistical_outlier_removal_begin(cloud, mean_k, stddev_mul, out)
        res = self.statistical_outlier_removal_end()
        return res.copy() if own else res

    def flat_filter_begin(self, cloud, lidar_z, out, normal_k=10, normal_thresh=0.2):
        """height_filtering -> normal_filtering -> flatten of the prefilter nodelet in one call (b200reg_flat_filter), first half."""
        if isinstance(cloud, DeviceCloud):
            if out.n < cloud.n:
                raise ValueError("output buffer smaller than the input cloud")
            self._ck(_lib.load().b200reg_flat_filter_device_begin(self._h, cloud.ptr, cloud.n, float(lidar_z), int(normal_k), float(normal_thresh), out.ptr))
            self._ror_pending = (cloud, out)
            return
        c = _lib.as_cloud(cloud)
        if out.dtype != np.float32 or out.ndim != 2 or out.shape[1] != 4 or not out.flags.c_contiguous or len(out) < len(c):
            raise ValueError("out must be a C-contiguous (M, 4) float32 array with M >= len(cloud)")
        self._ck(_lib.load().b200reg_flat_filter_begin(self._h, c.ctypes.data if len(c) else None, len(c), 16, float(lidar_z), int(normal_k), float(normal_thresh), out.ctypes.data, len(out)))
        self._ror_pending = (c, out)

    def flat_filter_end(self):
        return self.radius_outlier_removal_end()

    def flat_filter(self, cloud, lidar_z, out=None, normal_k=10, normal_thresh=0.2):
        own = out is None
        if own:
            if isinstance(cloud, DeviceCloud):
                raise ValueError("a device cloud needs a device output buffer")
            out = np.empty((max(len(cloud), 1), 4), np.float32)
        self.flat_filter_begin(cloud, lidar_z, out, normal_k, normal_thresh)
        res = self.flat_filter_end()
        return res.copy() if own else res

    def flat_filter_last_nz(self, n_points):
        nz = np.zeros(max(n_points, 1), np.float32)
        self._ck(_lib.load().b200reg_flat_filter_last_nz(self._h, nz.ctypes.data, n_points))
        return nz[:n_points]

    def statistical_last_stats(self, n_points=0):
        """{mean, stddev, threshold, valid, exact_pass[, distances]} of the last statistical call (b200reg_statistical_last_stats)."""
        st = np.zeros(3, np.float64)
        valid = C.c_ulonglong()
        exact = C.c_int()
        dist = np.zeros(max(n_points, 1), np.float32)
        self._ck(_lib.load().b200reg_statistical_last_stats(self._h, st.ctypes.data, C.addressof(valid), C.addressof(exact), dist.ctypes.data if n_points else None, n_points))
        return dict(mean=st[0], stddev=st[1], threshold=st[2], valid=valid.value, exact_pass=bool(exact.value), distances=dist[:n_points])

    def setSmBudget(self, n_sm):
        """At most n_sm CTAs (one per SM) for this handle's persistent kernels (b200reg_set_sm_budget)."""
        self._ck(_lib.load().b200reg_set_sm_budget(self._h, int(n_sm)))

    def voxelgrid_last_layout(self, n_voxels, n_points):
        vid = np.zeros(max(n_voxels, 1), np.uint32)
        cnt = np.zeros(max(n_voxels, 1), np.uint32)
        key = np.zeros(max(n_points, 1), np.uint32)
        grid = np.zeros(6, np.int32)
        ovf = C.c_int()
        self._ck(_lib.load().b200reg_voxelgrid_last_layout(self._h, vid.ctypes.data, cnt.ctypes.data, n_voxels, key.ctypes.data, n_points, grid.ctypes.data, C.byref(ovf)))
        return dict(voxel_id=vid[:n_voxels], count=cnt[:n_voxels], key=key[:n_points], min_b=grid[:3], div_b=grid[3:], overflow=bool(ovf.value))


class NormalDistributionsTransform(Registration):
    """Replaces pclomp::NormalDistributionsTransform ("NDT_OMP") [REF registrations.cpp:105-119]."""

    method = _lib.METHOD_NDT
    reg_name = "NormalDistributionsTransform"

    def setResolution(self, r):
        self._ck(_lib.load().b200reg_set_resolution(self._h, float(r)))

    def setNeighborhoodSearchMethod(self, m):
        self._ck(_lib.load().b200reg_set_nn_search(self._h, int(m)))

    def getTransformationProbability(self):
        v = C.c_double()
        self._ck(_lib.load().b200reg_get_transformation_probability(self._h, C.byref(v)))
        return v.value

    def ndt_leaves(self):
        n = C.c_size_t()
        self._ck(_lib.load().b200reg_ndt_num_leaves(self._h, C.byref(n)))
        n = n.value
        m = max(n, 1)
        idx = np.zeros(m, np.uint64); npts = np.zeros(m, np.int32); mean = np.zeros((m, 3)); cov = np.zeros((m, 9)); icov = np.zeros((m, 9))
        cen = np.zeros((m, 3), np.float32); grid = np.zeros(6, np.int32)
        self._ck(_lib.load().b200reg_ndt_get_leaves(self._h, idx.ctypes.data, npts.ctypes.data, mean.ctypes.data, cov.ctypes.data, icov.ctypes.data, cen.ctypes.data, grid.ctypes.data))
        return dict(idx=idx[:n], n=npts[:n], mean=mean[:n], cov=cov[:n].reshape(-1, 3, 3), icov=icov[:n].reshape(-1, 3, 3), centroid=cen[:n], min_b=grid[:3], div_b=grid[3:])

    def ndt_derivatives(self, p):
        p = np.ascontiguousarray(p, np.float64)
        s = C.c_double(); g = np.zeros(6); H = np.zeros(36)
        self._ck(_lib.load().b200reg_ndt_derivatives(self._h, p.ctypes.data, C.byref(s), g.ctypes.data, H.ctypes.data))
        return s.value, g, H.reshape(6, 6)


class FastGICP(Registration):
    """Replaces fast_gicp::FastGICP ("FAST_GICP") [REF registrations.cpp:27-36]."""

    method = _lib.METHOD_GICP
    reg_name = "FastGICP"

    def setMaxCorrespondenceDistance(self, d):
        self._ck(_lib.load().b200reg_set_max_correspondence_distance(self._h, float(d)))

    def setCorrespondenceRandomness(self, k):
        self._ck(_lib.load().b200reg_set_correspondence_randomness(self._h, int(k)))

    def setOptions(self, regularization=_lib.REG_PLANE, lsq_optimizer=_lib.LSQ_LM, rotation_epsilon=2e-3):
        """setRegularizationMethod / setLSQType / setRotationEpsilon of fast_gicp (never called by the reference)."""
        self._ck(_lib.load().b200reg_set_gicp_options(self._h, int(regularization), int(lsq_optimizer), float(rotation_epsilon)))

    def covariances(self, which, n):
        """3x3 covariances of the source (which = 0) or target (1) cloud, (n, 3, 3) float64."""
        out = np.zeros((max(n, 1), 9))
        self._ck(_lib.load().b200reg_gicp_get_covariances(self._h, int(which), out.ctypes.data, n))
        return out[:n].reshape(-1, 3, 3)


class VoxelGrid:
    """pcl::VoxelGrid<PointXYZ> as the nodelets use it: setLeafSize + setInputCloud + filter
    [REF apps/prefiltering_nodelet.cpp:59-63,249-260; apps/scan_matching_odometry_nodelet.cpp:85-89,155-165]."""

    def __init__(self, device=0):
        self._reg = Registration(device=device)
        self._leaf = (0.0, 0.0, 0.0)
        self._input = None
        self.is_dense = False  # distance_filter marks its output is_dense = false [REF apps/prefiltering_nodelet.cpp:286]
        self.min_points_per_voxel = 0

    def setLeafSize(self, lx, ly, lz):
        self._leaf = (float(np.float32(lx)), float(np.float32(ly)), float(np.float32(lz)))

    def setInputCloud(self, cloud, is_dense=False):
        self._input = cloud
        self.is_dense = is_dense

    def setMinimumPointsNumberPerVoxel(self, n):
        self.min_points_per_voxel = int(n)

    def filter(self, out=None):
        if isinstance(self._input, DeviceCloud):
            if not isinstance(out, DeviceCloud):
                raise ValueError("VoxelGrid.filter: a device-resident input needs a caller-owned DeviceCloud output (out=...) with room for len(input) points")
            return self._reg.voxelgrid_filter_device(self._input, self._leaf, out, self.min_points_per_voxel, self.is_dense)
        return self._reg.voxelgrid_filter(self._input, self._leaf, self.min_points_per_voxel, self.is_dense, out=out)

    def setDistanceFilter(self, use, near_thresh=1.0, far_thresh=100.0):
        self._reg.setDistanceFilter(use, near_thresh, far_thresh)

    def filter_begin(self, out):
        """filter() split in two for a pipelined front end: enqueue now, collect with filter_end()."""
        self._reg.voxelgrid_filter_begin(self._input, self._leaf, out, self.min_points_per_voxel, self.is_dense)

    def filter_end(self):
        return self._reg.voxelgrid_filter_end()

    def setSmBudget(self, n_sm):
        self._reg.setSmBudget(n_sm)

    def last_layout(self, n_voxels, n_points):
        return self._reg.voxelgrid_last_layout(n_voxels, n_points)


class RadiusOutlierRemoval:
    """pcl::RadiusOutlierRemoval<PointXYZ> as the prefiltering nodelet sets it up
    [REF apps/prefiltering_nodelet.cpp:88-96,262-273]: setRadiusSearch + setMinNeighborsInRadius + filter."""

    def __init__(self, device=0, registration=None):
        self._reg = registration if registration is not None else Registration(device=device)
        self.radius = 0.8
        self.min_neighbors = 2
        self._input = None

    def setRadiusSearch(self, r):
        self.radius = float(r)

    def setMinNeighborsInRadius(self, n):
        self.min_neighbors = int(n)

    def setInputCloud(self, cloud):
        self._input = cloud

    def filter(self, out=None):
        return self._reg.radius_outlier_removal(self._input, self.radius, self.min_neighbors, out=out)

    def filter_begin(self, out):
        self._reg.radius_outlier_removal_begin(self._input, self.radius, self.min_neighbors, out)

    def filter_end(self):
        return self._reg.radius_outlier_removal_end()


class StatisticalOutlierRemoval:
    """pcl::StatisticalOutlierRemoval<PointXYZ> as the prefiltering nodelet sets it up (its default outlier filter)
    [REF apps/prefiltering_nodelet.cpp:77-87,262-273]: setMeanK + setStddevMulThresh + filter."""

    def __init__(self, device=0, registration=None):
        self._reg = registration if registration is not None else Registration(device=device)
        self.mean_k = 1          # pcl's own defaults; the nodelet sets 20 / 1.0
        self.stddev_mul = 0.0
        self._input = None

    def setMeanK(self, k):
        self.mean_k = int(k)

    def setStddevMulThresh(self, m):
        self.stddev_mul = float(m)

    def setInputCloud(self, cloud):
        self._input = cloud

    def filter(self, out=None):
        return self._reg.statistical_outlier_removal(self._input, self.mean_k, self.stddev_mul, out=out)

    def filter_begin(self, out):
        self._reg.statistical_outlier_removal_begin(self._input, self.mean_k, self.stddev_mul, out)

    def filter_end(self):
        return self._reg.statistical_outlier_removal_end()

    def last_stats(self, n_points=0):
        return self._reg.statistical_last_stats(n_points)


def select_registration_method(params=None, device=0, out=sys.stdout):
    """hdl_graph_slam::select_registration_method(ros::NodeHandle&) over a dict of private params."""
    p = dict(params or {})
    registration_method = p.get("registration_method", "NDT_OMP")
    if registration_method == "FAST_GICP":
        print("registration: FAST_GICP", file=out)
        gicp = FastGICP(device=device)
        gicp.setNumThreads(p.get("reg_num_threads", 0))
        gicp.setTransformationEpsilon(p.get("reg_transformation_epsilon", 0.01))
        gicp.setMaximumIterations(p.get("reg_maximum_iterations", 64))
        gicp.setMaxCorrespondenceDistance(p.get("reg_max_correspondence_distance", 2.5))
        gicp.setCorrespondenceRandomness(p.get("reg_correspondence_randomness", 20))
        return gicp
    if registration_method in ("FAST_VGICP", "FAST_VGICP_CUDA", "ICP") or "GICP" in registration_method:
        # selectable in the reference, not on the path this engine replaces:
        # a maintainer keeps the reference's own branch for these.
        raise NotImplementedError(f"registration_method={registration_method} stays on the reference's CPU implementation; "
                                  "b200reg replaces NDT_OMP and FAST_GICP")
    if "NDT" not in registration_method:
        print(f"warning: unknown registration type({registration_method})", file=sys.stderr)
        print("       : use NDT", file=sys.stderr)
    ndt_resolution = p.get("reg_resolution", 0.5)
    if "OMP" not in registration_method:
        # "NDT" and every unknown string end in the reference's single-thread pcl::NormalDistributionsTransform branch
        # [REF src/hdl_graph_slam/registrations.cpp:88-99], which is not on the path this engine replaces
        raise NotImplementedError(f"registration_method={registration_method} selects the reference's single-thread pcl::NormalDistributionsTransform, which stays on its CPU implementation; "
                                  "b200reg replaces NDT_OMP and FAST_GICP")
    num_threads = p.get("reg_num_threads", 0)
    nn_search_method = p.get("reg_nn_search_method", "DIRECT7")
    print(f"registration: NDT_OMP {nn_search_method} {ndt_resolution:g} ({num_threads} threads)", file=out)
    ndt = NormalDistributionsTransform(device=device)
    if num_threads > 0:
        ndt.setNumThreads(num_threads)
    ndt.setTransformationEpsilon(p.get("reg_transformation_epsilon", 0.01))
    ndt.setMaximumIterations(p.get("reg_maximum_iterations", 64))
    ndt.setResolution(ndt_resolution)
    if nn_search_method == "KDTREE":
        ndt.setNeighborhoodSearchMethod(_lib.KDTREE)
    elif nn_search_method == "DIRECT1":
        ndt.setNeighborhoodSearchMethod(_lib.DIRECT1)
    else:
        ndt.setNeighborhoodSearchMethod(_lib.DIRECT7)
    return ndt
