"""ORACLE — test infrastructure only.

ctypes binding of oracle/liboracle.so (the CPU restatement of the reference's
scan-registration path, see oracle/oracle.hpp).  Imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs;
never by delta_graph_slam_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

KDTREE, DIRECT26, DIRECT7, DIRECT1 = 0, 1, 2, 3
NDT, GICP = 0, 1
REG_NONE, REG_MIN_EIG, REG_NORMALIZED_MIN_EIG, REG_PLANE, REG_FROBENIUS = range(5)
HDL64, DENSE128 = 0, 1


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    if force or not os.path.exists(so):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"] + (["-B"] if force else []))
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        f32p, f64p, i32p, u32p, u64p = (np.ctypeslib.ndpointer(dtype=t, flags="C_CONTIGUOUS") for t in (np.float32, np.float64, np.int32, np.uint32, np.uint64))
        L.orc_max_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
        L.orc_set_num_threads.restype = None
        L.orc_voxelgrid.restype = C.c_longlong
        L.orc_voxelgrid.argtypes = [f32p, C.c_longlong, C.c_float, C.c_float, C.c_float, C.c_uint, C.c_int, f32p, u32p, u32p, u32p, i32p, i32p]
        L.orc_reg_create.restype = C.c_void_p
        L.orc_reg_create.argtypes = [C.c_int]
        L.orc_reg_destroy.argtypes = [C.c_void_p]
        L.orc_reg_set.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        L.orc_reg_set.restype = C.c_int
        L.orc_reg_set_target.argtypes = [C.c_void_p, f32p, C.c_longlong]
        L.orc_reg_set_source.argtypes = [C.c_void_p, f32p, C.c_longlong]
        L.orc_reg_align.argtypes = [C.c_void_p, f32p, C.c_void_p]
        L.orc_reg_align.restype = C.c_int
        L.orc_reg_get_result.argtypes = [C.c_void_p, f32p, i32p, i32p, f64p]
        L.orc_reg_fitness.argtypes = [C.c_void_p, C.c_double]
        L.orc_reg_fitness.restype = C.c_double
        L.orc_reg_inlier_fraction.argtypes = [C.c_void_p, f32p, C.c_longlong, C.c_double]
        L.orc_reg_inlier_fraction.restype = C.c_double
        L.orc_ndt_num_leaves.argtypes = [C.c_void_p]
        L.orc_ndt_num_leaves.restype = C.c_longlong
        L.orc_ndt_grid.argtypes = [C.c_void_p, i32p, i32p]
        L.orc_ndt_get_leaves.argtypes = [C.c_void_p, u64p, i32p, f64p, f64p, f64p, f32p]
        L.orc_ndt_derivatives.argtypes = [C.c_void_p, f64p, f64p, f64p, C.c_int]
        L.orc_ndt_derivatives.restype = C.c_double
        L.orc_ndt_trace.argtypes = [C.c_void_p, f64p, C.c_longlong]
        L.orc_ndt_trace.restype = C.c_longlong
        L.orc_ndt_hessian.argtypes = [C.c_void_p, f64p, f64p]
        L.orc_ndt_hessian.restype = None
        L.orc_gicp_covariances.argtypes = [C.c_void_p, C.c_int, f64p]
        L.orc_knn.argtypes = [f32p, C.c_longlong, f32p, C.c_longlong, C.c_int, i32p, f32p]
        L.orc_distance_filter.argtypes = [f32p, C.c_longlong, C.c_double, C.c_double, f32p]
        L.orc_distance_filter.restype = C.c_longlong
        L.orc_transform_cloud_d.argtypes = [f32p, C.c_longlong, np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS"), C.c_int, f32p]
        L.orc_transform_cloud_d.restype = None
        L.orc_radius_outlier_removal.argtypes = [f32p, C.c_longlong, C.c_double, C.c_int, f32p]
        L.orc_radius_outlier_removal.restype = C.c_longlong
        L.orc_statistical_outlier_removal.argtypes = [f32p, C.c_longlong, C.c_int, C.c_double, f32p, f64p, f32p]
        L.orc_statistical_outlier_removal.restype = C.c_longlong
        L.orc_flat_filter.argtypes = [f32p, C.c_longlong, C.c_double, C.c_int, C.c_float, f32p, f32p]
        L.orc_flat_filter.restype = C.c_longlong
        L.orc_map_cloud.argtypes = [f32p, np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS"), C.c_longlong, f32p, C.c_double, f32p, f64p]
        L.orc_map_cloud.restype = C.c_longlong
        L.orc_sym_eigen3.argtypes = [f64p, f64p, f64p]
        L.orc_inverse3.argtypes = [f64p, f64p]
        L.orc_svd_solve6.argtypes = [f64p, f64p, f64p]
        L.orc_ldlt_solve6.argtypes = [f64p, f64p, f64p]
        L.orc_euler_xyz.argtypes = [f32p, f32p]
        L.orc_transform_from_p.argtypes = [f64p, f32p]
        L.orc_synth_scan.argtypes = [C.c_int, C.c_ulonglong, C.c_ulonglong, f64p, f32p]
        L.orc_synth_scan.restype = C.c_longlong
        L.orc_synth_num_rays.argtypes = [C.c_int]
        L.orc_synth_num_rays.restype = C.c_longlong
        L.orc_synth_traj.argtypes = [C.c_longlong, C.c_ulonglong, f64p]
        L.orc_synth_pose.argtypes = [f64p, f64p]
        _LIB = L
    return _LIB


def _cloud(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2 and a.shape[1] == 4, "clouds are (N, 4) float32 (x, y, z, 1)"
    return a


def _colmajor(T):
    """4x4 (numpy, row/col indexing) -> 16 floats in Eigen column-major order."""
    return np.ascontiguousarray(np.asarray(T, dtype=np.float32).T).reshape(16)


def _from_colmajor(v):
    return np.array(v, dtype=np.float32).reshape(4, 4).T.copy()


def voxelgrid(cloud, leaf, min_points_per_voxel=0, is_dense=False):
    """pcl::VoxelGrid::filter.  Returns dict(out, voxel_id, count, key, min_b, div_b, overflow)."""
    cloud = _cloud(cloud)
    n = cloud.shape[0]
    leaf3 = (leaf, leaf, leaf) if np.isscalar(leaf) else tuple(leaf)
    out = np.empty((max(n, 1), 4), np.float32)
    vid = np.empty(max(n, 1), np.uint32)
    cnt = np.empty(max(n, 1), np.uint32)
    key = np.empty(max(n, 1), np.uint32)
    grid = np.zeros(6, np.int32)
    ovf = np.zeros(1, np.int32)
    m = lib().orc_voxelgrid(cloud if n else np.zeros((1, 4), np.float32), n, leaf3[0], leaf3[1], leaf3[2], min_points_per_voxel, int(is_dense), out, vid, cnt, key, grid, ovf)
    if ovf[0]:
        return dict(out=out[:m].copy(), voxel_id=vid[:0], count=cnt[:0], key=key[:n].copy(), min_b=grid[:3], div_b=grid[3:], overflow=True)
    return dict(out=out[:m].copy(), voxel_id=vid[:m].copy(), count=cnt[:m].copy(), key=key[:n].copy(), min_b=grid[:3].copy(), div_b=grid[3:].copy(), overflow=False)


def distance_filter(cloud, near_thresh=1.0, far_thresh=100.0):
    """PrefilteringNodelet::distance_filter [REF apps/prefiltering_nodelet.cpp:275-291]."""
    cloud = _cloud(cloud)
    out = np.empty((max(len(cloud), 1), 4), np.float32)
    m = lib().orc_distance_filter(cloud if len(cloud) else np.zeros((1, 4), np.float32), len(cloud), float(near_thresh), float(far_thresh), out)
    return out[:m].copy()


def transform_cloud_d(cloud, matrix4x4, is_dense=False):
    """pcl::transformPointCloud(cloud, out, Matrix4d) as the prefilter's base_link step calls it
    [REF apps/prefiltering_nodelet.cpp:137-147]."""
    cloud = _cloud(cloud)
    out = np.empty((max(len(cloud), 1), 4), np.float32)
    m = np.ascontiguousarray(np.asarray(matrix4x4, np.float64).reshape(4, 4).T).reshape(-1)  # column-major
    lib().orc_transform_cloud_d(cloud if len(cloud) else np.zeros((1, 4), np.float32), len(cloud), m, int(is_dense), out)
    return out[: len(cloud)].copy()


def radius_outlier_removal(cloud, radius=0.8, min_neighbors=2):
    """pcl::RadiusOutlierRemoval as configured at [REF apps/prefiltering_nodelet.cpp:88-96]."""
    cloud = _cloud(cloud)
    out = np.empty((max(len(cloud), 1), 4), np.float32)
    m = lib().orc_radius_outlier_removal(cloud if len(cloud) else np.zeros((1, 4), np.float32), len(cloud), float(radius), int(min_neighbors), out)
    return out[:m].copy()


def statistical_outlier_removal(cloud, mean_k=20, stddev_mul=1.0, details=False):
    """pcl::StatisticalOutlierRemoval as configured at [REF apps/prefiltering_nodelet.cpp:77-87] (the nodelet's default
    outlier_removal_method).  details=True also returns {mean, stddev, threshold} and the per-point mean distances."""
    cloud = _cloud(cloud)
    n = len(cloud)
    out = np.empty((max(n, 1), 4), np.float32)
    stats = np.zeros(3, np.float64)
    dist = np.zeros(max(n, 1), np.float32)
    m = lib().orc_statistical_outlier_removal(cloud if n else np.zeros((1, 4), np.float32), n, int(mean_k), float(stddev_mul), out, stats, dist)
    if details:
        return out[:m].copy(), dict(mean=stats[0], stddev=stats[1], threshold=stats[2], distances=dist[:n].copy())
    return out[:m].copy()


def flat_filter(cloud, lidar_z, k=10, thresh=0.2, details=False):
    """filtered2D of the prefiltering nodelet [REF apps/prefiltering_nodelet.cpp:155-158]: height_filtering (z > lidar z) ->
    normal_filtering (pcl::NormalEstimation k = 10, |n_z| < 0.2) -> flatten (z := 0).  details=True also returns |n_z| per input point."""
    cloud = _cloud(cloud)
    n = len(cloud)
    out = np.empty((max(n, 1), 4), np.float32)
    nz = np.zeros(max(n, 1), np.float32)
    m = lib().orc_flat_filter(cloud if n else np.zeros((1, 4), np.float32), n, float(lidar_z), int(k), float(thresh), out, nz)
    if details:
        return out[:m].copy(), nz[:n].copy()
    return out[:m].copy()


def map_cloud(clouds, poses, resolution, details=False):
    """MapCloudGenerator::generate [REF src/hdl_graph_slam/map_cloud_generator.cpp:13-49]: keyframe clouds transformed by their
    4x4 float poses, concatenated, and (resolution > 0) reduced to pcl's occupied octree voxel centres in its own order."""
    clouds = [_cloud(c) for c in clouds]
    offsets = np.zeros(len(clouds) + 1, np.int64)
    offsets[1:] = np.cumsum([len(c) for c in clouds])
    cat = np.concatenate(clouds) if clouds and offsets[-1] else np.zeros((1, 4), np.float32)
    T = np.ascontiguousarray(np.stack([np.asarray(p, np.float32).T.reshape(16) for p in poses]) if len(poses) else np.zeros((1, 16), np.float32), np.float32)  # column-major
    out = np.empty((max(int(offsets[-1]), 1), 4), np.float32)
    info = np.zeros(4, np.float64)
    m = lib().orc_map_cloud(np.ascontiguousarray(cat), offsets, len(clouds), T, float(resolution), out, info)
    if details:
        return out[:m].copy(), dict(min=info[:3].copy(), depth=int(info[3]))
    return out[:m].copy()


class Registration:
    """Oracle registration object with the pcl::Registration call surface the reference uses."""

    def __init__(self, method=NDT, **params):
        self._h = lib().orc_reg_create(method)
        self.method = method
        self._src_n = 0
        for k, v in params.items():
            self.set(k, v)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_reg_destroy(self._h)
            self._h = None

    def set(self, name, value):
        if lib().orc_reg_set(self._h, name.encode(), float(value)) != 0:
            raise KeyError(name)

    def setInputTarget(self, cloud):
        c = _cloud(cloud)
        lib().orc_reg_set_target(self._h, c if len(c) else np.zeros((1, 4), np.float32), len(c))

    def setInputSource(self, cloud):
        c = _cloud(cloud)
        self._src_n = len(c)
        lib().orc_reg_set_source(self._h, c if len(c) else np.zeros((1, 4), np.float32), len(c))

    def align(self, guess=None, want_aligned=False):
        g = _colmajor(np.eye(4) if guess is None else guess)
        if want_aligned:
            out = np.zeros((max(self._src_n, 1), 4), np.float32)
            lib().orc_reg_align(self._h, g, out.ctypes.data_as(C.c_void_p))
            return out[: self._src_n]
        lib().orc_reg_align(self._h, g, None)
        return None

    def _result(self):
        T = np.zeros(16, np.float32)
        conv = np.zeros(1, np.int32)
        it = np.zeros(1, np.int32)
        info = np.zeros(3, np.float64)
        lib().orc_reg_get_result(self._h, T, conv, it, info)
        return _from_colmajor(T), bool(conv[0]), int(it[0]), info

    def getFinalTransformation(self):
        return self._result()[0]

    def hasConverged(self):
        return self._result()[1]

    def getFinalNumIteration(self):
        return self._result()[2]

    def info(self):
        return self._result()[3]

    def getFitnessScore(self, max_range=np.finfo(np.float64).max):
        return lib().orc_reg_fitness(self._h, float(max_range))

    def inlierFraction(self, aligned, max_dist=0.5):
        a = _cloud(aligned)
        return lib().orc_reg_inlier_fraction(self._h, a, len(a), max_dist)

    # ---- NDT introspection
    def ndt_leaves(self):
        n = lib().orc_ndt_num_leaves(self._h)
        idx = np.zeros(max(n, 1), np.uint64)
        npts = np.zeros(max(n, 1), np.int32)
        mean = np.zeros((max(n, 1), 3))
        cov = np.zeros((max(n, 1), 9))
        icov = np.zeros((max(n, 1), 9))
        cen = np.zeros((max(n, 1), 3), np.float32)
        lib().orc_ndt_get_leaves(self._h, idx, npts, mean, cov, icov, cen)
        mb = np.zeros(3, np.int32)
        db = np.zeros(3, np.int32)
        lib().orc_ndt_grid(self._h, mb, db)
        return dict(idx=idx[:n], n=npts[:n], mean=mean[:n], cov=cov[:n].reshape(-1, 3, 3), icov=icov[:n].reshape(-1, 3, 3), centroid=cen[:n], min_b=mb, div_b=db)

    def ndt_derivatives(self, p, compute_hessian=True):
        g = np.zeros(6)
        H = np.zeros(36)
        s = lib().orc_ndt_derivatives(self._h, np.ascontiguousarray(p, np.float64), g, H, int(compute_hessian))
        return s, g, H.reshape(6, 6)

    def ndt_trace(self, cap=1024):
        """Line-search evaluations of the last NDT align: rows {nr_iterations, step_iterations, a_t, score, phi_t, d_phi_t, psi_t, d_psi_t,
        open_interval, interval_converged, phi_0, d_phi_0}."""
        out = np.zeros((cap, 12))
        n = lib().orc_ndt_trace(self._h, out, cap)
        return out[: min(n, cap)]

    def ndt_hessian(self, p):
        """pclomp computeHessian (double arithmetic, serial) at pose p: what closes a More-Thuente line search upstream."""
        H = np.zeros(36)
        lib().orc_ndt_hessian(self._h, np.ascontiguousarray(p, np.float64), H)
        return H.reshape(6, 6)

    def gicp_covariances(self, which, n):
        out = np.zeros((max(n, 1), 9))
        lib().orc_gicp_covariances(self._h, which, out)
        return out[:n].reshape(-1, 3, 3)


def knn(points, queries, k):
    p, q = _cloud(points), _cloud(queries)
    idx = np.zeros((len(q), k), np.int32)
    d2 = np.zeros((len(q), k), np.float32)
    lib().orc_knn(p, len(p), q, len(q), k, idx, d2)
    return idx, d2


def transform_from_p(p):
    T = np.zeros(16, np.float32)
    lib().orc_transform_from_p(np.ascontiguousarray(p, np.float64), T)
    return _from_colmajor(T)


def euler_xyz(T):
    out = np.zeros(3, np.float32)
    lib().orc_euler_xyz(_colmajor(T), out)
    return out


# ---- synthetic scans (CPU build of the generator) ---------------------------
def synth_traj(k, seed=7):
    T = np.zeros(16)
    lib().orc_synth_traj(k, seed, T)
    return T.reshape(4, 4)


def synth_pose(xyzrpy):
    T = np.zeros(16)
    lib().orc_synth_pose(np.ascontiguousarray(xyzrpy, np.float64), T)
    return T.reshape(4, 4)


def synth_scan(pose, sensor=HDL64, scene_seed=1, noise_seed=1000):
    rays = lib().orc_synth_num_rays(sensor)
    out = np.zeros((rays, 4), np.float32)
    n = lib().orc_synth_scan(sensor, scene_seed, noise_seed, np.ascontiguousarray(pose, np.float64).reshape(16), out)
    return out[:n].copy()
