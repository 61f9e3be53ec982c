"""ctypes binding of libb200reg.so (include/b200reg.h).

The product path: there is no CPU fallback.  If the CUDA library is missing or cannot be loaded
this module raises; if no B200 is visible `b200reg_create` fails and the wrappers raise
`B200RegError`.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200REG_LIB") or os.path.join(_HERE, "libb200reg.so")  # B200REG_LIB: developer override (A/B builds of the same engine)

OK, E_INVALID, E_CUDA, E_STATE, E_CAPACITY = 0, -1, -2, -3, -4
METHOD_NONE, METHOD_NDT, METHOD_GICP = 0, 1, 2
KDTREE, DIRECT26, DIRECT7, DIRECT1 = 0, 1, 2, 3
REG_NONE, REG_MIN_EIG, REG_NORMALIZED_MIN_EIG, REG_PLANE, REG_FROBENIUS = range(5)
LSQ_GN, LSQ_LM = 0, 1

# every symbol include/b200reg.h declares (tests check the library exports each one)
EXPORTS = [
    "b200reg_default_config", "b200reg_create", "b200reg_destroy", "b200reg_last_error", "b200reg_version",
    "b200reg_set_resolution", "b200reg_set_nn_search", "b200reg_set_transformation_epsilon", "b200reg_set_maximum_iterations",
    "b200reg_set_max_correspondence_distance", "b200reg_set_correspondence_randomness", "b200reg_set_gicp_options", "b200reg_gicp_get_covariances",
    "b200reg_set_target", "b200reg_set_source", "b200reg_set_target_device", "b200reg_set_source_device", "b200reg_promote_source_to_target", "b200reg_prepare_promotion", "b200reg_set_side_budget",
    "b200reg_align", "b200reg_has_converged", "b200reg_get_final_transformation", "b200reg_get_num_iterations",
    "b200reg_get_transformation_probability", "b200reg_get_result", "b200reg_get_fitness_score", "b200reg_calc_fitness_score", "b200reg_get_inlier_fraction",
    "b200reg_voxelgrid_filter", "b200reg_voxelgrid_filter_device", "b200reg_voxelgrid_last_layout",
    "b200reg_voxelgrid_filter_begin", "b200reg_voxelgrid_filter_device_begin", "b200reg_voxelgrid_filter_host_to_device_begin", "b200reg_voxelgrid_filter_end",
    "b200reg_odometry_default_config", "b200reg_odometry_create", "b200reg_odometry_destroy", "b200reg_odometry_last_error", "b200reg_odometry_reset", "b200reg_odometry_matching",
    "b200reg_odometry_matching_device", "b200reg_odometry_get_state",
    "b200reg_frontend_default_config", "b200reg_frontend_create", "b200reg_frontend_destroy", "b200reg_frontend_last_error", "b200reg_frontend_reset", "b200reg_frontend_registration",
    "b200reg_frontend_filter", "b200reg_frontend_odometry", "b200reg_frontend_begin", "b200reg_frontend_begin_device", "b200reg_frontend_step", "b200reg_frontend_step_device", "b200reg_frontend_run_device", "b200reg_frontend_get_timing", "b200reg_set_sm_budget", "b200reg_set_distance_filter", "b200reg_set_input_transform", "b200reg_distance_filter", "b200reg_distance_filter_device",
    "b200reg_radius_outlier_removal", "b200reg_radius_outlier_removal_device", "b200reg_radius_outlier_removal_begin", "b200reg_radius_outlier_removal_device_begin",
    "b200reg_radius_outlier_removal_end",
    "b200reg_statistical_outlier_removal", "b200reg_statistical_outlier_removal_device", "b200reg_statistical_outlier_removal_begin", "b200reg_statistical_outlier_removal_device_begin",
    "b200reg_statistical_outlier_removal_end", "b200reg_statistical_last_stats",
    "b200reg_flat_filter", "b200reg_flat_filter_device", "b200reg_flat_filter_begin", "b200reg_flat_filter_device_begin", "b200reg_flat_filter_end", "b200reg_flat_filter_last_nz",
    "b200reg_cloud_put", "b200reg_cloud_sync", "b200reg_cloud_put_device", "b200reg_set_source_cached", "b200reg_set_target_cached", "b200reg_cloud_drop", "b200reg_cloud_clear", "b200reg_cloud_count", "b200reg_align_batch", "b200reg_calc_fitness_batch", "b200reg_get_batch_timing",
    "b200reg_map_cloud", "b200reg_map_cloud_cached",
    "b200reg_batch_create", "b200reg_batch_destroy", "b200reg_batch_last_error", "b200reg_batch_cloud_put", "b200reg_batch_cloud_drop", "b200reg_batch_run", "b200reg_batch_get_info",
    "b200reg_ndt_num_leaves", "b200reg_ndt_get_leaves", "b200reg_ndt_derivatives", "b200reg_set_timing", "b200reg_get_counters", "b200reg_get_profile", "b200reg_set_profile", "b200reg_get_trace", "b200reg_set_sort_path", "b200reg_get_nn_stats", "b200reg_get_stream",
]


class Config(C.Structure):
    _fields_ = [
        ("device", C.c_int), ("method", C.c_int), ("resolution", C.c_double), ("nn_search", C.c_int),
        ("transformation_epsilon", C.c_double), ("maximum_iterations", C.c_int), ("step_size", C.c_double),
        ("outlier_ratio", C.c_double), ("max_correspondence_distance", C.c_double), ("correspondence_randomness", C.c_int),
        ("rotation_epsilon", C.c_double), ("regularization", C.c_int), ("lsq_optimizer", C.c_int), ("num_threads", C.c_int),
    ]


class OdometryConfig(C.Structure):
    _fields_ = [("keyframe_delta_trans", C.c_double), ("keyframe_delta_angle", C.c_double), ("keyframe_delta_time", C.c_double), ("transform_thresholding", C.c_int),
                ("max_acceptable_trans", C.c_double), ("max_acceptable_angle", C.c_double)]


class FrontEndConfig(C.Structure):
    _fields_ = [("device", C.c_int), ("registration", Config), ("odometry", OdometryConfig), ("downsample_resolution", C.c_double), ("use_distance_filter", C.c_int),
                ("distance_near_thresh", C.c_double), ("distance_far_thresh", C.c_double), ("filter_sms", C.c_int), ("prepare_promotion", C.c_int), ("side_sms", C.c_int)]


class Result(C.Structure):
    _fields_ = [
        ("transformation", C.c_float * 16), ("fitness", C.c_double), ("score", C.c_double), ("converged", C.c_int32),
        ("iterations", C.c_int32), ("evaluations", C.c_int32), ("passes", C.c_int32), ("hits", C.c_int64),
    ]


class Pair(C.Structure):
    _fields_ = [("target_id", C.c_int64), ("source_id", C.c_int64), ("guess", C.c_float * 16)]


RESULT_DTYPE = np.dtype([("transformation", np.float32, (16,)), ("fitness", np.float64), ("score", np.float64), ("converged", np.int32), ("iterations", np.int32),
                         ("evaluations", np.int32), ("passes", np.int32), ("hits", np.int64)])
PAIR_DTYPE = np.dtype([("target_id", np.int64), ("source_id", np.int64), ("guess", np.float32, (16,))])
assert RESULT_DTYPE.itemsize == C.sizeof(Result) and PAIR_DTYPE.itemsize == C.sizeof(Pair)


class B200RegError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"b200reg error {code}: {msg}")
        self.code = code


_lib = None


def load():
    """Load libb200reg.so; raises OSError (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OSError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` (make -C delta_graph_slam_b200/csrc)")
    L = C.CDLL(LIB_PATH)
    vp, szp = C.c_void_p, C.POINTER(C.c_size_t)
    L.b200reg_version.restype = C.c_char_p
    L.b200reg_last_error.restype = C.c_char_p
    L.b200reg_last_error.argtypes = [vp]
    L.b200reg_default_config.argtypes = [C.c_int, C.POINTER(Config)]
    L.b200reg_default_config.restype = None
    L.b200reg_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.b200reg_destroy.argtypes = [vp]
    for name, extra in [("b200reg_set_resolution", C.c_double), ("b200reg_set_nn_search", C.c_int), ("b200reg_set_transformation_epsilon", C.c_double),
                        ("b200reg_set_maximum_iterations", C.c_int), ("b200reg_set_max_correspondence_distance", C.c_double),
                        ("b200reg_set_correspondence_randomness", C.c_int)]:
        getattr(L, name).argtypes = [vp, extra]
    L.b200reg_set_gicp_options.argtypes = [vp, C.c_int, C.c_int, C.c_double]
    L.b200reg_gicp_get_covariances.argtypes = [vp, C.c_int, vp, C.c_size_t]
    L.b200reg_set_target.argtypes = [vp, vp, C.c_size_t, C.c_size_t]
    L.b200reg_set_source.argtypes = [vp, vp, C.c_size_t, C.c_size_t]
    L.b200reg_set_target_device.argtypes = [vp, vp, C.c_size_t]
    L.b200reg_set_source_device.argtypes = [vp, vp, C.c_size_t]
    L.b200reg_promote_source_to_target.argtypes = [vp]
    L.b200reg_prepare_promotion.argtypes = [vp]
    L.b200reg_set_side_budget.argtypes = [vp, C.c_int]
    L.b200reg_align.argtypes = [vp, vp, vp]
    L.b200reg_has_converged.argtypes = [vp, C.POINTER(C.c_int)]
    L.b200reg_get_final_transformation.argtypes = [vp, vp]
    L.b200reg_get_num_iterations.argtypes = [vp, C.POINTER(C.c_int)]
    L.b200reg_get_transformation_probability.argtypes = [vp, C.POINTER(C.c_double)]
    L.b200reg_get_result.argtypes = [vp, C.POINTER(Result)]
    L.b200reg_get_fitness_score.argtypes = [vp, C.c_double, C.POINTER(C.c_double)]
    L.b200reg_calc_fitness_score.argtypes = [vp, vp, C.c_double, C.POINTER(C.c_double)]
    L.b200reg_get_inlier_fraction.argtypes = [vp, C.c_double, C.POINTER(C.c_double)]
    L.b200reg_voxelgrid_filter.argtypes = [vp, vp, C.c_size_t, C.c_size_t, C.POINTER(C.c_float), C.c_uint, C.c_int, vp, C.c_size_t, szp]
    L.b200reg_voxelgrid_filter_device.argtypes = [vp, vp, C.c_size_t, C.POINTER(C.c_float), C.c_uint, C.c_int, vp, szp]
    L.b200reg_voxelgrid_filter_begin.argtypes = [vp, vp, C.c_size_t, C.c_size_t, C.POINTER(C.c_float), C.c_uint, C.c_int, vp, C.c_size_t]
    L.b200reg_voxelgrid_filter_device_begin.argtypes = [vp, vp, C.c_size_t, C.POINTER(C.c_float), C.c_uint, C.c_int, vp]
    L.b200reg_voxelgrid_filter_host_to_device_begin.argtypes = [vp, vp, C.c_size_t, C.c_size_t, C.POINTER(C.c_float), C.c_uint, C.c_int, vp]
    L.b200reg_voxelgrid_filter_end.argtypes = [vp, szp]
    L.b200reg_odometry_default_config.argtypes = [C.POINTER(OdometryConfig)]
    L.b200reg_odometry_default_config.restype = None
    L.b200reg_odometry_create.argtypes = [vp, C.POINTER(OdometryConfig), C.POINTER(vp)]
    L.b200reg_odometry_destroy.argtypes = [vp]
    L.b200reg_odometry_last_error.argtypes = [vp]
    L.b200reg_odometry_last_error.restype = C.c_char_p
    L.b200reg_odometry_reset.argtypes = [vp]
    L.b200reg_odometry_matching.argtypes = [vp, C.c_double, vp, C.c_size_t, C.c_size_t, vp, vp, vp]
    L.b200reg_odometry_matching_device.argtypes = [vp, C.c_double, vp, C.c_size_t, vp, vp]
    L.b200reg_odometry_get_state.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), vp, vp, C.POINTER(Result)]
    L.b200reg_frontend_default_config.argtypes = [C.POINTER(FrontEndConfig)]
    L.b200reg_frontend_default_config.restype = None
    L.b200reg_frontend_create.argtypes = [C.POINTER(FrontEndConfig), C.POINTER(vp)]
    L.b200reg_frontend_destroy.argtypes = [vp]
    L.b200reg_frontend_last_error.argtypes = [vp]
    L.b200reg_frontend_last_error.restype = C.c_char_p
    L.b200reg_frontend_reset.argtypes = [vp]
    for name in ("b200reg_frontend_registration", "b200reg_frontend_filter", "b200reg_frontend_odometry"):
        getattr(L, name).argtypes = [vp]
        getattr(L, name).restype = vp
    L.b200reg_frontend_begin.argtypes = [vp, C.c_double, vp, C.c_size_t, C.c_size_t, vp, C.c_size_t]
    L.b200reg_frontend_begin_device.argtypes = [vp, C.c_double, vp, C.c_size_t]
    L.b200reg_frontend_step.argtypes = [vp, C.c_double, vp, C.c_size_t, C.c_size_t, vp, C.c_size_t, szp, vp, vp]
    L.b200reg_frontend_step_device.argtypes = [vp, C.c_double, vp, C.c_size_t, szp, vp]
    L.b200reg_frontend_get_timing.argtypes = [vp, vp]
    L.b200reg_frontend_run_device.argtypes = [vp, vp, vp, vp, C.c_size_t, vp, vp, vp, C.POINTER(C.c_int)]
    L.b200reg_set_sm_budget.argtypes = [vp, C.c_int]
    L.b200reg_set_distance_filter.argtypes = [vp, C.c_int, C.c_double, C.c_double]
    L.b200reg_set_input_transform.argtypes = [vp, C.POINTER(C.c_double)]
    L.b200reg_distance_filter.argtypes = [vp, vp, C.c_size_t, C.c_size_t, C.c_double, C.c_double, vp, C.c_size_t, szp]
    L.b200reg_distance_filter_device.argtypes = [vp, vp, C.c_size_t, C.c_double, C.c_double, vp, szp]
    L.b200reg_radius_outlier_removal.argtypes = [vp, vp, C.c_size_t, C.c_size_t, C.c_double, C.c_int, vp, C.c_size_t, szp]
    L.b200reg_radius_outlier_removal_device.argtypes = [vp, vp, C.c_size_t, C.c_double, C.c_int, vp, szp]
    L.b200reg_radius_outlier_removal_begin.argtypes = [vp, vp, C.c_size_t, C.c_size_t, C.c_double, C.c_int, vp, C.c_size_t]
    L.b200reg_radius_outlier_removal_device_begin.argtypes = [vp, vp, C.c_size_t, C.c_double, C.c_int, vp]
    L.b200reg_radius_outlier_removal_end.argtypes = [vp, szp]
    L.b200reg_statistical_outlier_removal.argtypes = [vp, vp, C.c_size_t, C.c_size_t, C.c_int, C.c_double, vp, C.c_size_t, szp]
    L.b200reg_statistical_outlier_removal_device.argtypes = [vp, vp, C.c_size_t, C.c_int, C.c_double, vp, szp]
    L.b200reg_statistical_outlier_removal_begin.argtypes = [vp, vp, C.c_size_t, C.c_size_t, C.c_int, C.c_double, vp, C.c_size_t]
    L.b200reg_statistical_outlier_removal_device_begin.argtypes = [vp, vp, C.c_size_t, C.c_int, C.c_double, vp]
    L.b200reg_statistical_outlier_removal_end.argtypes = [vp, szp]
    L.b200reg_statistical_last_stats.argtypes = [vp, vp, vp, vp, vp, C.c_size_t]
    L.b200reg_flat_filter.argtypes = [vp, vp, C.c_size_t, C.c_size_t, C.c_double, C.c_int, C.c_double, vp, C.c_size_t, szp]
    L.b200reg_flat_filter_device.argtypes = [vp, vp, C.c_size_t, C.c_double, C.c_int, C.c_double, vp, szp]
    L.b200reg_flat_filter_begin.argtypes = [vp, vp, C.c_size_t, C.c_size_t, C.c_double, C.c_int, C.c_double, vp, C.c_size_t]
    L.b200reg_flat_filter_device_begin.argtypes = [vp, vp, C.c_size_t, C.c_double, C.c_int, C.c_double, vp]
    L.b200reg_flat_filter_end.argtypes = [vp, szp]
    L.b200reg_flat_filter_last_nz.argtypes = [vp, vp, C.c_size_t]
    L.b200reg_voxelgrid_last_layout.argtypes = [vp, vp, vp, C.c_size_t, vp, C.c_size_t, vp, C.POINTER(C.c_int)]
    L.b200reg_cloud_put.argtypes = [vp, C.c_int64, vp, C.c_size_t, C.c_size_t]
    L.b200reg_cloud_put_device.argtypes = [vp, C.c_int64, vp, C.c_size_t]
    L.b200reg_set_source_cached.argtypes = [vp, C.c_int64]
    L.b200reg_set_target_cached.argtypes = [vp, C.c_int64]
    L.b200reg_cloud_drop.argtypes = [vp, C.c_int64]
    L.b200reg_cloud_sync.argtypes = [vp]
    L.b200reg_cloud_clear.argtypes = [vp]
    L.b200reg_cloud_count.argtypes = [vp, szp]
    L.b200reg_align_batch.argtypes = [vp, vp, C.c_size_t, C.c_int, C.c_double, vp]
    L.b200reg_calc_fitness_batch.argtypes = [vp, vp, C.c_size_t, C.c_double, vp]
    L.b200reg_get_batch_timing.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.b200reg_map_cloud.argtypes = [vp, vp, vp, vp, C.c_size_t, C.c_double, vp, C.c_size_t, szp, vp]
    L.b200reg_map_cloud_cached.argtypes = [vp, vp, vp, C.c_size_t, C.c_double, vp, C.c_size_t, szp, vp]
    L.b200reg_batch_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]
    L.b200reg_batch_destroy.argtypes = [vp]
    L.b200reg_batch_last_error.argtypes = [vp]
    L.b200reg_batch_last_error.restype = C.c_char_p
    L.b200reg_batch_cloud_put.argtypes = [vp, C.c_int64, vp, C.c_size_t, C.c_size_t]
    L.b200reg_batch_cloud_drop.argtypes = [vp, C.c_int64]
    L.b200reg_batch_run.argtypes = [vp, vp, C.c_size_t, C.c_int, C.c_double, vp]
    L.b200reg_batch_get_info.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double)]
    L.b200reg_ndt_num_leaves.argtypes = [vp, szp]
    L.b200reg_ndt_get_leaves.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp]
    L.b200reg_ndt_derivatives.argtypes = [vp, vp, C.POINTER(C.c_double), vp, vp]
    L.b200reg_get_profile.argtypes = [vp, vp]
    L.b200reg_set_profile.argtypes = [vp, C.c_int]
    L.b200reg_get_trace.argtypes = [vp, vp, C.c_size_t, szp]
    L.b200reg_get_nn_stats.argtypes = [vp, vp]
    L.b200reg_set_sort_path.argtypes = [C.c_int]
    L.b200reg_set_timing.argtypes = [vp, C.c_int]
    L.b200reg_get_counters.argtypes = [vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.POINTER(C.c_double)]
    L.b200reg_get_stream.argtypes = [vp, C.POINTER(vp)]
    # synthetic scan generator (bench / test infrastructure living in the same library)
    L.b200synth_num_rays.argtypes = [C.c_int]
    L.b200synth_num_rays.restype = C.c_longlong
    L.b200synth_scan_device.argtypes = [C.c_int, C.c_int, C.c_ulonglong, C.c_ulonglong, vp, vp]
    L.b200synth_scan_device.restype = C.c_longlong
    L.b200synth_traj.argtypes = [C.c_longlong, C.c_ulonglong, vp]
    L.b200synth_pose.argtypes = [vp, vp]
    _lib = L
    return L


def check(handle, rc):
    if rc != OK:
        msg = load().b200reg_last_error(handle).decode() if handle else "no handle"
        raise B200RegError(rc, msg)


def as_cloud(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim != 2 or a.shape[1] != 4:
        raise ValueError("clouds are (N, 4) float32 arrays (x, y, z, pad) — the pcl::PointXYZ layout")
    return a


def colmajor(T):
    return np.ascontiguousarray(np.asarray(T, dtype=np.float32).T).reshape(16)


def from_colmajor(v):
    return np.array(v, dtype=np.float32).reshape(4, 4).T.copy()
