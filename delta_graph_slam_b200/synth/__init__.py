"""Synthetic lidar scans (bench / test infrastructure): device build of synth_scene.h."""
import ctypes as C

import numpy as np

from .. import _lib

HDL64, DENSE128 = 0, 1


def num_rays(sensor=HDL64):
    return _lib.load().b200synth_num_rays(sensor)


def traj_kitti_like(k, seed=7):
    T = np.zeros(16)
    _lib.load().b200synth_traj(k, seed, T.ctypes.data)
    return T.reshape(4, 4)


def pose(xyzrpy):
    v = np.ascontiguousarray(xyzrpy, np.float64)
    T = np.zeros(16)
    _lib.load().b200synth_pose(v.ctypes.data, T.ctypes.data)
    return T.reshape(4, 4)


def scan_to_device(d_out_ptr, pose_rowmajor, sensor=HDL64, scene_seed=1, noise_seed=1000, device=0):
    """Ray-cast one scan on the GPU into device memory at d_out_ptr (capacity num_rays float4)."""
    P = np.ascontiguousarray(pose_rowmajor, np.float64).reshape(16)
    n = _lib.load().b200synth_scan_device(device, sensor, scene_seed, noise_seed, P.ctypes.data, C.c_void_p(d_out_ptr))
    if n < 0:
        raise _lib.B200RegError(_lib.E_CUDA, "b200synth_scan_device failed")
    return int(n)
