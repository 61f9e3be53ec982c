"""Host-side mirror of hdl_graph_slam::MapCloudGenerator [REF src/hdl_graph_slam/map_cloud_generator.cpp:13-49;
include/hdl_graph_slam/map_cloud_generator.hpp]: generate(keyframes, resolution) over the engine's b200reg_map_cloud.

A keyframe snapshot is anything with `.cloud` ((N, 4) float32) and `.pose` (4x4, the Isometry3d of KeyFrameSnapshot
[REF include/hdl_graph_slam/keyframe.hpp:68-84]; cast to float as the reference does, :23)."""
import ctypes as C
import sys

import numpy as np

from . import _lib
from .registration import Registration


class KeyFrameSnapshot:
    def __init__(self, pose, cloud):
        self.pose = np.asarray(pose, np.float64)
        self.cloud = cloud


class MapCloudGenerator:
    def __init__(self, device=0, registration=None):
        self._reg = registration if registration is not None else Registration(device=device)

    def generate(self, keyframes, resolution, details=False):
        """The map cloud ((M, 4) float32), or None for an empty keyframe list (the reference prints a warning and
        returns nullptr, :14-17)."""
        if not keyframes:
            print("warning: keyframes empty!!", file=sys.stderr)
            return None
        clouds = [_lib.as_cloud(k.cloud) for k in keyframes]
        n = len(clouds)
        ptrs = (C.c_void_p * n)(*[c.ctypes.data if len(c) else None for c in clouds])
        counts = (C.c_size_t * n)(*[len(c) for c in clouds])
        poses = np.ascontiguousarray(np.stack([np.asarray(k.pose, np.float64).astype(np.float32).T.reshape(16) for k in keyframes]), np.float32)
        total = int(sum(len(c) for c in clouds))
        out = np.empty((max(total, 1), 4), np.float32)
        n_out = C.c_size_t()
        info = np.zeros(4, np.float64)
        self._reg._ck(_lib.load().b200reg_map_cloud(self._reg._h, ptrs, counts, poses.ctypes.data, n, float(resolution), out.ctypes.data, len(out), C.byref(n_out), info.ctypes.data))
        res = out[: n_out.value].copy()
        if details:
            return res, dict(min=info[:3].copy(), depth=int(info[3]))
        return res

    def generate_cached(self, keyframe_ids, poses, resolution):
        """The same over keyframe clouds already cached on the device under their ids (Registration.cloudPut)."""
        ids = np.ascontiguousarray(keyframe_ids, np.int64)
        P = np.ascontiguousarray(np.stack([np.asarray(p, np.float64).astype(np.float32).T.reshape(16) for p in poses]), np.float32)
        cap = 1 << 20
        while True:
            out = np.empty((cap, 4), np.float32)
            n_out = C.c_size_t()
            rc = _lib.load().b200reg_map_cloud_cached(self._reg._h, ids.ctypes.data, P.ctypes.data, len(ids), float(resolution), out.ctypes.data, len(out), C.byref(n_out), None)
            if rc == _lib.E_CAPACITY:
                cap = int(n_out.value)
                continue
            self._reg._ck(rc)
            return out[: n_out.value].copy()
