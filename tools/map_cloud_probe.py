"""MapCloudGenerator::generate [REF src/hdl_graph_slam/map_cloud_generator.cpp:13-49] at a realistic size — n keyframes of a
down-sampled HDL-64 sequence (4 m apart), map resolution 0.05 m (the reference's map_cloud_resolution default) — through
host clouds (b200reg_map_cloud) and from the keyframe cache (b200reg_map_cloud_cached), next to the oracle on the host CPU.
python tools/map_cloud_probe.py [n_keyframes]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import delta_graph_slam_b200 as eng  # noqa: E402
from delta_graph_slam_b200.map_cloud_generator import KeyFrameSnapshot, MapCloudGenerator  # noqa: E402
from oracle import oracle_py as O  # noqa: E402  (input generation, the CPU leg and the check only)


def main():
    n_kf = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    res = 0.05
    poses = [O.synth_traj(8 * k) for k in range(n_kf)]
    clouds = [O.voxelgrid(O.synth_scan(poses[k], noise_seed=3000 + k), 0.1)["out"] for k in range(n_kf)]
    kfs = [KeyFrameSnapshot(poses[k], clouds[k]) for k in range(n_kf)]
    gen = MapCloudGenerator()
    t_host, t_cached = [], []
    for rep in range(4):
        t0 = time.perf_counter()
        out = gen.generate(kfs, res)
        t_host.append((time.perf_counter() - t0) * 1e3)
    for k in range(n_kf):
        gen._reg.cloudPut(k, clouds[k])
    gen._reg.cloudSync()
    for rep in range(4):
        t0 = time.perf_counter()
        out_c = gen.generate_cached(list(range(n_kf)), poses, res)
        t_cached.append((time.perf_counter() - t0) * 1e3)
    t0 = time.perf_counter()
    want = O.map_cloud(clouds, poses, res)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    same = lambda a, b: len(a) == len(b) and np.array_equal(np.asarray(a).view(np.uint32), np.asarray(b).view(np.uint32))
    print(json.dumps(dict(keyframes=n_kf, points_in=int(sum(len(c) for c in clouds)), points_out=int(len(want)), resolution=res,
                          map_cloud_ms_host_clouds=float(np.median(t_host[1:])), map_cloud_ms_cached_keyframes=float(np.median(t_cached[1:])),
                          oracle_ms=cpu_ms, oracle_cores=1, bit_identical=bool(same(out, want) and same(out_c, want)))))


if __name__ == "__main__":
    main()
