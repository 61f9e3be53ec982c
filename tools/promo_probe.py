"""Developer probe: what a prepared promotion buys per keyframe switch."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import delta_graph_slam_b200 as eng
from delta_graph_slam_b200 import synth
import bench

frames = 200
rays = synth.num_rays(synth.HDL64)
d_raw = torch.empty((rays, 4), dtype=torch.float32, device="cuda:0")
vg = eng.VoxelGrid(); vg.setLeafSize(0.1, 0.1, 0.1)
d_ds = torch.empty((frames, 70000, 4), dtype=torch.float32, device="cuda:0")
clouds = []
for k in range(frames):
    n = synth.scan_to_device(d_raw.data_ptr(), synth.traj_kitti_like(k), synth.HDL64, 1, 1000 + k, 0)
    vg.setInputCloud(eng.DeviceCloud(d_raw.data_ptr(), n, d_raw), is_dense=False)
    tmp = torch.empty((rays, 4), dtype=torch.float32, device="cuda:0")
    f = vg.filter(out=eng.DeviceCloud(tmp.data_ptr(), rays, tmp))
    d_ds[k, : f.n].copy_(tmp[: f.n]); torch.cuda.synchronize()
    clouds.append(eng.DeviceCloud(d_ds[k].data_ptr(), f.n, d_ds))

for mode in ("off", "predict", "always", "off"):
    odo = eng.ScanMatchingOdometry(dict(bench.ODOM_PARAMS, prepare_promotion=mode != "off"), out=bench.DEVNULL)
    odo.registration.setSmBudget(108); odo.registration.setSideBudget(16)
    for rep in range(2):
        odo.keyframe = None
        t_kf, t_other, n_kf = 0.0, 0.0, 0
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k, c in enumerate(clouds):
            if mode == "always" and k > 0:
                odo._last_step = 10.0
            nk = odo.num_keyframes
            a = time.perf_counter()
            odo.matching(0.1 * k, c)
            b = time.perf_counter() - a
            if odo.num_keyframes != nk: t_kf += b; n_kf += 1
            else: t_other += b
        tot = time.perf_counter() - t0
    print(f"{mode:8s} {tot / frames * 1e6:7.1f} us/frame  keyframe-switch frames {n_kf}: {t_kf / max(n_kf, 1) * 1e6:7.1f} us each, others {t_other / (frames - n_kf) * 1e6:7.1f} us each, hints {odo.promotions_prepared}", flush=True)
