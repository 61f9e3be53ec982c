"""Developer probe: a short device-resident odometry sequence through the native front end, for launch lists
(ncu --metrics gpu__time_duration.sum) and quick timing.  python tools/seq_probe.py [NDT|GICP] [frames] [repeats]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
import delta_graph_slam_b200 as eng  # noqa: E402
from delta_graph_slam_b200 import synth  # noqa: E402

method = sys.argv[1] if len(sys.argv) > 1 else "NDT"
F = int(sys.argv[2]) if len(sys.argv) > 2 else 30
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
rays = synth.num_rays(synth.HDL64)
d_raw = torch.empty((F, rays, 4), dtype=torch.float32, device="cuda:0")
counts = [synth.scan_to_device(d_raw[k].data_ptr(), synth.traj_kitti_like(k), synth.HDL64, scene_seed=1, noise_seed=1000 + k, device=0) for k in range(F)]
clouds = [eng.DeviceCloud(d_raw[k].data_ptr(), counts[k], d_raw) for k in range(F)]
params = {**(bench.GICP_ODOM_PARAMS if method == "GICP" else bench.ODOM_PARAMS), **bench.PREFILTER_PARAMS}
fe = eng.NativeFrontEnd(params, filter_sms=40, prepare_promotion=2 if method == "NDT" else 0)
fe.run_device(clouds)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(reps):
    poses, res, nf = fe.run_device(clouds)
dt = (time.perf_counter() - t0) / reps
print(f"{method}: {F} frames, {1e6 * dt / F:.1f} us per frame, {(F - 1) / dt:.0f} registrations/s, keyframes {fe.num_keyframes()}, host {fe.timing()}")
