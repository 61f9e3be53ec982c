"""Small deterministic driver for ncu: a few frames of the bench's device-resident odometry leg
(VoxelGrid 0.1 + NDT DIRECT7 keyframe odometry).  Not a bench: numbers printed here are not reported."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import delta_graph_slam_b200 as eng
from delta_graph_slam_b200 import synth
import bench

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rays = synth.num_rays(synth.HDL64)
d_raw = torch.empty((frames, rays, 4), dtype=torch.float32, device="cuda:0")
counts = [synth.scan_to_device(d_raw[k].data_ptr(), synth.traj_kitti_like(k), synth.HDL64, 1, 1000 + k, 0) for k in range(frames)]
clouds = [eng.DeviceCloud(d_raw[k].data_ptr(), counts[k], d_raw) for k in range(frames)]
d_ds = torch.empty((rays, 4), dtype=torch.float32, device="cuda:0")
ds_buf = eng.DeviceCloud(d_ds.data_ptr(), rays, d_ds)
pre = eng.Prefilter(bench.PREFILTER_PARAMS, out=bench.DEVNULL)
odo = eng.ScanMatchingOdometry(bench.GICP_ODOM_PARAMS if len(sys.argv) > 2 and sys.argv[2] == "gicp" else bench.ODOM_PARAMS, out=bench.DEVNULL)
poses = bench.run_sequence(pre, odo, clouds, out_buf=ds_buf)
if not (len(sys.argv) > 2 and sys.argv[2] == "gicp"):
    odo.registration.getFitnessScore()
print("frames", frames, "keyframes", odo.num_keyframes, "last pose t", poses[-1][:3, 3])
