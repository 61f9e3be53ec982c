"""Developer probe: the pipelined front end (eng.FrontEnd) against the plain loop, over SM budgets.
usage: python tools/pipeline_probe.py [frames]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import delta_graph_slam_b200 as eng
from delta_graph_slam_b200 import synth
import bench

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 300
rays = synth.num_rays(synth.HDL64)
d_raw = torch.empty((frames, rays, 4), dtype=torch.float32, device="cuda:0")
counts = [synth.scan_to_device(d_raw[k].data_ptr(), synth.traj_kitti_like(k), synth.HDL64, 1, 1000 + k, 0) for k in range(frames)]
dev_clouds = [eng.DeviceCloud(d_raw[k].data_ptr(), counts[k], d_raw) for k in range(frames)]
h_raw = torch.empty((frames, rays, 4), dtype=torch.float32, pin_memory=True)
h_raw.copy_(d_raw)
torch.cuda.synchronize()
host_clouds = [h_raw.numpy()[k, : counts[k]] for k in range(frames)]
d_ds = torch.empty((3, rays, 4), dtype=torch.float32, device="cuda:0")
ds_bufs = [eng.DeviceCloud(d_ds[j].data_ptr(), rays, d_ds) for j in range(3)]
h_out = torch.empty((3, rays, 4), dtype=torch.float32, pin_memory=True).numpy()
h_bufs = [h_out[j] for j in range(3)]


def run(mode, clouds, bufs, filter_sms, reg_sms, reps=3):
    pre = eng.Prefilter(bench.PREFILTER_PARAMS, out=bench.DEVNULL)
    odo = eng.ScanMatchingOdometry(bench.ODOM_PARAMS, out=bench.DEVNULL)
    pre.filter.setSmBudget(filter_sms)
    odo.registration.setSmBudget(reg_sms)
    odo.registration.setTiming(True)
    fe = eng.FrontEnd(pre, odo, bufs, filter_sms=0)
    best = 1e9
    for rep in range(reps):
        odo.keyframe = None
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if mode == "pipe":
            poses = fe.run(clouds)
        else:
            poses = [odo.matching(0.1 * k, pre.downsample(c, out=bufs[k % 3])) for k, c in enumerate(clouds)]
        best = min(best, time.perf_counter() - t0)
    c = odo.registration.counters()
    return best / frames * 1e6, c["align_kernel_ms"] / max(c["timed_aligns"], 1) * 1e3, poses


for name, clouds, bufs in (("device", dev_clouds, ds_bufs), ("host", host_clouds, h_bufs)):
    base = None
    for mode, fs, rs in (("seq", 148, 148), ("seq", 148, 128), ("seq", 148, 108), ("seq", 148, 96), ("seq", 40, 108),
                         ("pipe", 148, 148), ("pipe", 24, 124), ("pipe", 32, 116), ("pipe", 40, 108), ("pipe", 52, 96), ("pipe", 64, 84)):
        us, kern, poses = run(mode, clouds, bufs, fs, rs)
        tag = ""
        if mode == "seq" and rs == 108 and fs == 40:
            base = poses
        if mode == "pipe" and rs == 108 and base is not None:
            tag = " poses==seq(40,108): %s" % all(np.array_equal(a, b) for a, b in zip(poses, base))
        print(f"{name:6s} {mode:4s} filter_sms {fs:3d} reg_sms {rs:3d}: {us:7.1f} us/frame ({1e6 / us:7.1f} reg/s)  align kernel {kern:6.1f} us{tag}", flush=True)
