"""Developer probe: where a frame of the device-resident odometry leg spends its wall time."""
import os, sys, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import delta_graph_slam_b200 as eng
from delta_graph_slam_b200 import synth, _lib
import bench

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rays = synth.num_rays(synth.HDL64)
d_raw = torch.empty((frames, rays, 4), dtype=torch.float32, device="cuda:0")
counts = [synth.scan_to_device(d_raw[k].data_ptr(), synth.traj_kitti_like(k), synth.HDL64, 1, 1000 + k, 0) for k in range(frames)]
clouds = [eng.DeviceCloud(d_raw[k].data_ptr(), counts[k], d_raw) for k in range(frames)]
d_ds = torch.empty((rays, 4), dtype=torch.float32, device="cuda:0")
ds_buf = eng.DeviceCloud(d_ds.data_ptr(), rays, d_ds)
pre = eng.Prefilter(bench.PREFILTER_PARAMS, out=bench.DEVNULL)
odo = eng.ScanMatchingOdometry(bench.ODOM_PARAMS, out=bench.DEVNULL)
reg = odo.registration
acc = collections.defaultdict(float)
cnt = collections.defaultdict(int)
def wrap(obj, name):
    f = getattr(obj, name)
    def g(*a, **k):
        t = time.perf_counter(); r = f(*a, **k); acc[name] += time.perf_counter() - t; cnt[name] += 1; return r
    setattr(obj, name, g)
for n in ("setInputTarget", "setInputSource", "align", "hasConverged", "getFinalTransformation", "promoteSourceToTarget"):
    wrap(reg, n)
wrap(pre, "downsample")
timing = len(sys.argv) > 2 and sys.argv[2] == "timing"
if timing:
    reg.setTiming(True)
for rep in range(2):
    acc.clear(); cnt.clear()
    odo.keyframe = None
    t0 = time.perf_counter()
    for k, c in enumerate(clouds):
        f = pre.downsample(c, out=ds_buf)
        odo.matching(0.1 * k, f)
    tot = time.perf_counter() - t0
print(f"{frames} frames: {tot / frames * 1e6:.1f} us / frame")
if timing:
    c = reg.counters()
    print(f"event-timed align kernel: {c['align_kernel_ms'] / c['timed_aligns'] * 1e3:.1f} us over {c['timed_aligns']} launches")
for k, v in sorted(acc.items(), key=lambda x: -x[1]):
    print(f"  {k:28s} calls {cnt[k]:5d}  {v / frames * 1e6:8.1f} us/frame  {v / cnt[k] * 1e6:8.1f} us/call")
print(f"  python + numpy outside the calls: {(tot - sum(acc.values())) / frames * 1e6:.1f} us/frame")
