"""Developer probe: device time of setInputTarget (NDT grid build) and of the VoxelGrid filter over SM budgets."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import delta_graph_slam_b200 as eng
from delta_graph_slam_b200 import synth
import bench

rays = synth.num_rays(synth.HDL64)
d_raw = torch.empty((rays, 4), dtype=torch.float32, device="cuda:0")
n = synth.scan_to_device(d_raw.data_ptr(), synth.traj_kitti_like(3), synth.HDL64, 1, 1003, 0)
raw = eng.DeviceCloud(d_raw.data_ptr(), n, d_raw)
d_ds = torch.empty((rays, 4), dtype=torch.float32, device="cuda:0")
vg = eng.VoxelGrid(); vg.setLeafSize(0.1, 0.1, 0.1)


def timed(stream_ptr, fn, reps=30):
    st = torch.cuda.ExternalStream(stream_ptr, device="cuda:0")
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); fn(); e1.record(st); e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts)), float(np.min(ts))


for budget in (148, 96, 64, 40, 32, 24, 16):
    vg.setSmBudget(budget)
    vg.setInputCloud(raw, is_dense=False)
    f = [None]
    def run():
        f[0] = vg.filter(out=eng.DeviceCloud(d_ds.data_ptr(), rays, d_ds))
    med, mn = timed(vg._reg.stream(), run)
    print(f"VoxelGrid 0.1 of {n} points, budget {budget:3d}: median {med:6.1f} us  min {mn:6.1f} us  -> {f[0].n} points", flush=True)
ds = f[0]
for budget in (148, 96, 64, 48, 32, 24, 16, 8):
    ndt = eng.select_registration_method(bench.ODOM_PARAMS, out=bench.DEVNULL)
    ndt.setSmBudget(budget)
    med, mn = timed(ndt.stream(), lambda: ndt.setInputTargetDevice(ds.ptr, ds.n))
    print(f"NDT setInputTarget of {ds.n} points, sort budget {budget:3d}: median {med:6.1f} us  min {mn:6.1f} us", flush=True)
