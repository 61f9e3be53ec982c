"""Developer probe: pcl::VoxelGrid (0.1 m) of one dense128 (1 M-ray) and one HDL-64 scan through each sort path
(b200reg_set_sort_path), device-resident, CUDA events around the call.  Under gpurun; with ncu --metrics
gpu__time_duration.sum it gives the per-kernel launch list of the filter."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import delta_graph_slam_b200 as eng  # noqa: E402
from delta_graph_slam_b200 import _lib, synth  # noqa: E402

L = _lib.load()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
paths = [int(c) for c in sys.argv[2]] if len(sys.argv) > 2 else [0, 3, 2, 1]     # e.g. "2": the onesweep path only
sensors = sys.argv[3].split(",") if len(sys.argv) > 3 else ["dense128", "hdl64"]
out = {}
for name, sensor in (("dense128", synth.DENSE128), ("hdl64", synth.HDL64)):
    if name not in sensors:
        continue
    rays = synth.num_rays(sensor)
    d_raw = torch.empty((rays, 4), dtype=torch.float32, device="cuda:0")
    d_out = torch.empty((rays, 4), dtype=torch.float32, device="cuda:0")
    n = synth.scan_to_device(d_raw.data_ptr(), synth.traj_kitti_like(3), sensor, scene_seed=1, noise_seed=77, device=0)
    ref = None
    for path in paths:
        L.b200reg_set_sort_path(path)
        vg = eng.VoxelGrid()
        vg.setLeafSize(0.1, 0.1, 0.1)
        vg.setInputCloud(eng.DeviceCloud(d_raw.data_ptr(), n, d_raw), is_dense=False)
        ob = eng.DeviceCloud(d_out.data_ptr(), rays, d_out)
        st = torch.cuda.ExternalStream(vg._reg.stream(), device="cuda:0")
        for _ in range(3):
            f = vg.filter(out=ob)
        ms = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            f = vg.filter(out=ob)
            e1.record(st)
            e1.synchronize()
            ms.append(e0.elapsed_time(e1))
        res = d_out[: f.n].cpu().numpy().copy()
        if ref is None:
            ref = res
        out[f"{name}_path{path}"] = dict(points=n, voxels=f.n, us=1e3 * float(np.median(ms)), us_min=1e3 * float(np.min(ms)), same=bool(np.array_equal(res.view(np.uint32), ref.view(np.uint32))),
                                         gbs=(16 * n + 16 * f.n) / (1e-3 * float(np.median(ms))) / 1e9)
    L.b200reg_set_sort_path(0)
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/vg_probe.json", "w") as fh:
    json.dump(out, fh, indent=1)
