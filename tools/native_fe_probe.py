"""Developer probe: the native front end (b200reg_frontend_run_device) over a device-resident synthetic sequence under
different scheduling options — prepared promotions off / predicted / always, SM split.  Poses must not change.  Under gpurun.
  python tools/native_fe_probe.py [frames]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
import delta_graph_slam_b200 as eng  # noqa: E402
from delta_graph_slam_b200 import synth  # noqa: E402

F = int(sys.argv[1]) if len(sys.argv) > 1 else 300
rays = synth.num_rays(synth.HDL64)
d_raw = torch.empty((F, rays, 4), dtype=torch.float32, device="cuda:0")
counts = [synth.scan_to_device(d_raw[k].data_ptr(), synth.traj_kitti_like(k), synth.HDL64, scene_seed=1, noise_seed=1000 + k, device=0) for k in range(F)]
clouds = [eng.DeviceCloud(d_raw[k].data_ptr(), counts[k], d_raw) for k in range(F)]
params = {**bench.ODOM_PARAMS, **bench.PREFILTER_PARAMS}
base = None
out = []
for (prep, fsm, side) in ((0, 40, 16), (1, 40, 16), (2, 40, 16), (1, 48, 24), (2, 48, 24), (0, 32, 16), (2, 56, 32)):
    fe = eng.NativeFrontEnd(params, filter_sms=fsm, prepare_promotion=prep, side_sms=side)
    fe.set_timing(True)
    for _ in range(2):
        poses, res, nf = fe.run_device(clouds)
    fe.timing()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        poses, res, nf = fe.run_device(clouds)
    dt = (time.perf_counter() - t0) / 3
    tm = fe.timing()
    c = fe.counters()
    if base is None:
        base = poses
    out.append(dict(prepare=prep, filter_sms=fsm, side_sms=side, us_per_frame=1e6 * dt / F, regs_per_s=(F - 1) / dt, same_poses=bool(np.array_equal(poses, base)),
                    align_kernel_us=1e3 * c["align_kernel_ms"] / max(c["timed_aligns"], 1), host={k: round(v, 1) if isinstance(v, float) else v for k, v in tm.items()}))
    print(json.dumps(out[-1]), flush=True)
    del fe
