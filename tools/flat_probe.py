"""One flat-cloud filter call (height -> normal -> flatten) on a down-sampled synthetic HDL-64 scan, device-resident, for ncu:
the profiled region is one call after warm-up (torch.cuda.profiler start / stop; run ncu with --profile-from-start off).
Prints the median wall time of the synchronous call.  python tools/flat_probe.py"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import delta_graph_slam_b200 as eng  # noqa: E402
from oracle import oracle_py as O  # noqa: E402  (input generation and the check only)


def main():
    raw = O.synth_scan(O.synth_traj(3), noise_seed=1003)
    ds = O.voxelgrid(O.distance_filter(raw, 0.1, 100.0), 0.1, is_dense=False)["out"]
    d_in = torch.from_numpy(ds).cuda()
    d_out = torch.zeros_like(d_in)
    cin, cout = eng.DeviceCloud(d_in.data_ptr(), len(ds), d_in), eng.DeviceCloud(d_out.data_ptr(), len(ds), d_out)
    reg = eng.Registration()
    ts = []
    for _ in range(int(os.environ.get("FLAT_PROBE_REPS", "30"))):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = reg.flat_filter(cin, 0.0, out=cout)
        ts.append((time.perf_counter() - t0) * 1e3)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    res = reg.flat_filter(cin, 0.0, out=cout)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    want = O.flat_filter(ds, 0.0)
    got = d_out[: res.n].cpu().numpy()
    print(json.dumps(dict(points=len(ds), above_lidar=int((ds[:, 2] > 0).sum()), kept=res.n, flat_filter_ms=float(np.median(ts[3:])),
                          bit_identical=bool(len(got) == len(want) and np.array_equal(got.view(np.uint32), want.view(np.uint32))))))


if __name__ == "__main__":
    main()
