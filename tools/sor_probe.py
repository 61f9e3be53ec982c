"""Time the two outlier filters on one down-sampled HDL-64 scan (device-resident in / out, wall clock around the
synchronous call, median of 50) next to the OpenMP oracle on the host cores.  python tools/sor_probe.py"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import delta_graph_slam_b200 as eng  # noqa: E402
from oracle import oracle_py as O  # noqa: E402  (checker only)


def med(fn, reps=int(os.environ.get("SOR_PROBE_REPS", "50")), warm=2):
    for _ in range(warm):
        fn()
    t = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        t.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(t))


def main():
    raw = O.synth_scan(O.synth_traj(3), noise_seed=1003)
    ds = O.voxelgrid(O.distance_filter(raw, 0.1, 100.0), 0.1, is_dense=False)["out"]
    d_in = torch.from_numpy(ds).cuda()
    d_out = torch.zeros_like(d_in)
    cin, cout = eng.DeviceCloud(d_in.data_ptr(), len(ds), d_in), eng.DeviceCloud(d_out.data_ptr(), len(ds), d_out)
    sor = eng.StatisticalOutlierRemoval()
    sor.setMeanK(20)
    sor.setStddevMulThresh(1.0)
    sor.setInputCloud(cin)
    ror = eng.RadiusOutlierRemoval(registration=sor._reg)
    ror.setRadiusSearch(0.5)
    ror.setMinNeighborsInRadius(2)
    ror.setInputCloud(cin)
    res = dict(points=len(ds))
    res["statistical_ms"] = med(lambda: sor.filter(out=cout))
    st = sor.last_stats()
    res["statistical_kept"] = sor.filter(out=cout).n
    res["exact_pass"] = st["exact_pass"]
    got = d_out[: res["statistical_kept"]].cpu().numpy()
    res["radius_ms"] = med(lambda: ror.filter(out=cout))
    t0 = time.perf_counter()
    want = O.statistical_outlier_removal(ds, 20, 1.0)
    res["oracle_statistical_ms"] = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    O.radius_outlier_removal(ds, 0.5, 2)
    res["oracle_radius_ms"] = (time.perf_counter() - t0) * 1e3
    res["oracle_cores"] = os.cpu_count()
    res["bit_identical"] = bool(len(got) == len(want) and np.array_equal(got.view(np.uint32), want.view(np.uint32)))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
