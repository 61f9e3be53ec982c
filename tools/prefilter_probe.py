"""The whole prefilter nodelet per scan through HOST buffers (page-locked in / out): distance gate + VoxelGrid 0.1 m ->
statistical outlier filter (20, 1.0) -> filtered3D, then height / normal / flatten -> filtered2D, next to the OpenMP oracle.
python tools/prefilter_probe.py [n_scans]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import delta_graph_slam_b200 as eng  # noqa: E402
from oracle import oracle_py as O  # noqa: E402  (input generation, the CPU leg and the check only)

PARAMS = dict(downsample_method="VOXELGRID", downsample_resolution=0.1, use_distance_filter=True, distance_near_thresh=0.1, distance_far_thresh=100.0,
              outlier_removal_method="STATISTICAL", statistical_mean_k=20, statistical_stddev=1.0)


def main():
    n_scans = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    scans = [O.synth_scan(O.synth_traj(k), noise_seed=1000 + k) for k in range(n_scans)]
    cap = max(len(s) for s in scans)
    pin = lambda: torch.empty((cap, 4), dtype=torch.float32, pin_memory=True).numpy()
    h_in, h_ds, h_3d, h_2d = pin(), pin(), pin(), pin()
    pre = eng.Prefilter(PARAMS, out=open(os.devnull, "w"))
    ok = True
    ts = []
    for rep in range(3):
        for s in scans:
            h_in[: len(s)] = s
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            f3d = pre.filter3d(h_in[: len(s)], out=h_ds, out2=h_3d)
            f2d = pre.filter2d(f3d, 0.0, out=h_2d)
            ts.append((time.perf_counter() - t0) * 1e3)
            if rep == 0:
                w3 = O.statistical_outlier_removal(O.voxelgrid(O.distance_filter(s, 0.1, 100.0), 0.1, is_dense=False)["out"], 20, 1.0)
                w2 = O.flat_filter(w3, 0.0)
                ok = ok and np.array_equal(np.array(f3d).view(np.uint32), w3.view(np.uint32)) and np.array_equal(np.array(f2d).view(np.uint32), w2.view(np.uint32))
    t0 = time.perf_counter()
    for s in scans[:4]:
        O.flat_filter(O.statistical_outlier_removal(O.voxelgrid(O.distance_filter(s, 0.1, 100.0), 0.1, is_dense=False)["out"], 20, 1.0), 0.0)
    cpu_ms = (time.perf_counter() - t0) * 1e3 / min(4, n_scans)
    print(json.dumps(dict(scans=n_scans, points_per_scan=int(np.mean([len(s) for s in scans])), prefilter_ms_per_scan=float(np.median(ts[n_scans:])),
                          scans_per_s=1e3 / float(np.median(ts[n_scans:])), oracle_ms_per_scan=cpu_ms, oracle_cores=os.cpu_count(), bit_identical=bool(ok))))


if __name__ == "__main__":
    main()
