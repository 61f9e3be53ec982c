import os, sys, io, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import delta_graph_slam_b200 as d
from oracle import oracle_py as O
P0, P1 = O.synth_traj(0), O.synth_traj(1)
s0, s1 = O.synth_scan(P0, noise_seed=1000), O.synth_scan(P1, noise_seed=1001)
vg = d.VoxelGrid(); vg.setLeafSize(0.1, 0.1, 0.1)
vg.setInputCloud(s0); v0 = vg.filter()
vg.setInputCloud(s1); v1 = vg.filter()
ndt = d.select_registration_method(dict(registration_method="NDT_OMP", reg_resolution=1.0), out=io.StringIO())
ndt.setInputTarget(v0); ndt.setInputSource(v1)
ndt.setTiming(True)
for mode in ("back-to-back", "with filter between", "with sleep between"):
    c0 = ndt.counters()
    t = time.perf_counter()
    for _ in range(50):
        ndt.align(None)
        if mode == "with filter between":
            vg.filter()
        elif mode == "with sleep between":
            time.sleep(0.0005)
    wall = (time.perf_counter() - t) / 50 * 1e6
    c1 = ndt.counters()
    n = c1["timed_aligns"] - c0["timed_aligns"]
    print(f"{mode:22s}: wall/iter {wall:7.1f} us   event-timed kernel {(c1['align_kernel_ms'] - c0['align_kernel_ms']) / n * 1e3:7.1f} us over {n} launches")
