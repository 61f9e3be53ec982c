"""Developer probe: cycle counters of k_gicp_align on the cfg-1 pair (b200reg_get_profile)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import delta_graph_slam_b200 as d
from oracle import oracle_py as O
import bench
P0, P1 = O.synth_traj(0), O.synth_traj(1)
v0 = O.voxelgrid(O.synth_scan(P0, noise_seed=1000), 0.1)["out"]
v1 = O.voxelgrid(O.synth_scan(P1, noise_seed=1001), 0.1)["out"]
g = d.select_registration_method(bench.GICP_ODOM_PARAMS, out=bench.DEVNULL)
g.setProfile(True)
g.setInputTarget(v0); g.setInputSource(v1)
for _ in range(3):
    g.setInputSource(v1)
    t = time.perf_counter(); g.align(None); dt = time.perf_counter() - t
v = np.zeros(16, np.int64)
d._lib.check(g._h, d._lib.load().b200reg_get_profile(g._h, v.ctypes.data))
names = ("near", "far_queue", "error_pass", "block_reduce", "group_barrier", "row_sum", "step", "n_linearize", "n_error", "far_queries_cta0", "warp0_brute_calls", "warp0_far_cycles", "warp0_brute_cycles", "fq_queue_fill", "fq_search", "fq_finish")
print("align (with covariances of the source) %.1f us" % (dt * 1e6), g.getResult()["iterations"], "iterations")
print({n: int(x) for n, x in zip(names, v[:16])})
print({n: round(float(x) / 1965.0, 1) for n, x in zip(names[:7], v[:7])}, "us total")
