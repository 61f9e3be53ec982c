"""Developer probe: three FAST_GICP aligns of the cfg-1 pair (profiling off) — the target of an ncu capture of k_gicp_align."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import delta_graph_slam_b200 as d
from oracle import oracle_py as O
import bench
P0, P1 = O.synth_traj(0), O.synth_traj(1)
v0 = O.voxelgrid(O.synth_scan(P0, noise_seed=1000), 0.1)["out"]
v1 = O.voxelgrid(O.synth_scan(P1, noise_seed=1001), 0.1)["out"]
g = d.select_registration_method(bench.GICP_ODOM_PARAMS, out=bench.DEVNULL)
g.setInputTarget(v0)
for _ in range(3):
    g.setInputSource(v1)
    g.align(None)
print(g.getResult()["iterations"], "iterations")
