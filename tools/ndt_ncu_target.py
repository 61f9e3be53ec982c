"""Developer probe: five NDT DIRECT7 aligns of the cfg-1 pair at the front end's registration budget (96 SMs, profiling
off) — the target of an ncu capture of k_ndt_align<7, false>."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import delta_graph_slam_b200 as d
from oracle import oracle_py as O
P0, P1 = O.synth_traj(0), O.synth_traj(1)
v0 = O.voxelgrid(O.synth_scan(P0, noise_seed=1000), 0.1)["out"]
v1 = O.voxelgrid(O.synth_scan(P1, noise_seed=1001), 0.1)["out"]
ndt = d.select_registration_method(dict(registration_method="NDT_OMP", reg_resolution=1.0, reg_nn_search_method="DIRECT7"), out=open(os.devnull, "w"))
ndt.setSmBudget(int(sys.argv[1]) if len(sys.argv) > 1 else 96)
ndt.setInputTarget(v0)
ndt.setInputSource(v1)
for _ in range(5):
    ndt.align(None)
r = ndt.getResult()
print(r["iterations"], "iterations", r["passes"], "passes")
