"""Developer timing probe (wall clock around synchronous C-ABI calls) — not a bench."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import delta_graph_slam_b200 as d
from oracle import oracle_py as O

def t(f, n=20, warm=3):
    for _ in range(warm): f()
    ts = []
    for _ in range(n):
        a = time.perf_counter(); f(); ts.append(time.perf_counter() - a)
    ts = np.array(ts) * 1e6
    return f"{np.median(ts):9.1f} us (min {ts.min():.1f})"

P0, P1 = O.synth_traj(0), O.synth_traj(1)
s0, s1 = O.synth_scan(P0, noise_seed=1000), O.synth_scan(P1, noise_seed=1001)
vg = d.VoxelGrid(); vg.setLeafSize(0.1, 0.1, 0.1)
vg.setInputCloud(s0); v0 = vg.filter()
vg.setInputCloud(s1); v1 = vg.filter()
print("raw", len(s0), "ds", len(v0), len(v1))
print("voxelgrid_filter(host->host)", t(lambda: vg.filter()))
for search in ("DIRECT7", "DIRECT1", "KDTREE"):
    ndt = d.select_registration_method(dict(registration_method="NDT_OMP", reg_resolution=1.0, reg_nn_search_method=search), out=open(os.devnull, "w"))
    print(search, "set_target", t(lambda: ndt.setInputTarget(v0)))
    ndt.setInputSource(v1)
    ndt.setProfile(True)
    print(search, "set_source", t(lambda: ndt.setInputSource(v1)))
    print(search, "align(identity)", t(lambda: ndt.align(None)), ndt.getResult()["iterations"], "iters", ndt.getResult()["evaluations"], "evals", ndt.getResult()["hits"], "hits")
    print(search, "profile(cycles)", ndt.profile())
    print(search, "derivatives", t(lambda: ndt.ndt_derivatives(np.zeros(6))))
    print(search, "fitness", t(lambda: ndt.getFitnessScore()), ndt.getFitnessScore())
    print(search, "T err", np.abs(ndt.getFinalTransformation() - (np.linalg.inv(P0) @ P1)).max())
