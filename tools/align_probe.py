"""Developer probe of k_ndt_align on the cfg-1 pair (frames 0 / 1, 0.1 m VoxelGrid, NDT DIRECT7, identity guess):
CUDA-event duration of the production kernel (b200reg_set_timing) at SM budgets 148 and 108, then the per-phase cycle
counters of the profiled instantiation.  B200REG_LIB selects an alternative build of the library.  Under gpurun."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import delta_graph_slam_b200 as d  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

P0, P1 = O.synth_traj(0), O.synth_traj(1)
v0 = O.voxelgrid(O.synth_scan(P0, noise_seed=1000), 0.1)["out"]
v1 = O.voxelgrid(O.synth_scan(P1, noise_seed=1001), 0.1)["out"]
out = {"lib": os.environ.get("B200REG_LIB", "default"), "points": [len(v0), len(v1)]}
for budget in (148, 108):
    ndt = d.select_registration_method(dict(registration_method="NDT_OMP", reg_resolution=1.0, reg_nn_search_method="DIRECT7"), out=open(os.devnull, "w"))
    ndt.setSmBudget(budget)
    ndt.setInputTarget(v0)
    ndt.setInputSource(v1)
    for _ in range(20):
        ndt.align(None)
    ndt.setTiming(True)
    for _ in range(200):
        ndt.align(None)
    c = ndt.counters()
    r = ndt.getResult()
    out[f"sm{budget}"] = {"us_per_registration": 1e3 * c["align_kernel_ms"] / c["timed_aligns"], "iterations": r["iterations"], "evaluations": r["evaluations"], "passes": r["passes"],
                          "us_per_pass": 1e3 * c["align_kernel_ms"] / c["timed_aligns"] / r["passes"], "T03": float(r["transformation"][0, 3])}
    ndt.setTiming(False)
    ndt.setProfile(True)
    ndt.align(None)
    p = ndt.profile()
    n = max(p["n"], 1)
    out[f"sm{budget}"]["cycles_per_pass"] = {k: round(v / n) for k, v in p.items() if k not in ("n", "stage")}
    out[f"sm{budget}"]["stage_cycles"] = p["stage"]
print(json.dumps(out))
