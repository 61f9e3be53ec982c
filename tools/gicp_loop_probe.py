"""Developer probe: LoopDetector.matching's serial loop on a FAST_GICP handle (the launch file's loop detector), 16 candidates
against one new keyframe at full HDL-64 size — clouds handed over per pair (setInputSource: upload + covariances every time)
against the keyframe cache (b200reg_set_source_cached: covariances once per keyframe).  Wall clock per pair, under gpurun."""
import io, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import delta_graph_slam_b200 as eng
from delta_graph_slam_b200 import synth
from oracle import oracle_py as O

n_c = 16
clouds = [O.voxelgrid(O.synth_scan(O.synth_traj(4 * k), noise_seed=2000 + k), 0.1)["out"] for k in range(n_c + 1)]
guess = [np.linalg.inv(O.synth_traj(0)) @ O.synth_traj(4 * k) for k in range(1, n_c + 1)]
guess = [g.astype(np.float32) for g in guess]
out = {}
for mode in ("plain", "cached"):
    reg = eng.select_registration_method(dict(registration_method="FAST_GICP"), out=io.StringIO())
    if mode == "cached":
        for k, c in enumerate(clouds):
            reg.cloudPut(k, c)
        reg.cloudSync()
    res = []
    for rep in range(3):  # rep 0 warms up (and, cached, computes every keyframe's covariances once)
        t0 = time.perf_counter()
        if mode == "cached":
            reg.setInputTargetCached(0)
        else:
            reg.setInputTarget(clouds[0])
        for k in range(1, n_c + 1):
            if mode == "cached":
                reg.setInputSourceCached(k)
            else:
                reg.setInputSource(clouds[k])
            reg.align(guess[k - 1])
            f = reg.getFitnessScore()
            if rep == 2:
                res.append((reg.getFinalTransformation().copy(), f, reg.hasConverged()))
        dt = time.perf_counter() - t0
    out[mode] = (1e6 * dt / n_c, res)
same = all(np.array_equal(a[0], b[0]) and a[1] == b[1] and a[2] == b[2] for a, b in zip(out["plain"][1], out["cached"][1]))
print(f"points per cloud ~{int(np.mean([len(c) for c in clouds]))}; us per pair (setInputSource + align + getFitnessScore): plain {out['plain'][0]:.0f}, cached {out['cached'][0]:.0f}; results bit-identical: {same}; converged {sum(r[2] for r in out['cached'][1])} of {n_c}")
