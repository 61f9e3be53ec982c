"""Generate tests/golden/*.npz with the CPU oracle at fixed seeds (SURVEY.md §8c pin vi).

The reference holds no golden vectors and its numerics cannot run here, so these
fixtures are produced by the oracle restatement; they freeze its behaviour so later refactors of the
oracle or the engine cannot drift silently.  Inputs are stored with the outputs (the synthetic
generator depends on libm).  Re-run only on purpose:  python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle_py as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
DBL_MAX = np.finfo(np.float64).max


def voxelgrid_case():
    rng = np.random.default_rng(2024)
    pts = np.ones((6000, 4), np.float32)
    pts[:, :3] = rng.normal(0, 4, (6000, 3)).astype(np.float32)
    pts[::13, 1] = np.nan
    pts[7::101, 2] = -np.inf
    pts[200:230] = pts[199]
    r = O.voxelgrid(pts, (0.3, 0.2, 0.5), min_points_per_voxel=0, is_dense=False)
    r2 = O.voxelgrid(pts, 0.25, min_points_per_voxel=2, is_dense=False)
    np.savez_compressed(os.path.join(OUT, "voxelgrid.npz"), pts=pts, leaf=np.array([0.3, 0.2, 0.5], np.float32), out=r["out"], voxel_id=r["voxel_id"], count=r["count"], key=r["key"],
                        min_b=r["min_b"], div_b=r["div_b"], out_min2=r2["out"], count_min2=r2["count"])


def prefilter_case():
    """distance_filter -> VoxelGrid -> RadiusOutlierRemoval with the reference launch file's parameters
    [REF launch/delta_graph_slam.launch:31-42] on a quarter of a synthetic scan plus stray / non-finite points."""
    rng = np.random.default_rng(77)
    raw = O.synth_scan(O.synth_traj(4), noise_seed=1004)[::4].copy()
    stray = np.ones((60, 4), np.float32)
    stray[:, :3] = rng.uniform(-160, 160, (60, 3)).astype(np.float32)
    bad = np.ones((3, 4), np.float32)
    bad[0, 0], bad[1, 1], bad[2, 2] = np.nan, np.inf, -np.inf
    pts = np.concatenate([raw[:500], stray[:30], bad, raw[500:], stray[30:]]).astype(np.float32)
    gated = O.distance_filter(pts, 0.1, 100.0)
    ds = O.voxelgrid(gated, 0.1, is_dense=False)["out"]
    kept = O.radius_outlier_removal(ds, 0.5, 2)
    # the nodelet's own default outlier filter: STATISTICAL 20 - 1.0 [REF apps/prefiltering_nodelet.cpp:77-80]
    kept_stat, det = O.statistical_outlier_removal(ds, 20, 1.0, details=True)
    np.savez_compressed(os.path.join(OUT, "prefilter.npz"), pts=pts, near_far=np.array([0.1, 100.0]), leaf=np.float32(0.1), radius_min=np.array([0.5, 2.0]),
                        n_gated=np.int64(len(gated)), ds=ds, kept=kept, meank_mul=np.array([20.0, 1.0]), kept_stat=kept_stat,
                        stat=np.array([det["mean"], det["stddev"], det["threshold"]]), stat_dist=det["distances"])


def clouds():
    P0, P1 = O.synth_traj(0), O.synth_traj(1)
    s0 = O.synth_scan(P0, noise_seed=1000)[::4]
    s1 = O.synth_scan(P1, noise_seed=1001)[::4]
    return O.voxelgrid(s0, 0.4)["out"], O.voxelgrid(s1, 0.4)["out"], np.linalg.inv(P0) @ P1


def ndt_case(tgt, src):
    guess = np.eye(4, dtype=np.float32)
    guess[:3, 3] = [0.3, 0.05, 0.0]
    out = dict(tgt=tgt, src=src, guess=guess)
    for name, code in (("direct7", O.DIRECT7), ("direct1", O.DIRECT1), ("kdtree", O.KDTREE)):
        reg = O.Registration(O.NDT, resolution=1.0, nn_search=code, trans_eps=0.01, max_iter=64)
        reg.setInputTarget(tgt)
        reg.setInputSource(src)
        reg.align(guess)
        info = reg.info()
        out[f"{name}_T"] = reg.getFinalTransformation()
        out[f"{name}_meta"] = np.array([reg.hasConverged(), reg.getFinalNumIteration(), info[1], info[2]], np.int64)
        out[f"{name}_score"] = np.array([info[0], reg.getFitnessScore(), reg.getFitnessScore(1.0)])
        p = np.array([0.4, -0.1, 0.03, 0.01, -0.02, 0.05])
        s, g, H = reg.ndt_derivatives(p)
        out[f"{name}_deriv"] = np.concatenate([[s], g, H.ravel()])
    reg = O.Registration(O.NDT, resolution=1.0)
    reg.setInputTarget(tgt)
    L = reg.ndt_leaves()
    out.update(leaf_idx=L["idx"], leaf_n=L["n"], leaf_mean=L["mean"], leaf_icov=L["icov"])
    np.savez_compressed(os.path.join(OUT, "ndt.npz"), **out)


def gicp_case(tgt, src):
    guess = np.eye(4, dtype=np.float32)
    guess[:3, 3] = [0.3, 0.05, 0.0]
    out = dict(guess=guess)
    for name, lsq in (("lm", 1), ("gn", 0)):
        reg = O.Registration(O.GICP, trans_eps=0.01, max_iter=64, max_corr_dist=2.5, k_corr=20, lsq=lsq)
        reg.setInputTarget(tgt)
        reg.setInputSource(src)
        reg.align(guess)
        out[f"{name}_T"] = reg.getFinalTransformation()
        out[f"{name}_meta"] = np.array([reg.hasConverged(), reg.getFinalNumIteration()], np.int64)
        out[f"{name}_fitness"] = np.array([reg.getFitnessScore()])
        if lsq == 1:
            out["cov_src"] = reg.gicp_covariances(0, len(src))[::50]
            out["cov_tgt"] = reg.gicp_covariances(1, len(tgt))[::50]
    np.savez_compressed(os.path.join(OUT, "gicp.npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    voxelgrid_case()
    prefilter_case()
    tgt, src, gt = clouds()
    ndt_case(tgt, src)
    gicp_case(tgt, src)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
