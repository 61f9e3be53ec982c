"""Golden vectors (tests/golden/*.npz, made by tools/make_golden.py with the oracle at fixed seeds).
CPU: the oracle must still reproduce them (guards the checker against drift).  GPU: the engine
through the C ABI against the same vectors — bit-exact for voxel work, 1e-4 m / 1e-4 rad / 1e-5
relative fitness for registrations."""
import os

import numpy as np
import pytest

from helpers import rot_angle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLD, name))


def close_T(Ta, Tb, tol_t=1e-4, tol_r=1e-4):
    return np.max(np.abs(Ta[:3, 3] - Tb[:3, 3])) < tol_t and rot_angle(Ta[:3, :3], Tb[:3, :3]) < tol_r


# ---------------------------------------------------------------- CPU: oracle vs golden
def test_oracle_voxelgrid_golden(oracle):
    g = load("voxelgrid.npz")
    r = oracle.voxelgrid(g["pts"], tuple(g["leaf"]), is_dense=False)
    for k in ("voxel_id", "count", "key", "min_b", "div_b"):
        assert np.array_equal(r[k], g[k]), k
    assert np.array_equal(r["out"].view(np.uint32), g["out"].view(np.uint32))
    r2 = oracle.voxelgrid(g["pts"], 0.25, min_points_per_voxel=2, is_dense=False)
    assert np.array_equal(r2["out"].view(np.uint32), g["out_min2"].view(np.uint32)) and np.array_equal(r2["count"], g["count_min2"])


def test_oracle_prefilter_golden(oracle):
    g = load("prefilter.npz")
    gated = oracle.distance_filter(g["pts"], *[float(x) for x in g["near_far"]])
    assert len(gated) == int(g["n_gated"])
    ds = oracle.voxelgrid(gated, float(g["leaf"]), is_dense=False)["out"]
    assert np.array_equal(ds.view(np.uint32), g["ds"].view(np.uint32))
    kept = oracle.radius_outlier_removal(ds, float(g["radius_min"][0]), int(g["radius_min"][1]))
    assert np.array_equal(kept.view(np.uint32), g["kept"].view(np.uint32))
    kept, det = oracle.statistical_outlier_removal(ds, int(g["meank_mul"][0]), float(g["meank_mul"][1]), details=True)
    assert np.array_equal(kept.view(np.uint32), g["kept_stat"].view(np.uint32)) and np.array_equal(det["distances"].view(np.uint32), g["stat_dist"].view(np.uint32))
    assert np.array_equal(np.array([det["mean"], det["stddev"], det["threshold"]]), g["stat"])


@pytest.mark.parametrize("name,code", [("direct7", 2), ("direct1", 3), ("kdtree", 0)])
def test_oracle_ndt_golden(oracle, name, code):
    g = load("ndt.npz")
    reg = oracle.Registration(oracle.NDT, resolution=1.0, nn_search=code, trans_eps=0.01, max_iter=64)
    reg.setInputTarget(g["tgt"])
    reg.setInputSource(g["src"])
    reg.align(g["guess"])
    meta = g[f"{name}_meta"]
    assert [int(reg.hasConverged()), reg.getFinalNumIteration(), int(reg.info()[1])] == meta[:3].tolist()
    assert close_T(reg.getFinalTransformation(), g[f"{name}_T"], 1e-6, 1e-6)  # OpenMP summation order may move the last bits
    sc = g[f"{name}_score"]
    assert abs(reg.getFitnessScore() - sc[1]) <= 1e-9 * sc[1] and abs(reg.getFitnessScore(1.0) - sc[2]) <= 1e-9 * sc[2]
    s, gr, H = reg.ndt_derivatives(np.array([0.4, -0.1, 0.03, 0.01, -0.02, 0.05]))
    want = g[f"{name}_deriv"]
    assert np.allclose(np.concatenate([[s], gr, H.ravel()]), want, rtol=1e-9, atol=1e-9 * np.abs(want).max())


@pytest.mark.parametrize("name,lsq", [("lm", 1), ("gn", 0)])
def test_oracle_gicp_golden(oracle, name, lsq):
    g, n = load("gicp.npz"), load("ndt.npz")
    reg = oracle.Registration(oracle.GICP, trans_eps=0.01, max_iter=64, max_corr_dist=2.5, k_corr=20, lsq=lsq)
    reg.setInputTarget(n["tgt"])
    reg.setInputSource(n["src"])
    reg.align(g["guess"])
    assert [int(reg.hasConverged()), reg.getFinalNumIteration()] == g[f"{name}_meta"].tolist()
    assert close_T(reg.getFinalTransformation(), g[f"{name}_T"], 1e-6, 1e-6)
    if lsq == 1:
        assert np.allclose(reg.gicp_covariances(0, len(n["src"]))[::50], g["cov_src"], atol=1e-12)


# ---------------------------------------------------------------- GPU: engine vs golden
@pytest.mark.gpu
def test_engine_voxelgrid_golden():
    import delta_graph_slam_b200 as eng
    g = load("voxelgrid.npz")
    vg = eng.VoxelGrid()
    vg.setLeafSize(*[float(x) for x in g["leaf"]])
    vg.setInputCloud(g["pts"], is_dense=False)
    out = vg.filter()
    lay = vg.last_layout(len(out), len(g["pts"]))
    assert np.array_equal(out.view(np.uint32), g["out"].view(np.uint32))
    for k in ("voxel_id", "count", "key", "min_b", "div_b"):
        assert np.array_equal(lay[k], g[k]), k
    vg.setLeafSize(0.25, 0.25, 0.25)
    vg.setMinimumPointsNumberPerVoxel(2)
    vg.setInputCloud(g["pts"], is_dense=False)
    assert np.array_equal(vg.filter().view(np.uint32), g["out_min2"].view(np.uint32))


@pytest.mark.gpu
def test_engine_prefilter_golden():
    import delta_graph_slam_b200 as eng
    g = load("prefilter.npz")
    pre = eng.Prefilter(dict(downsample_method="VOXELGRID", downsample_resolution=float(g["leaf"]), use_distance_filter=True, distance_near_thresh=float(g["near_far"][0]),
                             distance_far_thresh=float(g["near_far"][1]), outlier_removal_method="RADIUS", radius_radius=float(g["radius_min"][0]),
                             radius_min_neighbors=int(g["radius_min"][1])), out=open(os.devnull, "w"))
    ds = pre.downsample(g["pts"])
    assert np.array_equal(np.asarray(ds).view(np.uint32), g["ds"].view(np.uint32))
    kept = pre.outlier_removal(ds)
    assert np.array_equal(np.asarray(kept).view(np.uint32), g["kept"].view(np.uint32))
    sor = eng.StatisticalOutlierRemoval()
    sor.setMeanK(int(g["meank_mul"][0]))
    sor.setStddevMulThresh(float(g["meank_mul"][1]))
    sor.setInputCloud(g["ds"])
    assert np.array_equal(sor.filter().view(np.uint32), g["kept_stat"].view(np.uint32))
    st = sor.last_stats(len(g["ds"]))
    assert np.array_equal(st["distances"].view(np.uint32), g["stat_dist"].view(np.uint32)) and abs(st["threshold"] - g["stat"][2]) <= 1e-11 * g["stat"][2]


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["direct7", "direct1", "kdtree"])
def test_engine_ndt_golden(name):
    import io
    import delta_graph_slam_b200 as eng
    g = load("ndt.npz")
    ndt = eng.select_registration_method(dict(registration_method="NDT_OMP", reg_resolution=1.0, reg_nn_search_method=name.upper()), out=io.StringIO())
    ndt.setInputTarget(g["tgt"])
    ndt.setInputSource(g["src"])
    ndt.align(g["guess"])
    meta = g[f"{name}_meta"]
    res = ndt.getResult()
    assert [int(res["converged"]), res["iterations"], res["evaluations"]] == meta[:3].tolist()
    assert close_T(ndt.getFinalTransformation(), g[f"{name}_T"])
    sc = g[f"{name}_score"]
    assert abs(res["score"] - sc[0]) <= 1e-5 * abs(sc[0])
    assert abs(ndt.getFitnessScore() - sc[1]) <= 1e-5 * sc[1] and abs(ndt.getFitnessScore(1.0) - sc[2]) <= 1e-5 * sc[2]
    L = ndt.ndt_leaves()
    assert np.array_equal(L["idx"], g["leaf_idx"]) and np.array_equal(L["n"], g["leaf_n"])
    assert np.allclose(L["mean"], g["leaf_mean"], rtol=0, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("name,lsq", [("lm", 1), ("gn", 0)])
def test_engine_gicp_golden(name, lsq):
    import io
    import delta_graph_slam_b200 as eng
    g, n = load("gicp.npz"), load("ndt.npz")
    reg = eng.select_registration_method(dict(registration_method="FAST_GICP", reg_transformation_epsilon=0.01, reg_maximum_iterations=64, reg_max_correspondence_distance=2.5,
                                              reg_correspondence_randomness=20), out=io.StringIO())
    reg.setOptions(lsq_optimizer=lsq)
    reg.setInputTarget(n["tgt"])
    reg.setInputSource(n["src"])
    reg.align(g["guess"])
    assert [int(reg.hasConverged()), reg.getFinalNumIteration()] == g[f"{name}_meta"].tolist()
    assert close_T(reg.getFinalTransformation(), g[f"{name}_T"])
    assert abs(reg.getFitnessScore() - g[f"{name}_fitness"][0]) <= 1e-5 * g[f"{name}_fitness"][0]
    if lsq == 1:
        assert np.allclose(reg.covariances(0, len(n["src"]))[::50], g["cov_src"], atol=1e-6)
