"""Pins of the oracle's fast_gicp restatement, configured as
[REF src/hdl_graph_slam/registrations.cpp:27-36]: k-NN covariances with PLANE regularisation
against numpy, recovery of a known transform with LM and GN, the convergence contract."""
import numpy as np
import pytest


def rot_angle(Ra, Rb):
    R = np.asarray(Ra, np.float64).T @ np.asarray(Rb, np.float64)
    w = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / 2.0
    return float(np.arctan2(np.linalg.norm(w), (np.trace(R) - 1.0) / 2.0))


@pytest.mark.parametrize("reg_method", ["PLANE", "NONE", "MIN_EIG", "NORMALIZED_MIN_EIG", "FROBENIUS"])
def test_gicp_covariances_match_numpy(oracle, scans, reg_method):
    code = dict(NONE=0, MIN_EIG=1, NORMALIZED_MIN_EIG=2, PLANE=3, FROBENIUS=4)[reg_method]
    cloud = scans["ds1"][::5].copy()
    reg = oracle.Registration(oracle.GICP, regularization=code, k_corr=20)
    reg.setInputSource(cloud)
    C = reg.gicp_covariances(0, len(cloud))
    idx, _ = oracle.knn(cloud, cloud, 20)
    assert np.array_equal(idx[:, 0], np.arange(len(cloud))) or np.all(np.linalg.norm(cloud[idx[:, 0], :3] - cloud[:, :3], axis=1) == 0), "the query point is its own nearest neighbour"
    for i in range(0, len(cloud), 97):
        nb = cloud[idx[i], :3].astype(np.float64)
        d = nb - nb.mean(axis=0)
        cov = d.T @ d / 20.0
        w, V = np.linalg.eigh(cov)
        if reg_method == "NONE":
            want = cov
        elif reg_method == "PLANE":
            want = V @ np.diag([1e-3, 1.0, 1.0]) @ V.T
        elif reg_method == "MIN_EIG":
            want = V @ np.diag(np.maximum(w, 1e-3)) @ V.T
        elif reg_method == "NORMALIZED_MIN_EIG":
            want = V @ np.diag(np.maximum(w / w[2], 1e-3)) @ V.T
        else:
            Ci = np.linalg.inv(cov + 1e-3 * np.eye(3))
            want = np.linalg.inv(Ci / np.linalg.norm(Ci))
        gap = (w[1] - w[0]) / max(w[2], 1e-30)
        if reg_method in ("PLANE",) and gap < 1e-6:
            continue  # degenerate normal direction: any vector of the 2-D eigenspace is valid
        assert np.max(np.abs(C[i] - want)) < 1e-7 * max(1.0, np.abs(want).max()) / max(min(gap, 1.0), 1e-3)


@pytest.mark.parametrize("lsq", ["LM", "GN"])
def test_gicp_recovers_known_transform(oracle, scans, lsq):
    tgt = scans["ds0"][::2].copy()
    P = oracle.synth_pose(np.array([0.25, -0.10, 0.03, 0.005, -0.004, 0.015]))
    Pinv = np.linalg.inv(P)
    src = np.ones_like(tgt[::2])
    src[:, :3] = (tgt[::2, :3].astype(np.float64) @ Pinv[:3, :3].T + Pinv[:3, 3]).astype(np.float32)
    reg = oracle.Registration(oracle.GICP, trans_eps=1e-4, rot_eps=1e-5, max_iter=64, max_corr_dist=2.5, lsq=1 if lsq == "LM" else 0)
    reg.setInputTarget(tgt)
    reg.setInputSource(src)
    reg.align(None)
    T = reg.getFinalTransformation().astype(np.float64)
    assert reg.hasConverged()
    assert np.max(np.abs(T[:3, 3] - P[:3, 3])) < 0.005
    assert rot_angle(P[:3, :3], T[:3, :3]) < np.deg2rad(0.05)
    assert reg.getFitnessScore() < 1e-4


def test_gicp_consecutive_scans_and_contract(oracle, scans):
    """Frames 0 -> 1 of the synthetic sequence with the reference's parameters (epsilon 0.01,
    64 iterations, 2.5 m correspondences, k = 20): converges next to the ground truth; with
    max_iterations = 1 the loop stops un-converged; an impossible correspondence distance leaves
    the guess untouched."""
    reg = oracle.Registration(oracle.GICP, trans_eps=0.01, max_iter=64, max_corr_dist=2.5, k_corr=20)
    reg.setInputTarget(scans["ds0"])
    reg.setInputSource(scans["ds1"])
    reg.align(None)
    T = reg.getFinalTransformation()
    gt = scans["gt"]
    assert reg.hasConverged() and reg.getFinalNumIteration() <= 10
    assert np.max(np.abs(T[:3, 3] - gt[:3, 3])) < 0.02
    assert rot_angle(gt[:3, :3], T[:3, :3]) < np.deg2rad(0.1)
    one = oracle.Registration(oracle.GICP, trans_eps=1e-9, rot_eps=1e-12, max_iter=1, max_corr_dist=2.5)
    one.setInputTarget(scans["ds0"])
    one.setInputSource(scans["ds1"])
    one.align(None)
    assert not one.hasConverged() and one.getFinalNumIteration() == 0
    none = oracle.Registration(oracle.GICP, trans_eps=0.01, max_iter=8, max_corr_dist=1e-6)
    none.setInputTarget(scans["ds0"][::50])
    none.setInputSource(scans["ds1"][::50])
    guess = np.eye(4, dtype=np.float32)
    guess[0, 3] = 0.123
    none.align(guess)
    assert np.allclose(none.getFinalTransformation(), guess, atol=1e-6)
