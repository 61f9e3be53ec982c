"""Pins of the oracle's ndt_omp restatement (SURVEY.md §8c pins i, ii, v; Appendix A.3 / A.4 / A.6):
voxel statistics against numpy, analytic gradient / Hessian against finite differences of the
score, Eigen's eulerAngles(0,1,2) factorisation, and recovery of a known transform."""
import numpy as np
import pytest


def test_ndt_voxel_statistics_match_numpy(oracle, scans):
    reg = oracle.Registration(oracle.NDT, resolution=1.0)
    tgt = scans["ds0"]
    reg.setInputTarget(tgt)
    L = reg.ndt_leaves()
    # independent assignment: floor(x * (1 / leaf)) in float32, linear index over the bounding lattice
    ijk = np.floor(tgt[:, :3] * np.float32(1.0)).astype(np.int64)
    mn, mx = ijk.min(axis=0), ijk.max(axis=0)
    div = mx - mn + 1
    assert np.array_equal(L["min_b"], mn) and np.array_equal(L["div_b"], div)
    key = (ijk - mn) @ np.array([1, div[0], div[0] * div[1]])
    uniq, counts = np.unique(key, return_counts=True)
    assert np.array_equal(L["idx"], uniq.astype(np.uint64))
    assert np.array_equal(np.where(L["n"] == -1, 6, L["n"]) >= 6, counts >= 6), "-1 marks a voxel with >= 6 points rejected by the eigenvalue test"
    checked = 0
    for slot in np.nonzero(L["n"] >= 6)[0][::7]:
        pts = tgt[key == uniq[slot], :3].astype(np.float64)
        n = len(pts)
        assert n == L["n"][slot]
        mean = pts.mean(axis=0)
        np.testing.assert_allclose(L["mean"][slot], mean, rtol=0, atol=1e-12)
        # population covariance x (n - 1) / n — the upstream quirk of A.3 step 3
        cov = np.cov(pts.T, bias=True) * (n - 1.0) / n
        w, V = np.linalg.eigh(cov)
        if w[0] < 0.01 * w[2]:  # eigenvalue clamp (min_covar_eigvalue_mult_ = 0.01)
            w = np.maximum(w, 0.01 * w[2])
            cov = V @ np.diag(w) @ V.T
        scale = np.abs(cov).max()
        assert np.max(np.abs(L["cov"][slot] - cov)) < 1e-9 * scale
        icov = np.linalg.inv(cov)
        assert np.max(np.abs(L["icov"][slot] - icov)) < 1e-7 * np.abs(icov).max()
        np.testing.assert_allclose(L["cov"][slot] @ L["icov"][slot], np.eye(3), atol=1e-8)
        checked += 1
    assert checked > 100
    # voxels under the 6-point rule carry their count and are never probed
    assert np.all(L["n"][counts < 6] == counts[counts < 6])


@pytest.mark.parametrize("search", ["DIRECT7", "DIRECT1", "KDTREE"])
def test_ndt_gradient_and_hessian_match_finite_differences(oracle, scans, search):
    code = dict(KDTREE=0, DIRECT7=2, DIRECT1=3)[search]
    reg = oracle.Registration(oracle.NDT, resolution=1.0, nn_search=code)
    reg.setInputTarget(scans["ds0"])
    reg.setInputSource(scans["ds1"][::4])
    p0 = np.array([0.45, 0.03, -0.01, 0.004, -0.006, 0.012])
    s0, g0, H0 = reg.ndt_derivatives(p0)
    assert s0 > 0 and np.all(np.isfinite(g0)) and np.allclose(H0, H0.T, rtol=1e-6, atol=1e-6 * np.abs(H0).max())
    # per-hit arithmetic is float: central differences with a step well above float noise
    h = np.array([2e-3] * 3 + [2e-4] * 3)
    g_fd = np.zeros(6)
    H_fd = np.zeros((6, 6))
    for i in range(6):
        e = np.zeros(6)
        e[i] = h[i]
        sp, gp, _ = reg.ndt_derivatives(p0 + e, compute_hessian=False)
        sm, gm, _ = reg.ndt_derivatives(p0 - e, compute_hessian=False)
        g_fd[i] = (sp - sm) / (2 * h[i])
        H_fd[:, i] = (gp - gm) / (2 * h[i])
    # the voxel membership of a few points changes inside the stencil (the score is only piecewise
    # smooth), hence percent-level rather than 1e-4 agreement on real scans
    assert np.max(np.abs(g_fd - g0)) < (0.05 if search == "DIRECT1" else 0.03) * np.max(np.abs(g0))
    # upstream's Hessian deliberately deviates from the exact second derivative in one h_ang row
    # (the (+sy) term kept in A.4), and is the Gauss-Newton-style form; compare the dominant block
    assert np.max(np.abs(H_fd[:3, :3] - H0[:3, :3])) < 0.05 * np.max(np.abs(H0[:3, :3]))


def test_ndt_gradient_smooth_synthetic_case(oracle):
    """A target made of fat isotropic blobs and a source inside them: the score is smooth over the
    finite-difference stencil, so the analytic gradient must match to 1e-4 relative."""
    rng = np.random.default_rng(0)
    centres = np.array([[x + 0.5, y + 0.5, z + 0.5] for x in range(-3, 3) for y in range(-3, 3) for z in range(-1, 1)], np.float32)
    tgt = np.ones((len(centres) * 40, 4), np.float32)
    tgt[:, :3] = (centres[:, None, :] + rng.normal(0, 0.12, (len(centres), 40, 3)).clip(-0.45, 0.45)).reshape(-1, 3)
    src = np.ones((len(centres) * 3, 4), np.float32)
    src[:, :3] = (centres[:, None, :] + rng.normal(0, 0.05, (len(centres), 3, 3)).clip(-0.2, 0.2)).reshape(-1, 3)
    reg = oracle.Registration(oracle.NDT, resolution=1.0, nn_search=oracle.DIRECT1)
    reg.setInputTarget(tgt)
    reg.setInputSource(src)
    p0 = np.array([0.02, -0.015, 0.01, 0.002, -0.003, 0.004])
    _, g0, H0 = reg.ndt_derivatives(p0)
    h = np.array([1e-3] * 3 + [1e-3] * 3)
    g_fd = np.zeros(6)
    for i in range(6):
        e = np.zeros(6)
        e[i] = h[i]
        g_fd[i] = (reg.ndt_derivatives(p0 + e, False)[0] - reg.ndt_derivatives(p0 - e, False)[0]) / (2 * h[i])
    assert np.max(np.abs(g_fd - g0)) < 2e-3 * np.max(np.abs(g0))


def test_euler_xyz_is_a_valid_factorisation_with_first_angle_in_0_pi(oracle):
    rng = np.random.default_rng(2)
    for _ in range(200):
        ang = rng.uniform(-0.5, 0.5, 3)
        T = oracle.transform_from_p(np.concatenate([rng.uniform(-1, 1, 3), ang]))
        e = oracle.euler_xyz(T)
        assert 0.0 <= e[0] <= np.pi + 1e-6, "Eigen 3.3 eulerAngles(0,1,2): first angle in [0, pi]"
        T2 = oracle.transform_from_p(np.concatenate([T[:3, 3], e.astype(np.float64)]))
        assert np.max(np.abs(T2 - T)) < 5e-6


@pytest.mark.parametrize("search,tol_t", [("DIRECT7", 0.02), ("DIRECT1", 0.03), ("KDTREE", 0.02)])
def test_ndt_recovers_known_transform(oracle, scans, search, tol_t):
    """The source IS the target moved by a known rigid transform: align from the identity must
    bring it back (pin i: <= 2 cm / 0.1 deg from a perturbed guess)."""
    code = dict(KDTREE=0, DIRECT7=2, DIRECT1=3)[search]
    tgt = scans["ds0"]
    P = oracle.synth_pose(np.array([0.30, -0.12, 0.02, 0.004, -0.003, 0.02]))
    Pinv = np.linalg.inv(P)
    src = np.ones_like(tgt[::2])
    src[:, :3] = (tgt[::2, :3].astype(np.float64) @ Pinv[:3, :3].T + Pinv[:3, 3]).astype(np.float32)
    reg = oracle.Registration(oracle.NDT, resolution=1.0, nn_search=code, trans_eps=0.001, max_iter=64)
    reg.setInputTarget(tgt)
    reg.setInputSource(src)
    reg.align(None)
    T = reg.getFinalTransformation().astype(np.float64)
    assert reg.hasConverged()
    assert np.max(np.abs(T[:3, 3] - P[:3, 3])) < tol_t
    R = P[:3, :3].T @ T[:3, :3]
    assert np.arccos(np.clip((np.trace(R) - 1) / 2, -1, 1)) < np.deg2rad(0.1)
    assert reg.getFitnessScore() < 0.01


def test_ndt_align_contract(oracle, scans):
    """pcl::Registration::align semantics: max_iterations caps the Newton loop, the aligned cloud is
    the source under the final transform, an empty source leaves converged_ false."""
    reg = oracle.Registration(oracle.NDT, resolution=1.0, nn_search=oracle.DIRECT7, trans_eps=1e-9, max_iter=3)
    reg.setInputTarget(scans["ds0"])
    reg.setInputSource(scans["ds1"])
    out = reg.align(None, want_aligned=True)
    assert reg.getFinalNumIteration() <= 3 + 2
    T = reg.getFinalTransformation()
    want = (scans["ds1"][:, :3] @ T[:3, :3].T + T[:3, 3]).astype(np.float32)
    assert np.max(np.abs(out[:, :3] - want)) < 1e-4
