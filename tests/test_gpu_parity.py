"""GPU parity tests: the CUDA engine (through the C ABI) against the CPU oracle on the same
seeded synthetic scans.  Bars (BASELINE.json): voxel assignment / counts / order bit-exact;
final transforms within 1e-4 m / 1e-4 rad; fitness within 1e-5 relative."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL_T = 1e-4    # metres
TOL_R = 1e-4    # radians
TOL_FIT = 1e-5  # relative


def rot_angle(Ra, Rb):
    """Rotation angle between two float32 rotation matrices.  arccos of the trace cannot resolve
    angles below ~3e-4 rad from float32 entries (1 - cos(theta) drops under the float epsilon), so
    the angle is taken from the skew part, sin(theta) = |vee(R - R^T)| / 2."""
    R = Ra.astype(np.float64).T @ Rb.astype(np.float64)
    w = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / 2.0
    s = float(np.linalg.norm(w))
    c = float((np.trace(R) - 1.0) / 2.0)
    return float(np.arctan2(s, c))


@pytest.fixture(scope="module")
def eng():
    import delta_graph_slam_b200 as d
    return d


def test_voxelgrid_bit_exact(eng, oracle, scans):
    vg = eng.VoxelGrid()
    vg.setLeafSize(0.1, 0.1, 0.1)
    for raw in (scans["raw0"], scans["raw1"]):
        ref = oracle.voxelgrid(raw, 0.1, is_dense=False)
        vg.setInputCloud(raw, is_dense=False)
        out = vg.filter()
        lay = vg.last_layout(len(out), len(raw))
        assert len(out) == len(ref["out"])
        assert np.array_equal(lay["min_b"], ref["min_b"]) and np.array_equal(lay["div_b"], ref["div_b"])
        assert np.array_equal(lay["key"], ref["key"]), "per-point voxel assignment must be bit-exact"
        assert np.array_equal(lay["voxel_id"], ref["voxel_id"]), "output order (ascending voxel index) must be bit-exact"
        assert np.array_equal(lay["count"], ref["count"]), "per-voxel point counts must be bit-exact"
        # both sides sum in ascending input order without contraction -> identical bits
        assert np.array_equal(out.view(np.uint32), ref["out"].view(np.uint32))


def test_voxelgrid_edge_cases(eng, oracle):
    vg = eng.VoxelGrid()
    vg.setLeafSize(0.1, 0.1, 0.1)
    # empty
    vg.setInputCloud(np.zeros((0, 4), np.float32))
    assert len(vg.filter()) == 0
    # single point, duplicates, NaN rows skipped when !is_dense
    rng = np.random.default_rng(3)
    pts = np.ones((1000, 4), np.float32)
    pts[:, :3] = rng.uniform(-5, 5, (1000, 3)).astype(np.float32)
    pts[::7, 1] = np.nan
    pts[5::31, 0] = np.inf
    pts[10:20] = pts[9]
    ref = oracle.voxelgrid(pts, 0.25, is_dense=False)
    vg.setLeafSize(0.25, 0.25, 0.25)
    vg.setInputCloud(pts, is_dense=False)
    out = vg.filter()
    lay = vg.last_layout(len(out), len(pts))
    assert np.array_equal(lay["key"], ref["key"])
    assert np.array_equal(lay["count"], ref["count"])
    assert np.array_equal(out.view(np.uint32), ref["out"].view(np.uint32))
    # leaf too small for the extent: PCL returns the input unfiltered
    far = np.ones((4, 4), np.float32)
    far[:, :3] = [[0, 0, 0], [4000, 0, 0], [0, 4000, 0], [0, 0, 4000]]
    vg.setLeafSize(0.001, 0.001, 0.001)
    vg.setInputCloud(far, is_dense=True)
    out = vg.filter()
    ref = oracle.voxelgrid(far, 0.001, is_dense=True)
    assert ref["overflow"] and len(out) == 4 and np.array_equal(out, far)
    # anisotropic leaf + min_points_per_voxel
    ref = oracle.voxelgrid(pts, (0.5, 0.25, 1.0), min_points_per_voxel=2, is_dense=False)
    vg.setLeafSize(0.5, 0.25, 1.0)
    vg.setMinimumPointsNumberPerVoxel(2)
    vg.setInputCloud(pts, is_dense=False)
    out = vg.filter()
    assert len(out) == len(ref["out"]) and np.array_equal(out.view(np.uint32), ref["out"].view(np.uint32))


@pytest.mark.parametrize("path", [0, 1, 2, 3])
def test_voxelgrid_edge_cases_on_every_sort_path(eng, oracle, path):
    """The ragged and degenerate clouds through each sort path (0 default, 1 three launches per digit, 2 one-sweep with the
    gather in its last pass, 3 cooperative): sizes around the tile and warp boundaries of the one-sweep passes (2048-key
    tiles, 256-record centroid tiles), a cloud whose points all share ONE voxel (no significant key bit: the one-sweep
    path's forced first pass), one voxel holding thousands of points (a run across many centroid tiles), NaN rows, the
    distance gate letting nothing through, and PCL's leaf-too-small copy."""
    from delta_graph_slam_b200 import _lib
    L = _lib.load()
    rng = np.random.default_rng(17 + path)
    try:
        assert L.b200reg_set_sort_path(path) == 0
        vg = eng.VoxelGrid()
        for n in (1, 2, 31, 32, 33, 255, 256, 257, 2047, 2048, 2049, 4097, 6000):
            pts = np.ones((n, 4), np.float32)
            pts[:, :3] = rng.uniform(-3, 3, (n, 3)).astype(np.float32)
            if n > 40:
                pts[::13, 2] = np.nan
            vg.setLeafSize(0.3, 0.3, 0.3)
            vg.setInputCloud(pts, is_dense=False)
            ref = oracle.voxelgrid(pts, 0.3, is_dense=False)
            out = vg.filter()
            assert len(out) == len(ref["out"]) and np.array_equal(out.view(np.uint32), ref["out"].view(np.uint32)), n
        # all points in one voxel
        one = np.ones((5000, 4), np.float32)
        one[:, :3] = rng.uniform(0.01, 0.09, (5000, 3)).astype(np.float32)
        vg.setLeafSize(0.1, 0.1, 0.1)
        vg.setInputCloud(one, is_dense=True)
        ref = oracle.voxelgrid(one, 0.1, is_dense=True)
        out = vg.filter()
        assert len(ref["out"]) == 1 and np.array_equal(out.view(np.uint32), ref["out"].view(np.uint32))
        # one crowded voxel among sparse ones: its run spans many 256-record tiles
        mix = np.ones((9000, 4), np.float32)
        mix[:, :3] = rng.uniform(-20, 20, (9000, 3)).astype(np.float32)
        mix[1000:7000, :3] = rng.uniform(5.01, 5.09, (6000, 3)).astype(np.float32)
        vg.setInputCloud(mix, is_dense=True)
        ref = oracle.voxelgrid(mix, 0.1, is_dense=True)
        out = vg.filter()
        assert ref["count"].max() >= 6000 and np.array_equal(out.view(np.uint32), ref["out"].view(np.uint32))
        # the gate lets nothing through; then PCL's leaf-too-small copy with the gate on
        vg.setDistanceFilter(True, 500.0, 600.0)
        assert len(vg.filter()) == 0
        vg.setDistanceFilter(True, 1.0, 100.0)
        vg.setLeafSize(1e-4, 1e-4, 1e-4)
        gated = oracle.distance_filter(mix, 1.0, 100.0)
        assert np.array_equal(vg.filter().view(np.uint32), gated.view(np.uint32))
    finally:
        L.b200reg_set_sort_path(0)


def test_ndt_target_grid(eng, oracle, scans):
    ref = oracle.Registration(oracle.NDT, resolution=1.0)
    ref.setInputTarget(scans["ds0"])
    L = ref.ndt_leaves()
    ndt = eng.NormalDistributionsTransform()
    ndt.setResolution(1.0)
    ndt.setInputTarget(scans["ds0"])
    G = ndt.ndt_leaves()
    assert np.array_equal(G["min_b"], L["min_b"]) and np.array_equal(G["div_b"], L["div_b"])
    assert np.array_equal(G["idx"], L["idx"]), "occupied voxel set must be bit-exact"
    assert np.array_equal(G["n"], L["n"]), "per-voxel counts / validity must be bit-exact"
    valid = L["n"] >= 6
    assert valid.sum() > 1000
    np.testing.assert_allclose(G["mean"], L["mean"], rtol=0, atol=1e-12)
    scale = np.abs(L["cov"][valid]).max(axis=(1, 2), keepdims=True)
    assert np.max(np.abs(G["cov"][valid] - L["cov"][valid]) / scale) < 1e-9
    scale = np.abs(L["icov"][valid]).max(axis=(1, 2), keepdims=True)
    assert np.max(np.abs(G["icov"][valid] - L["icov"][valid]) / scale) < 1e-7


@pytest.mark.parametrize("search", ["DIRECT7", "DIRECT1", "KDTREE", "DIRECT26"])
def test_ndt_derivatives(eng, oracle, scans, search):
    code = dict(KDTREE=0, DIRECT26=1, DIRECT7=2, DIRECT1=3)[search]
    ref = oracle.Registration(oracle.NDT, resolution=1.0, nn_search=code)
    ref.setInputTarget(scans["ds0"])
    ref.setInputSource(scans["ds1"])
    ndt = eng.NormalDistributionsTransform()
    ndt.setResolution(1.0)
    ndt.setNeighborhoodSearchMethod(code)
    ndt.setInputTarget(scans["ds0"])
    ndt.setInputSource(scans["ds1"])
    for p in ([0, 0, 0, 0, 0, 0], [1.0, 0.0, 0.0, 0.0, 0.0, 0.0063], [0.4, -0.1, 0.03, 0.01, -0.02, 0.05], [0.9, 0.05, 0.0, 3.13, 3.12, -3.1]):
        p = np.array(p, np.float64)
        s0, g0, H0 = ref.ndt_derivatives(p)
        s1, g1, H1 = ndt.ndt_derivatives(p)
        assert abs(s1 - s0) <= 2e-6 * abs(s0)
        assert np.max(np.abs(g1 - g0)) <= 2e-6 * np.max(np.abs(g0)) + 1e-3
        assert np.max(np.abs(H1 - H0)) <= 2e-6 * np.max(np.abs(H0))


@pytest.mark.parametrize("search", ["DIRECT7", "DIRECT1", "KDTREE"])
def test_ndt_align_parity(eng, oracle, scans, search):
    code = dict(KDTREE=0, DIRECT7=2, DIRECT1=3)[search]
    ref = oracle.Registration(oracle.NDT, resolution=1.0, nn_search=code, trans_eps=0.01, max_iter=64)
    ndt = eng.select_registration_method(dict(registration_method="NDT_OMP", reg_resolution=1.0, reg_nn_search_method=search))
    ref.setInputTarget(scans["ds0"])
    ndt.setInputTarget(scans["ds0"])
    ref.setInputSource(scans["ds1"])
    ndt.setInputSource(scans["ds1"])
    yaw = 0.01
    guess = np.eye(4, dtype=np.float32)
    guess[:2, :2] = [[np.cos(yaw), -np.sin(yaw)], [np.sin(yaw), np.cos(yaw)]]
    guess[:3, 3] = [0.7, 0.05, 0.0]
    for g in (None, guess):
        ref.align(g)
        aligned = ndt.align(g, want_aligned=True)
        T0, T1 = ref.getFinalTransformation(), ndt.getFinalTransformation()
        print(f"{search}: iters gpu={ndt.getFinalNumIteration()} ref={ref.getFinalNumIteration()} dt={np.abs(T1[:3, 3] - T0[:3, 3])} dR={rot_angle(T0[:3, :3], T1[:3, :3]):.3e} "
              f"fit gpu={ndt.getFitnessScore():.10f} ref={ref.getFitnessScore():.10f}")
        assert ndt.hasConverged() == ref.hasConverged()
        assert ndt.getFinalNumIteration() == ref.getFinalNumIteration(), "same Newton / line-search path"
        assert np.max(np.abs(T1[:3, 3] - T0[:3, 3])) < TOL_T
        assert rot_angle(T0[:3, :3], T1[:3, :3]) < TOL_R
        info = ref.info()
        res = ndt.getResult()
        assert res["evaluations"] == int(info[1])
        assert abs(res["score"] - info[0]) <= 1e-5 * abs(info[0])
        # aligned cloud = final transform applied to the source
        ref_al = (scans["ds1"][:, :3] @ T1[:3, :3].T + T1[:3, 3]).astype(np.float32)
        assert np.max(np.abs(aligned[:, :3] - ref_al)) < 1e-4
        f0, f1 = ref.getFitnessScore(), ndt.getFitnessScore()
        assert abs(f1 - f0) <= TOL_FIT * f0
        f0, f1 = ref.getFitnessScore(4.0), ndt.getFitnessScore(4.0)
        assert abs(f1 - f0) <= TOL_FIT * f0


def test_closing_hessian_of_a_line_search_float_per_hit_against_upstream_double(eng, oracle, scans):
    """Upstream closes a More-Thuente search that took extra trials with computeHessian — a serial, all-double sweep
    (oracle_ndt.cpp NDT::computeHessian, A.4) — where every other Hessian comes from computeDerivatives' float per-hit
    algebra.  The engine takes the float per-hit Hessian its last trial pass accumulated anyway (csrc/ndt_align.cuh).
    This quantifies that substitution at the level where it could matter: the Hessian itself and the Newton step it
    yields, and then checks registrations whose line searches DID iterate against the oracle (which runs the double sweep)."""
    ref = oracle.Registration(oracle.NDT, resolution=1.0, nn_search=oracle.DIRECT7, trans_eps=0.01, max_iter=64)
    ndt = eng.select_registration_method(dict(registration_method="NDT_OMP", reg_resolution=1.0, reg_nn_search_method="DIRECT7"))
    for r in (ref, ndt):
        r.setInputTarget(scans["ds0"])
        r.setInputSource(scans["ds1"])
    worst_h, worst_step = 0.0, 0.0
    for p in ([0, 0, 0, 0, 0, 0], [0.4, -0.1, 0.03, 0.01, -0.02, 0.05], [0.5, 0.0, 0.0, 0.0, 0.0, 0.003], [0.9, 0.05, 0.0, 3.13, 3.12, -3.1], [0.2, 0.3, -0.05, -0.03, 0.02, -0.08]):
        p = np.array(p, np.float64)
        _, g, H_engine = ndt.ndt_derivatives(p)       # float per-hit, double sums: what the engine's trial pass holds
        H_double = ref.ndt_hessian(p)                  # upstream's closing sweep
        rel = np.max(np.abs(H_engine - H_double)) / np.max(np.abs(H_double))
        step_e, step_d = np.linalg.solve(H_engine, -g), np.linalg.solve(H_double, -g)
        # upstream normalises delta_p and walks at most step_size = 0.1 along it: the pose change of the next iteration
        walk = lambda d: d / np.linalg.norm(d) * min(np.linalg.norm(d), 0.1)
        worst_h, worst_step = max(worst_h, rel), max(worst_step, float(np.max(np.abs(walk(step_e) - walk(step_d)))))
    print(f"closing Hessian: float-per-hit vs double sweep, max relative entry difference {worst_h:.2e}, max difference of the next Newton step {worst_step:.2e} (m | rad)")
    assert worst_h < 2e-6 and worst_step < 1e-6  # two orders below the 1e-4 parity bar
    # registrations whose line search took more than one trial (evaluations > passes: a closing sweep was folded away)
    folded = 0
    for k, (tx, ty, yaw) in enumerate([(0.0, 0.0, 0.0), (0.7, 0.05, 0.01), (1.6, -0.3, 0.03), (-0.4, 0.5, -0.05), (2.2, 0.0, 0.0), (1.0, 0.8, 0.08), (0.2, -0.3, -0.02)]):
        guess = np.eye(4, dtype=np.float32)
        guess[:2, :2] = [[np.cos(yaw), -np.sin(yaw)], [np.sin(yaw), np.cos(yaw)]]
        guess[:3, 3] = [tx, ty, 0.0]
        ref.align(guess)
        ndt.align(guess)
        res, info = ndt.getResult(), ref.info()
        assert res["iterations"] == ref.getFinalNumIteration() and res["evaluations"] == int(info[1]), f"guess {k}: same path"
        T0, T1 = ref.getFinalTransformation(), ndt.getFinalTransformation()
        if res["evaluations"] > res["passes"]:
            folded += 1
            assert np.max(np.abs(T1[:3, 3] - T0[:3, 3])) < TOL_T and rot_angle(T0[:3, :3], T1[:3, :3]) < TOL_R, f"guess {k}"
    assert folded >= 2, "the guesses must exercise line searches with extra trials"


def test_ndt_recovers_ground_truth(eng, scans):
    ndt = eng.select_registration_method(dict(registration_method="NDT_OMP", reg_resolution=1.0))
    ndt.setInputTarget(scans["ds0"])
    ndt.setInputSource(scans["ds1"])
    ndt.align(None)
    T = ndt.getFinalTransformation()
    gt = scans["gt"]
    assert np.max(np.abs(T[:3, 3] - gt[:3, 3])) < 0.02
    assert rot_angle(T[:3, :3], gt[:3, :3].astype(np.float32)) < np.deg2rad(0.1)


def test_fitness_exact_nn(eng, oracle, scans):
    """Same transform on both sides: the exact-NN fitness must agree to rounding of the double sum."""
    ref = oracle.Registration(oracle.NDT, resolution=1.0)
    reg = eng.NormalDistributionsTransform()
    ref.setInputTarget(scans["ds0"])
    reg.setInputTarget(scans["ds0"])
    ref.setInputSource(scans["ds1"])
    reg.setInputSource(scans["ds1"])
    # identity (final_transformation_ before any align) and far-off transforms exercise the
    # ring search, the max_range gate and the brute-force pass for far outliers
    assert abs(reg.getFitnessScore() - ref.getFitnessScore()) <= 1e-12 * ref.getFitnessScore()
    idx, d2 = oracle.knn(scans["ds0"], scans["ds1"], 1)
    for T in (np.eye(4), scans["gt"], np.array([[1, 0, 0, 30.0], [0, 1, 0, -20.0], [0, 0, 1, 5.0], [0, 0, 0, 1]])):
        T = np.asarray(T, np.float32)
        q = np.ones_like(scans["ds1"])
        # float transform in pcl::transformPoint order
        for r in range(3):
            q[:, r] = ((T[r, 0] * scans["ds1"][:, 0] + T[r, 1] * scans["ds1"][:, 1]) + T[r, 2] * scans["ds1"][:, 2]) + T[r, 3]
        _, d2 = oracle.knn(scans["ds0"], q, 1)
        d2 = d2[:, 0].astype(np.float64)
        for max_range in (np.finfo(np.float64).max, 1.0, 0.01):
            sel = d2 <= max_range
            want = d2[sel].sum() / sel.sum() if sel.any() else np.finfo(np.float64).max
            got = reg.calcFitnessScore(T, max_range)
            assert abs(got - want) <= 1e-12 * want
        frac = reg.getInlierFraction(0.5)
        assert 0.0 <= frac <= 1.0
    # publish_scan_matching_status [REF apps/scan_matching_odometry_nodelet.cpp:318-332]: matching_error
    # and inlier fraction (k_sq_dists[0] < 0.5^2) of the aligned cloud after a real align
    ref.align(None)
    aligned = reg.align(None, want_aligned=True)
    for max_dist in (0.5, 0.1):
        want = ref.inlierFraction(aligned, max_dist)
        got = reg.getInlierFraction(max_dist)
        assert abs(got - want) <= 2.0 / len(aligned), (got, want)


def test_odometry_keyframe_promotion(eng, oracle, scans):
    """promoteSourceToTarget (keyframe = last source; setInputTarget(keyframe)) must equal uploading the same cloud again,
    and as in PCL the cloud stays the input source until a new one is set."""
    a = eng.NormalDistributionsTransform()
    b = eng.NormalDistributionsTransform()
    for r in (a, b):
        r.setResolution(1.0)
        r.setInputTarget(scans["ds0"])
    src = scans["ds1"]
    a.setInputSource(src)
    a.promoteSourceToTarget()      # changes role on the device
    b.setInputTarget(src.copy())   # uploaded
    La, Lb = a.ndt_leaves(), b.ndt_leaves()
    assert np.array_equal(La["idx"], Lb["idx"]) and np.array_equal(La["n"], Lb["n"])
    assert np.array_equal(La["icov"], Lb["icov"])
    # pcl still holds the promoted cloud as input_: an align now registers it against itself
    b.setInputSource(src)
    for r in (a, b):
        r.align(None)
    assert a.hasConverged() and np.array_equal(a.getFinalTransformation(), b.getFinalTransformation())
    assert a.getFitnessScore() == b.getFitnessScore() and a.getFitnessScore() < 1e-9
    # refilling the caller's buffer in place after setInputSource must not leak into a later setInputTarget of that array
    c = eng.NormalDistributionsTransform()
    c.setResolution(1.0)
    buf = scans["ds0"].copy()
    c.setInputTarget(scans["ds0"])
    c.setInputSource(buf)
    buf[: len(scans["ds1"])] = scans["ds1"][: len(buf)]
    c.setInputTarget(buf)          # the CURRENT contents are uploaded (no identity shortcut)
    d = eng.NormalDistributionsTransform()
    d.setResolution(1.0)
    d.setInputTarget(buf.copy())
    Lc, Ld = c.ndt_leaves(), d.ndt_leaves()
    assert np.array_equal(Lc["idx"], Ld["idx"]) and np.array_equal(Lc["icov"], Ld["icov"])


def test_cooperative_and_multi_kernel_sort_paths_agree(eng, oracle, scans):
    """The voxel key / sort / segmentation pipeline as ONE cooperative kernel (default at this size) against the
    three-launch-per-digit path and the one-sweep chained-scan sort: identical keys, order, runs, centroids and NDT
    leaves, bit for bit (all three are stable sorts of the same keys)."""
    from delta_graph_slam_b200 import _lib
    L = _lib.load()
    outs = []
    try:
        for path in (0, 1, 2, 3):
            assert L.b200reg_set_sort_path(path) == 0
            vg = eng.VoxelGrid()
            vg.setLeafSize(0.1, 0.1, 0.1)
            raw = scans["raw0"].copy()
            raw[::17, 0] = np.nan
            vg.setInputCloud(raw, is_dense=False)
            out = vg.filter()
            lay = vg.last_layout(len(out), len(raw))
            ndt = eng.NormalDistributionsTransform()
            ndt.setResolution(1.0)
            ndt.setInputTarget(scans["ds0"])
            G = ndt.ndt_leaves()
            small = eng.VoxelGrid()
            small.setLeafSize(0.5, 0.5, 0.5)
            small.setInputCloud(scans["ds1"][:777], is_dense=True)
            outs.append((out, lay, G, small.filter()))
    finally:
        L.b200reg_set_sort_path(0)
    assert L.b200reg_set_sort_path(4) != 0
    (o0, l0, g0, s0) = outs[0]
    for path, (o1, l1, g1, s1) in enumerate(outs[1:], start=1):
        assert np.array_equal(o0.view(np.uint32), o1.view(np.uint32)) and np.array_equal(s0.view(np.uint32), s1.view(np.uint32)), path
        for k in ("voxel_id", "count", "key", "min_b", "div_b"):
            assert np.array_equal(l0[k], l1[k]), (path, k)
        for k in ("idx", "n", "mean", "cov", "icov"):
            assert np.array_equal(g0[k], g1[k]), (path, k)
    ref = oracle.voxelgrid(scans["raw0"], 0.1, is_dense=False)
    vg = eng.VoxelGrid()
    vg.setLeafSize(0.1, 0.1, 0.1)
    vg.setInputCloud(scans["raw0"], is_dense=False)
    assert np.array_equal(vg.filter().view(np.uint32), ref["out"].view(np.uint32))
