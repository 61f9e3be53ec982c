"""InformationMatrixCalculator mirror [REF src/hdl_graph_slam/information_matrix_calculator.cpp:28-108,
include/hdl_graph_slam/information_matrix_calculator.hpp:46-49; call sites apps/delta_graph_slam_nodelet.cpp:572,820]."""
import math

import numpy as np
import pytest


def ref_weight(a, max_x, min_y, max_y, x):
    y = (1.0 - math.exp(-a * x)) / (1.0 - math.exp(-a * max_x))
    return min_y + (max_y - min_y) * y


def test_information_matrix_weighting_host_logic():
    from delta_graph_slam_b200.information_matrix import InformationMatrixCalculator
    c = InformationMatrixCalculator()
    assert (c.var_gain_a, c.min_stddev_x, c.max_stddev_x, c.min_stddev_q, c.max_stddev_q, c.fitness_score_thresh) == (20.0, 0.1, 5.0, 0.05, 0.2, 0.5)
    for f in (0.0, 0.01, 0.2, 0.5, 3.0):
        M = c._matrix(f)
        wx = float(np.float32(ref_weight(20.0, 0.5, 0.1 ** 2, 5.0 ** 2, f)))
        wq = float(np.float32(ref_weight(20.0, 0.5, 0.05 ** 2, 0.2 ** 2, f)))
        want = np.diag([1.0 / wx, 1.0 / wx, 1.0 / wq])
        assert np.array_equal(M, want)
    k = InformationMatrixCalculator(dict(use_const_inf_matrix=True, const_stddev_x=0.5, const_stddev_q=0.1))
    assert np.array_equal(k.calc_information_matrix(None, None, np.eye(4)), np.diag([2.0, 2.0, 10.0]))
    assert all(np.array_equal(m, np.diag([2.0, 2.0, 10.0])) for m in k.calc_information_matrices([(0, 1, np.eye(4))]))


@pytest.mark.gpu
def test_fitness_and_information_matrix_match_the_in_tree_loop(oracle, scans):
    import delta_graph_slam_b200 as eng
    from test_oracle_kdtree import reference_fitness
    tgt, src = scans["ds0"][::4].copy(), scans["ds1"][::6].copy()
    T = scans["gt"].astype(np.float64)
    calc = eng.InformationMatrixCalculator()
    for max_range in (np.finfo(np.float64).max, 0.05):
        want = reference_fitness(tgt, src, T.astype(np.float32), max_range)
        got = calc.calc_fitness_score(tgt, src, T, max_range)
        assert abs(got - want) <= 1e-12 * want
    M = calc.calc_information_matrix(tgt, src, T)
    assert np.array_equal(M, calc._matrix(reference_fitness(tgt, src, T.astype(np.float32), np.finfo(np.float64).max)))


@pytest.mark.gpu
def test_batched_edges_equal_single_calls(oracle, scans):
    import delta_graph_slam_b200 as eng
    clouds = {0: scans["ds0"], 1: scans["ds1"], 2: scans["ds1"][::2].copy(), 3: np.zeros((0, 4), np.float32)}
    T = scans["gt"].astype(np.float64)
    Ti = np.linalg.inv(T)
    edges = [(0, 1, T), (1, 0, Ti), (0, 2, T), (1, 2, np.eye(4)), (0, 3, T)]
    reg = eng.select_registration_method(dict(registration_method="NDT_OMP", reg_resolution=1.0), out=open("/dev/null", "w"))
    for cid, c in clouds.items():
        reg.cloudPut(cid, c)
    calc = eng.InformationMatrixCalculator(engine=reg)
    single = eng.InformationMatrixCalculator()
    for max_range in (np.finfo(np.float64).max, 0.25):
        got = reg.calcFitnessBatch([(a, b, np.asarray(P, np.float32)) for a, b, P in edges], max_range)
        for (a, b, P), g in zip(edges, got):
            want = single.calc_fitness_score(clouds[a], clouds[b], P, max_range) if len(clouds[b]) else np.finfo(np.float64).max
            assert g == want or abs(g - want) <= 1e-12 * want, (a, b, g, want)  # double sums in a different fixed order
    mats = calc.calc_information_matrices(edges[:4])
    for (a, b, P), M in zip(edges[:4], mats):
        assert np.allclose(M, single.calc_information_matrix(clouds[a], clouds[b], P), rtol=1e-6, atol=0)
    # fitness-only batches need no NDT handle
    plain = eng.Registration()
    for cid, c in clouds.items():
        plain.cloudPut(cid, c)
    a, b = plain.calcFitnessBatch([(0, 1, T.astype(np.float32))])[0], single.calc_fitness_score(clouds[0], clouds[1], T)
    assert abs(a - b) <= 1e-12 * b
