"""Host-side mirror of hdl_graph_slam::InformationMatrixCalculator
[REF include/hdl_graph_slam/information_matrix_calculator.hpp:20-49; src/hdl_graph_slam/information_matrix_calculator.cpp:28-108].

The reference calls calc_information_matrix(cloud1, cloud2, relpose) once per odometry edge and once
per accepted loop [REF apps/delta_graph_slam_nodelet.cpp:572,820]; every call builds a fresh kd-tree on
cloud1 and runs a serial nearest-neighbour loop over cloud2 (calc_fitness_score, :77-108).  Here the
fitness comes from the engine's exact-NN kernels; the edge weighting (`weight`, the 3x3 information
matrix of the 2-D pose graph) stays on the host, where it is a handful of flops.
"""
import math

import numpy as np

from .registration import DBL_MAX, Registration


class InformationMatrixCalculator:
    def __init__(self, params=None, device=0, engine=None):
        p = dict(params or {})
        # constructor defaults [REF information_matrix_calculator.cpp:28-39]
        self.use_const_inf_matrix = bool(p.get("use_const_inf_matrix", False))
        self.const_stddev_x = float(p.get("const_stddev_x", 0.5))
        self.const_stddev_q = float(p.get("const_stddev_q", 0.1))
        self.var_gain_a = float(p.get("var_gain_a", 20.0))
        self.min_stddev_x = float(p.get("min_stddev_x", 0.1))
        self.max_stddev_x = float(p.get("max_stddev_x", 5.0))
        self.min_stddev_q = float(p.get("min_stddev_q", 0.05))
        self.max_stddev_q = float(p.get("max_stddev_q", 0.2))
        self.fitness_score_thresh = float(p.get("fitness_score_thresh", 0.5))
        self._engine = engine  # a Registration whose keyframe cache holds the clouds (batch calls)
        self._single = None
        self._device = device

    # ---- the reference's private helper [REF information_matrix_calculator.hpp:46-49]
    def weight(self, a, max_x, min_y, max_y, x):
        y = (1.0 - math.exp(-a * x)) / (1.0 - math.exp(-a * max_x))
        return min_y + (max_y - min_y) * y

    def _matrix(self, fitness_score):
        """[REF information_matrix_calculator.cpp:53-75]; w_x / w_q are floats there."""
        inf = np.eye(3)
        if self.use_const_inf_matrix:
            inf[:2, :2] /= self.const_stddev_x
            inf[2:, 2:] /= self.const_stddev_q
            return inf
        min_var_x, max_var_x = self.min_stddev_x ** 2, self.max_stddev_x ** 2
        min_var_q, max_var_q = self.min_stddev_q ** 2, self.max_stddev_q ** 2
        w_x = float(np.float32(self.weight(self.var_gain_a, self.fitness_score_thresh, min_var_x, max_var_x, fitness_score)))
        w_q = float(np.float32(self.weight(self.var_gain_a, self.fitness_score_thresh, min_var_q, max_var_q, fitness_score)))
        inf[:2, :2] /= w_x
        inf[2:, 2:] /= w_q
        return inf

    # ---- calc_fitness_score(cloud1, cloud2, relpose, max_range) [REF :77-108]
    def calc_fitness_score(self, cloud1, cloud2, relpose, max_range=DBL_MAX):
        if self._single is None:
            self._single = Registration(device=self._device)
        self._single.setInputTarget(cloud1)
        self._single.setInputSource(cloud2)
        return self._single.calcFitnessScore(np.asarray(relpose, np.float64).astype(np.float32), max_range)

    def calc_information_matrix(self, cloud1, cloud2, relpose):
        if self.use_const_inf_matrix:
            return self._matrix(0.0)
        return self._matrix(self.calc_fitness_score(cloud1, cloud2, relpose))

    # ---- many edges at once on cached keyframes: [(id1, id2, relpose), ...]
    def calc_information_matrices(self, edges, max_range=DBL_MAX):
        edges = list(edges)
        if self.use_const_inf_matrix:
            return [self._matrix(0.0) for _ in edges]
        if self._engine is None:
            raise ValueError("calc_information_matrices needs the engine whose keyframe cache holds the clouds")
        scores = self._engine.calcFitnessBatch([(a, b, np.asarray(T, np.float64).astype(np.float32)) for a, b, T in edges], max_range)
        return [self._matrix(float(s)) for s in scores]
