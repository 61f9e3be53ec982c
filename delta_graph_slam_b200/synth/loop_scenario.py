"""Synthetic loop-closure batch (bench / test infrastructure): BASELINE.json configs[3].

`n_targets` new keyframes spaced along the kitti_like trajectory, each with `n_candidates`
candidate keyframes scanned from poses within U(+-dx, +-dy, +-dyaw) of the target pose; the initial
guess of a pair is its true relative pose perturbed by U(+-0.5 m, +-0.5 m, +-2 deg) and projected
to (x, y, yaw) the way LoopDetector::matching builds it from the 2-D graph estimates
[REF include/hdl_graph_slam/loop_detector.hpp:139-143; src/hdl_graph_slam/ros_utils.cpp:105-126].
Only poses, seeds and ids are produced here; the caller ray-casts the scans (GPU generator in
bench.py, CPU generator in the tests).  The lateral offset stays inside the street (the parked
cars of the scene start 3.6 m from the centre line).
"""
import numpy as np


def _pose(x, y, z, roll, pitch, yaw):
    cx, sx, cy, sy, cz, sz = np.cos(roll), np.sin(roll), np.cos(pitch), np.sin(pitch), np.cos(yaw), np.sin(yaw)
    T = np.eye(4)
    T[:3, :3] = [[cz * cy, cz * sy * sx - sz * cx, cz * sy * cx + sz * sx], [sz * cy, sz * sy * sx + cz * cx, sz * sy * cx - cz * sx], [-sy, cy * sx, cy * cx]]
    T[:3, 3] = [x, y, z]
    return T


def loop_scenario(traj, n_targets=256, n_candidates=16, spacing_frames=8, seed=11, first_frame=0, dx=3.0, dy=1.5, dyaw_deg=10.0, guess_dxy=0.5, guess_dyaw_deg=2.0):
    """traj(k) -> 4x4 sensor pose of frame k.  Returns dict(targets=[(id, pose, noise_seed)],
    candidates=[(id, pose, noise_seed)], pairs=[(target_id, candidate_id, guess4x4 float32, true_rel 4x4)])."""
    from ..loop_detector import transform2Dto3D
    rng = np.random.default_rng(seed)
    targets, candidates, pairs = [], [], []
    for t in range(n_targets):
        Pt = np.asarray(traj(first_frame + t * spacing_frames), np.float64)
        targets.append((t, Pt, 100000 + t))
        for c in range(n_candidates):
            cid = n_targets + t * n_candidates + c
            off = _pose(rng.uniform(-dx, dx), rng.uniform(-dy, dy), 0.0, 0.0, 0.0, np.deg2rad(rng.uniform(-dyaw_deg, dyaw_deg)))
            Pc = Pt @ off
            candidates.append((cid, Pc, 200000 + cid))
            rel = np.linalg.inv(Pt) @ Pc  # candidate (source) points -> target frame
            yaw = np.arctan2(rel[1, 0], rel[0, 0]) + np.deg2rad(rng.uniform(-guess_dyaw_deg, guess_dyaw_deg))
            g2 = np.array([[np.cos(yaw), -np.sin(yaw), rel[0, 3] + rng.uniform(-guess_dxy, guess_dxy)], [np.sin(yaw), np.cos(yaw), rel[1, 3] + rng.uniform(-guess_dxy, guess_dxy)], [0, 0, 1.0]])
            pairs.append((t, cid, transform2Dto3D(g2.astype(np.float32)), rel))
    return dict(targets=targets, candidates=candidates, pairs=pairs)
