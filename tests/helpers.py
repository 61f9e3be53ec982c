"""Test helpers: an oracle-backed stand-in for the engine's batch surface (cloudPut / alignBatch),
driven through the reference's serial sequence — setInputTarget once per target, then per candidate
setInputSource, align, getFitnessScore [REF include/hdl_graph_slam/loop_detector.hpp:124-156]."""
import numpy as np

from delta_graph_slam_b200 import _lib


class OracleBatchEngine:
    def __init__(self, oracle, resolution=1.0, nn_search=2, trans_eps=0.01, max_iter=64):
        self.reg = oracle.Registration(oracle.NDT, resolution=resolution, nn_search=nn_search, trans_eps=trans_eps, max_iter=max_iter)
        self.clouds = {}
        self.calls = 0

    def cloudPut(self, cloud_id, cloud):
        self.clouds[int(cloud_id)] = np.ascontiguousarray(cloud, np.float32)

    def alignBatch(self, pairs, with_fitness=True, fitness_max_range=np.finfo(np.float64).max):
        if not (isinstance(pairs, np.ndarray) and pairs.dtype == _lib.PAIR_DTYPE):
            from delta_graph_slam_b200.loop_batch import make_pairs
            pairs = make_pairs(pairs)
        out = np.zeros(len(pairs), _lib.RESULT_DTYPE)
        last_target = None
        self.calls += 1
        for i, p in enumerate(pairs):
            t, s = int(p["target_id"]), int(p["source_id"])
            if t != last_target:
                self.reg.setInputTarget(self.clouds[t])
                last_target = t
            self.reg.setInputSource(self.clouds[s])
            self.reg.align(_lib.from_colmajor(p["guess"]))
            out[i]["transformation"] = _lib.colmajor(self.reg.getFinalTransformation())
            out[i]["converged"] = int(self.reg.hasConverged())
            out[i]["iterations"] = self.reg.getFinalNumIteration()
            info = self.reg.info()
            out[i]["score"], out[i]["evaluations"], out[i]["hits"] = info[0], int(info[1]), int(info[2])
            out[i]["fitness"] = self.reg.getFitnessScore(fitness_max_range) if with_fitness else np.finfo(np.float64).max
        return out


def rot_angle(Ra, Rb):
    """Rotation angle between two rotation matrices from the skew part (arccos of the trace cannot
    resolve angles below ~3e-4 rad from float32 entries)."""
    R = np.asarray(Ra, np.float64).T @ np.asarray(Rb, np.float64)
    w = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / 2.0
    return float(np.arctan2(np.linalg.norm(w), (np.trace(R) - 1.0) / 2.0))


def small_loop_scenario(oracle, n_targets=3, n_candidates=4, leaf=0.2, stride=2, seed=11):
    """A small loop batch ray-cast on the CPU: returns (clouds {id: (N,4)}, pairs PAIR_DTYPE, true_rel list)."""
    from delta_graph_slam_b200.loop_batch import make_pairs
    from delta_graph_slam_b200.synth.loop_scenario import loop_scenario
    sc = loop_scenario(oracle.synth_traj, n_targets=n_targets, n_candidates=n_candidates, spacing_frames=8, seed=seed)
    clouds = {}
    for cid, P, ns in sc["targets"] + sc["candidates"]:
        raw = oracle.synth_scan(P, noise_seed=ns)[::stride]
        clouds[cid] = oracle.voxelgrid(raw, leaf)["out"]
    pairs = make_pairs([(t, c, g) for t, c, g, _ in sc["pairs"]])
    return clouds, pairs, [rel for _, _, _, rel in sc["pairs"]]


def bits_equal(a, b):
    """Same shape and the same float32 bit patterns."""
    a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def transform_delta(Ta, Tb):
    """(max |dt| in metres, rotation angle in radians) between two 4x4 transforms."""
    Ta, Tb = np.asarray(Ta), np.asarray(Tb)
    return float(np.max(np.abs(Ta[:3, 3].astype(np.float64) - Tb[:3, 3].astype(np.float64)))), rot_angle(Ta[:3, :3], Tb[:3, :3])
