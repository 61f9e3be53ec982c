"""Host-side mirror of hdl_graph_slam::LoopDetector [REF include/hdl_graph_slam/loop_detector.hpp:33-187].

Candidate selection (`find_candidates`, :83-111), the per-candidate initial guess
(`transform2Dto3D` of the 2-D relative pose, :139-143 and [REF src/hdl_graph_slam/ros_utils.cpp:105-126]),
the arg-min over fitness scores (:148-156) and the threshold (:163-170) are restated here as host
logic.  The registrations themselves — the serial `for candidate` loop of :137-156 — go to the
engine as ONE batch (`Registration.alignBatch`): every candidate is aligned against the new
keyframe and scored with getFitnessScore(fitness_score_max_range) on the GPU, the target's NDT
grid and exact-NN structure being built once (the hoisted setInputTarget of :124).

Objects that only expose the pcl::Registration calls (no alignBatch) are driven through the
reference's own serial sequence instead, so the same class runs on the CPU oracle in tests.
"""
import sys

import numpy as np

from collections import OrderedDict

from . import _lib
from .registration import DBL_MAX, select_registration_method


def transform2Dto3D(trans2D):
    """[REF src/hdl_graph_slam/ros_utils.cpp:105-126]: yaw from the 2x2 block (Rotation2Df::angle =
    atan2(m10, m00)), roll = pitch = 0, translation (x, y, 0); float throughout."""
    t = np.asarray(trans2D, np.float32)
    ang = np.arctan2(t[1, 0], t[0, 0]).astype(np.float32)
    c, s = np.cos(ang).astype(np.float32), np.sin(ang).astype(np.float32)
    out = np.eye(4, dtype=np.float32)
    out[0, 0], out[0, 1], out[1, 0], out[1, 1] = c, -s, s, c
    out[0, 3], out[1, 3] = t[0, 2], t[1, 2]
    return out


def isometry2d(x, y, yaw):
    c, s = np.cos(yaw), np.sin(yaw)
    return np.array([[c, -s, x], [s, c, y], [0.0, 0.0, 1.0]], np.float64)


class KeyFrame:
    """The fields of hdl_graph_slam::KeyFrame the loop detector reads [REF include/hdl_graph_slam/keyframe.hpp:25-59]:
    `cloud`, `accum_distance` and `estimate()` (the VertexSE2 estimate, a 2-D isometry)."""

    def __init__(self, kf_id, cloud, estimate2d, accum_distance):
        self.id = int(kf_id)
        self.cloud = cloud
        self._estimate = np.asarray(estimate2d, np.float64)
        self.accum_distance = float(accum_distance)

    def estimate(self):
        return self._estimate


class Loop:
    def __init__(self, key1, key2, relative_pose, score):
        self.key1, self.key2, self.relative_pose, self.score = key1, key2, relative_pose, score


def candidate_guess(new_keyframe, candidate):
    """guess = transform2Dto3D((new.estimate()^-1 * candidate.estimate()).cast<float>()) [:139-143]."""
    g2 = np.linalg.inv(new_keyframe.estimate()) @ candidate.estimate()
    return transform2Dto3D(g2.astype(np.float32))


def select_best(candidates, converged, scores, transforms):
    """The running arg-min of :149-156: a candidate replaces the best unless it did not converge
    or scores strictly worse (so the LAST of equal scores wins, as in the reference)."""
    best_score, best, rel = DBL_MAX, None, None
    for c, ok, sc, T in zip(candidates, converged, scores, transforms):
        if (not ok) or sc > best_score:
            continue
        best_score, best, rel = sc, c, T
    return best_score, best, rel


class LoopDetector:
    def __init__(self, params=None, device=0, out=sys.stdout, registration=None):
        p = dict(params or {})
        self.distance_thresh = p.get("distance_thresh", 5.0)
        self.accum_distance_thresh = p.get("accum_distance_thresh", 8.0)
        self.distance_from_last_edge_thresh = p.get("min_edge_interval", 5.0)
        self.fitness_score_max_range = p.get("fitness_score_max_range", DBL_MAX)
        self.fitness_score_thresh = p.get("fitness_score_thresh", 0.5)
        self.registration = registration if registration is not None else select_registration_method(p, device=device, out=out)
        self.last_edge_accum_distance = 0.0
        self.out = out
        # keyframe clouds resident on the device, least recently used first; beyond `b200_max_cached_keyframes`
        # (the mirror's own key; ~0.8 MB per HDL-64 keyframe, 4 MB once it has served as a target) the oldest are dropped
        self.max_cached_keyframes = int(p.get("b200_max_cached_keyframes", 16384))
        self._cached = OrderedDict()

    def get_distance_thresh(self):
        return self.distance_thresh

    def detect(self, keyframes, new_keyframes):
        loops = []
        for new_keyframe in new_keyframes:
            candidates = self.find_candidates(keyframes, new_keyframe)
            loop = self.matching(candidates, new_keyframe)
            if loop is not None:
                loops.append(loop)
        return loops

    def find_candidates(self, keyframes, new_keyframe):
        if new_keyframe.accum_distance - self.last_edge_accum_distance < self.distance_from_last_edge_thresh:
            return []
        candidates = []
        pos2 = new_keyframe.estimate()[:2, 2]
        for k in keyframes:
            if new_keyframe.accum_distance - k.accum_distance < self.accum_distance_thresh:
                continue
            if float(np.linalg.norm(k.estimate()[:2, 2] - pos2)) > self.distance_thresh:
                continue
            candidates.append(k)
        return candidates

    # ---- the registrations
    def _ensure_cached(self, kf):
        if kf.id in self._cached:
            self._cached.move_to_end(kf.id)
            return
        self.registration.cloudPut(kf.id, kf.cloud)
        self._cached[kf.id] = True

    def _evict(self, keep):
        """Drop least-recently-used keyframes above the cap (never one of the batch about to run)."""
        while len(self._cached) > self.max_cached_keyframes:
            victim = next((k for k in self._cached if k not in keep), None)
            if victim is None:
                return
            del self._cached[victim]
            self.registration.cloudDrop(victim)

    def batch_capable(self):
        """The engine's batch path registers NDT pairs (b200reg_align_batch); any other registration object — FAST_GICP,
        the launch file's choice [REF launch/delta_graph_slam.launch:95], or a plain pcl::Registration surface — is
        driven through the reference's own serial loop."""
        reg = self.registration
        return hasattr(reg, "alignBatch") and getattr(reg, "method", _lib.METHOD_NDT) in getattr(reg, "batch_methods", (_lib.METHOD_NDT,))

    def register_candidates(self, candidates, new_keyframe):
        """(converged[], scores[], transforms[]) of every candidate against the new keyframe."""
        reg = self.registration
        guesses = [candidate_guess(new_keyframe, c) for c in candidates]
        if self.batch_capable():
            self._ensure_cached(new_keyframe)
            for c in candidates:
                self._ensure_cached(c)
            self._evict({new_keyframe.id} | {c.id for c in candidates})
            res = reg.alignBatch([(new_keyframe.id, c.id, g) for c, g in zip(candidates, guesses)], with_fitness=True, fitness_max_range=self.fitness_score_max_range)
            return ([bool(r["converged"]) for r in res], [float(r["fitness"]) for r in res],
                    [np.array(r["transformation"], np.float32).reshape(4, 4).T.copy() for r in res])
        # the reference's serial sequence [REF include/hdl_graph_slam/loop_detector.hpp:124-156].  On an engine handle (FAST_GICP:
        # the launch file's loop detector) the clouds come from the keyframe cache: no upload per pair, and a keyframe's
        # covariances are computed once in its life (b200reg_set_source_cached); on a plain pcl::Registration surface the
        # clouds are handed over as the reference does
        cached = hasattr(reg, "setInputSourceCached") and hasattr(reg, "cloudPut")
        if cached:
            self._ensure_cached(new_keyframe)
            for c in candidates:
                self._ensure_cached(c)
            self._evict({new_keyframe.id} | {c.id for c in candidates})
            reg.setInputTargetCached(new_keyframe.id)
        else:
            reg.setInputTarget(new_keyframe.cloud)
        conv, scores, Ts = [], [], []
        for c, g in zip(candidates, guesses):
            if cached:
                reg.setInputSourceCached(c.id)
            else:
                reg.setInputSource(c.cloud)
            reg.align(g)
            scores.append(reg.getFitnessScore(self.fitness_score_max_range))
            conv.append(reg.hasConverged())
            Ts.append(reg.getFinalTransformation())
        return conv, scores, Ts

    def matching(self, candidate_keyframes, new_keyframe):
        if not candidate_keyframes:
            return None
        print("\n--- loop detection ---", file=self.out)
        print(f"num_candidates: {len(candidate_keyframes)}", file=self.out)
        conv, scores, Ts = self.register_candidates(candidate_keyframes, new_keyframe)
        best_score, best_matched, relative_pose = select_best(candidate_keyframes, conv, scores, Ts)
        print(f"best_score: {best_score:.3f}", file=self.out)
        if best_score > self.fitness_score_thresh:
            print("loop not found...", file=self.out)
            return None
        print("loop found!!", file=self.out)
        self.last_edge_accum_distance = new_keyframe.accum_distance
        return Loop(new_keyframe, best_matched, relative_pose, best_score)
