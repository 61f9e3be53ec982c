"""CPU tests of the loop-closure host logic: the LoopDetector mirror
[REF include/hdl_graph_slam/loop_detector.hpp:59-173], transform2Dto3D
[REF src/hdl_graph_slam/ros_utils.cpp:105-126], target sharding and the world_size-2 result gather
(gloo), with the oracle standing in for the engine."""
import io
import os
import socket
import sys

import numpy as np
import pytest

from helpers import OracleBatchEngine, small_loop_scenario
from delta_graph_slam_b200 import loop_batch
from delta_graph_slam_b200.loop_detector import KeyFrame, LoopDetector, candidate_guess, isometry2d, select_best, transform2Dto3D

DBL_MAX = np.finfo(np.float64).max


def test_transform2Dto3D():
    yaw = 0.3
    t2 = np.array([[np.cos(yaw), -np.sin(yaw), 1.5], [np.sin(yaw), np.cos(yaw), -2.0], [0, 0, 1]], np.float32)
    T = transform2Dto3D(t2)
    assert T.dtype == np.float32
    assert np.allclose(T[:2, :2], t2[:2, :2], atol=1e-7) and T[2, 2] == 1 and np.all(T[2, :2] == 0) and np.all(T[:2, 2] == 0)
    assert np.allclose(T[:3, 3], [1.5, -2.0, 0.0]) and np.all(T[3] == [0, 0, 0, 1])


def test_candidate_guess_is_relative_2d_pose():
    new = KeyFrame(1, None, isometry2d(10.0, 2.0, 0.5), 100.0)
    cand = KeyFrame(2, None, isometry2d(11.0, 2.5, 0.7), 10.0)
    g = candidate_guess(new, cand)
    rel = np.linalg.inv(new.estimate()) @ cand.estimate()
    assert np.allclose(g[:2, 3], rel[:2, 2], atol=1e-6)
    assert np.isclose(np.arctan2(g[1, 0], g[0, 0]), 0.2, atol=1e-6)


def test_find_candidates_rules():
    ld = LoopDetector(dict(distance_thresh=5.0, accum_distance_thresh=8.0, min_edge_interval=5.0), registration=object(), out=io.StringIO())
    kfs = [KeyFrame(i, None, isometry2d(float(i), 0.0, 0.0), accum_distance=float(i)) for i in range(30)]
    new = KeyFrame(99, None, isometry2d(3.0, 0.0, 0.0), accum_distance=29.5)
    ids = [k.id for k in ld.find_candidates(kfs, new)]
    # travelled distance >= 8 behind the new keyframe (29.5 - k >= 8 -> k <= 21) AND within 5 m in the plane (k <= 8)
    assert ids == list(range(0, 9))
    ld.last_edge_accum_distance = 26.0  # too close to the last loop edge
    assert ld.find_candidates(kfs, new) == []
    assert ld.matching([], new) is None


def test_select_best_follows_the_reference_loop():
    c = ["a", "b", "c", "d"]
    Ts = [np.eye(4) * k for k in range(4)]
    # non-converged candidates are skipped even with the best score; the LAST of equal scores wins
    score, best, T = select_best(c, [True, False, True, True], [0.3, 0.1, 0.2, 0.2], Ts)
    assert (score, best) == (0.2, "d") and T is Ts[3]
    score, best, T = select_best(c, [False] * 4, [0.3, 0.1, 0.2, 0.2], Ts)
    assert score == DBL_MAX and best is None


def test_shard_by_target_properties():
    rng = np.random.default_rng(0)
    tids = np.repeat(rng.permutation(40), rng.integers(1, 20, 40))
    for world in (1, 2, 4, 8):
        shards = loop_batch.shard_by_target(tids, world)
        allidx = np.concatenate(shards)
        assert sorted(allidx.tolist()) == list(range(len(tids))), "a partition of the pairs"
        owners = {}
        for r, s in enumerate(shards):
            for t in set(tids[s].tolist()):
                assert owners.setdefault(t, r) == r, "a target lives on exactly one rank"
        loads = [len(s) for s in shards]
        assert max(loads) - min(loads) <= 19, "balanced to within one target"


@pytest.fixture(scope="module")
def scenario(oracle):
    clouds, pairs, rels = small_loop_scenario(oracle, n_targets=2, n_candidates=3, leaf=0.4, stride=4)
    return clouds, pairs


def test_loop_detector_on_plain_registration_surface(oracle, scenario):
    """matching() through setInputTarget/Source, align, getFitnessScore equals the batch surface."""
    clouds, pairs = scenario
    old = [KeyFrame(int(p["source_id"]), clouds[int(p["source_id"])], isometry2d(0.4 * k, 0.0, 0.0), 1.0) for k, p in enumerate(pairs[:3])]
    new = KeyFrame(0, clouds[0], isometry2d(0.5, 0.0, 0.0), 50.0)
    params = dict(distance_thresh=5.0, accum_distance_thresh=8.0, min_edge_interval=5.0, fitness_score_thresh=10.0)
    serial = LoopDetector(params, registration=oracle.Registration(oracle.NDT, resolution=1.0, nn_search=2, trans_eps=0.01, max_iter=64), out=io.StringIO())
    batch = LoopDetector(params, registration=OracleBatchEngine(oracle), out=io.StringIO())
    la = serial.detect(old, [new])
    lb = batch.detect(old, [new])
    assert len(la) == len(lb) == 1
    assert la[0].key2.id == lb[0].key2.id and la[0].score == lb[0].score
    assert np.array_equal(la[0].relative_pose, lb[0].relative_pose)
    assert serial.last_edge_accum_distance == 50.0
    # a second new keyframe right behind the loop edge is suppressed by min_edge_interval
    assert serial.detect(old, [KeyFrame(1, clouds[1], isometry2d(0.6, 0, 0), 52.0)]) == []


def test_batch_path_is_ndt_only_and_the_keyframe_cache_is_bounded(oracle, scenario):
    """A registration object whose batch call cannot take its method (FAST_GICP, the launch file's loop-detector
    method [REF launch/delta_graph_slam.launch:95]) runs the reference's serial loop; cached keyframes beyond the cap
    are dropped least-recently-used first, never one of the running batch."""
    from delta_graph_slam_b200 import _lib
    clouds, pairs = scenario
    log = []

    class GicpLike(OracleBatchEngine):
        method = _lib.METHOD_GICP

        def alignBatch(self, *a, **k):
            raise AssertionError("the batch path must not be taken for a FAST_GICP handle")

        def __getattr__(self, name):  # the pcl::Registration surface, served by the oracle's NDT object
            return getattr(self.reg, name)

    old = [KeyFrame(int(p["source_id"]), clouds[int(p["source_id"])], isometry2d(0.4 * k, 0.0, 0.0), 1.0) for k, p in enumerate(pairs[:3])]
    new = KeyFrame(0, clouds[0], isometry2d(0.5, 0.0, 0.0), 50.0)
    params = dict(distance_thresh=5.0, accum_distance_thresh=8.0, min_edge_interval=5.0, fitness_score_thresh=10.0)
    serial = LoopDetector(params, registration=oracle.Registration(oracle.NDT, resolution=1.0, nn_search=2, trans_eps=0.01, max_iter=64), out=io.StringIO())
    gicp_like = LoopDetector(params, registration=GicpLike(oracle), out=io.StringIO())
    assert not gicp_like.batch_capable()
    la, lb = serial.detect(old, [new]), gicp_like.detect(old, [new])
    assert len(la) == len(lb) == 1 and la[0].key2.id == lb[0].key2.id and la[0].score == lb[0].score

    class Counting(OracleBatchEngine):
        def cloudPut(self, cid, cloud):
            log.append(("put", int(cid)))
            super().cloudPut(cid, cloud)

        def cloudDrop(self, cid):
            log.append(("drop", int(cid)))
            del self.clouds[int(cid)]

    ld = LoopDetector(dict(params, b200_max_cached_keyframes=3), registration=Counting(oracle), out=io.StringIO())
    ld.register_candidates(old[:2], new)              # 3 clouds cached: at the cap
    assert [e for e in log if e[0] == "drop"] == []
    ld.register_candidates(old[2:3], new)             # one more: the least recently used candidate goes
    assert [e for e in log if e[0] == "drop"] == [("drop", old[0].id)]
    assert set(ld._cached) == {new.id, old[1].id, old[2].id} and set(ld.registration.clouds) == set(ld._cached)
    ld.register_candidates(old[:3], new)              # a batch larger than the cap keeps all of its own clouds
    assert set(ld._cached) == {new.id, old[0].id, old[1].id, old[2].id}


def test_serial_loop_of_a_gicp_handle_takes_its_clouds_from_the_keyframe_cache(oracle, scenario):
    """An engine handle that has no batch path for its method but offers the cached setters (b200reg_set_source_cached /
    _set_target_cached: a FAST_GICP handle) is driven through the reference's sequence [REF include/hdl_graph_slam/
    loop_detector.hpp:124-156] with keyframe ids instead of clouds: every keyframe is put once, the new keyframe becomes
    the target once, each candidate the source once, and the outcome is the plain serial loop's."""
    from delta_graph_slam_b200 import _lib
    clouds, pairs = scenario
    log = []

    class CachedGicpLike(OracleBatchEngine):
        method = _lib.METHOD_GICP

        def alignBatch(self, *a, **k):
            raise AssertionError("the batch path must not be taken for a FAST_GICP handle")

        def cloudPut(self, cid, cloud):
            log.append(("put", int(cid)))
            super().cloudPut(cid, cloud)

        def cloudDrop(self, cid):
            log.append(("drop", int(cid)))
            del self.clouds[int(cid)]

        def setInputTargetCached(self, cid):
            log.append(("target", int(cid)))
            self.reg.setInputTarget(self.clouds[int(cid)])

        def setInputSourceCached(self, cid):
            log.append(("source", int(cid)))
            self.reg.setInputSource(self.clouds[int(cid)])

        def setInputTarget(self, cloud):
            raise AssertionError("clouds are not handed over when the handle has a keyframe cache")

        setInputSource = setInputTarget

        def __getattr__(self, name):
            return getattr(self.reg, name)

    old = [KeyFrame(int(p["source_id"]), clouds[int(p["source_id"])], isometry2d(0.4 * k, 0.0, 0.0), 1.0) for k, p in enumerate(pairs[:3])]
    new = KeyFrame(0, clouds[0], isometry2d(0.5, 0.0, 0.0), 50.0)
    params = dict(distance_thresh=5.0, accum_distance_thresh=8.0, min_edge_interval=5.0, fitness_score_thresh=10.0)
    serial = LoopDetector(params, registration=oracle.Registration(oracle.NDT, resolution=1.0, nn_search=2, trans_eps=0.01, max_iter=64), out=io.StringIO())
    cached = LoopDetector(dict(params, b200_max_cached_keyframes=3), registration=CachedGicpLike(oracle), out=io.StringIO())
    c0, s0, T0 = serial.register_candidates(old[:2], new)
    c1, s1, T1 = cached.register_candidates(old[:2], new)
    assert c0 == c1 and s0 == s1 and all(np.array_equal(a, b) for a, b in zip(T0, T1))
    assert log == [("put", new.id), ("put", old[0].id), ("put", old[1].id), ("target", new.id), ("source", old[0].id), ("source", old[1].id)]
    # the second detection puts only the keyframe it has not seen and evicts the least recently used one above the cap
    log.clear()
    cached.register_candidates(old[2:3], new)
    assert log == [("put", old[2].id), ("drop", old[0].id), ("target", new.id), ("source", old[2].id)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from oracle import oracle_py as oracle
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    clouds, pairs, _ = small_loop_scenario(oracle, n_targets=2, n_candidates=3, leaf=0.4, stride=4)
    shards = loop_batch.shard_by_target(pairs["target_id"], world)
    eng = OracleBatchEngine(oracle)
    for cid in loop_batch.needed_clouds(pairs, shards[rank]):  # a rank holds only its own share of the clouds
        eng.cloudPut(cid, clouds[cid])
    full, _ = loop_batch.align_batch_sharded(eng, pairs, rank=rank, world_size=world)
    q.put((rank, full.tobytes(), len(eng.clouds)))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_batch_world_size_2_gloo(oracle, scenario):
    """Two processes, gloo: each registers whole targets, one all-gather of the result records;
    both ranks end with the full array, identical to the single-process run."""
    import torch.multiprocessing as mp
    clouds, pairs = scenario
    ref_eng = OracleBatchEngine(oracle)
    for k, v in clouds.items():
        ref_eng.cloudPut(k, v)
    ref, shards1 = loop_batch.align_batch_sharded(ref_eng, pairs, rank=0, world_size=1)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, blob, n_clouds in got:
        assert blob == ref.tobytes(), f"rank {rank} holds the full, identical result array"
        assert n_clouds < len(clouds), "each rank uploads only the clouds of its own targets"
