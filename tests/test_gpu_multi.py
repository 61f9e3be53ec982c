"""b200reg_batch_* — the loop-candidate batch over several GPUs from ONE process, NCCL all-gather inside the library
(what a C++ LoopDetector [REF include/hdl_graph_slam/loop_detector.hpp:119-173] would link against).
On a one-GPU box the same code path runs with a single NCCL rank; with more GPUs visible (gpurun --gpus N) every device
takes whole targets and the gathered records must match the single-handle batch."""
import io

import numpy as np
import pytest

from helpers import rot_angle, small_loop_scenario

pytestmark = pytest.mark.gpu
PARAMS = dict(registration_method="NDT_OMP", reg_resolution=1.0, reg_nn_search_method="DIRECT7")


@pytest.fixture(scope="module")
def scenario(oracle):
    clouds, pairs, rels = small_loop_scenario(oracle, n_targets=4, n_candidates=3)
    return dict(clouds=clouds, pairs=pairs)


def single_handle(eng, sc):
    ndt = eng.select_registration_method(PARAMS, out=io.StringIO())
    for k, v in sc["clouds"].items():
        ndt.cloudPut(k, v)
    return ndt.alignBatch(sc["pairs"])


def T_of(rec):
    return np.array(rec["transformation"], np.float32).reshape(4, 4).T


def test_one_device_through_the_multi_gpu_entry_is_the_single_handle_batch(scenario):
    import delta_graph_slam_b200 as eng
    from delta_graph_slam_b200.loop_batch import MultiGpuBatch
    base = single_handle(eng, scenario)
    mg = MultiGpuBatch([0], PARAMS)
    for k, v in scenario["clouds"].items():
        mg.cloudPut(k, v)
    got = mg.alignBatch(scenario["pairs"])
    assert np.array_equal(got.view(np.uint8), base.view(np.uint8)), "same device, same launch shape: bit-identical records"
    info = mg.info()
    assert info["n_devices"] == 1 and info["pairs_per_device"] == [len(scenario["pairs"])]
    assert info["uses_nccl"] and info["nccl_version"] >= 20000, "NCCL carries the gather (this image ships libnccl.so.2)"
    # a second run reuses the resident clouds; without fitness the slot reads DBL_MAX
    again = mg.alignBatch(scenario["pairs"], with_fitness=False)
    assert np.array_equal(again["transformation"], base["transformation"]) and np.all(again["fitness"] == np.finfo(np.float64).max)
    # argument errors: unknown cloud id, duplicate devices, a device that does not exist
    bad = scenario["pairs"][:1].copy()
    bad["source_id"] = 987654
    with pytest.raises(eng.B200RegError) as e:
        mg.alignBatch(bad)
    assert e.value.code == eng._lib.E_INVALID
    with pytest.raises(eng.B200RegError):
        MultiGpuBatch([0, 0], PARAMS)
    with pytest.raises(eng.B200RegError):
        MultiGpuBatch([97], PARAMS)
    mg.cloudDrop(int(scenario["pairs"][0]["source_id"]))
    with pytest.raises(eng.B200RegError):
        mg.alignBatch(scenario["pairs"][:1])


def test_all_visible_devices_shard_whole_targets_and_gather(scenario):
    import torch
    import delta_graph_slam_b200 as eng
    from delta_graph_slam_b200.loop_batch import MultiGpuBatch, shard_by_target
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("one GPU visible: the multi-device gather runs under gpurun --gpus 2 (tools/multi_gpu_check.sh)")
    base = single_handle(eng, scenario)
    devices = list(range(min(n, 4)))
    mg = MultiGpuBatch(devices, PARAMS)
    for k, v in scenario["clouds"].items():
        mg.cloudPut(k, v)
    got = mg.alignBatch(scenario["pairs"])
    info = mg.info()
    shards = shard_by_target(scenario["pairs"]["target_id"], len(devices))
    assert info["pairs_per_device"] == [len(s) for s in shards] and info["uses_nccl"]
    # every device's share is the single-handle batch of exactly those pairs (same launch shape: same CTAs per
    # registration, same summation order), so the gathered records are that batch's records bit for bit
    ndt = eng.select_registration_method(PARAMS, out=io.StringIO())
    for k, v in scenario["clouds"].items():
        ndt.cloudPut(k, v)
    for shard in shards:
        if len(shard) == 0:
            continue
        idx = np.asarray(shard)
        alone = ndt.alignBatch(scenario["pairs"][idx])
        assert np.array_equal(got[idx].view(np.uint8), alone.view(np.uint8)), "a device's share = the single-handle batch of the same pairs"
    # against the whole batch on one handle: fewer pairs per launch -> more CTAs per registration -> another (fixed)
    # summation order.  Same convergence; poses within the bound the reference's own sensitivity to a last-bit change
    # of the sums allows (tests/test_gpu_odometry_sequence.py: up to ~1e-3 m on sensitive pairs), most far inside 1e-4
    tight = 0
    for a, b in zip(got, base):
        assert a["converged"] == b["converged"]
        dt, dr = np.max(np.abs(T_of(a)[:3, 3] - T_of(b)[:3, 3])), rot_angle(T_of(a)[:3, :3], T_of(b)[:3, :3])
        assert dt < 5e-3 and dr < 1e-3
        tight += int(dt < 1e-4 and dr < 1e-4 and a["iterations"] == b["iterations"])
    assert tight >= (3 * len(base)) // 4
    again = mg.alignBatch(scenario["pairs"])
    assert np.array_equal(again.view(np.uint8), got.view(np.uint8)), "a rerun on the same devices is bit-identical"
