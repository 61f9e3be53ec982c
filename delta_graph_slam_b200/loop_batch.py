"""Loop-candidate batches across the GPUs of one box.

The pairs of a batch are independent registrations, so the path shards with no data-path
collective: WHOLE TARGETS (a new keyframe with all of its candidates) are dealt to ranks, so each
target's NDT grid and exact-NN structure is built on exactly one GPU — the multi-GPU form of the
setInputTarget hoisted out of the candidate loop [REF include/hdl_graph_slam/loop_detector.hpp:124].
Each rank (one process per GPU) registers its share with `Registration.alignBatch`; the only
exchange is one all-gather of the fixed-size result records (transform, fitness, converged,
iterations — 104 bytes per pair) over NCCL, after which every rank holds the full result array and
the host does the per-target arg-min exactly as the reference does (:149-165).
"""
import numpy as np

from . import _lib

RESULT_DTYPE = _lib.RESULT_DTYPE
PAIR_DTYPE = _lib.PAIR_DTYPE


def make_pairs(triples):
    """[(target_id, source_id, guess 4x4 or None), ...] -> PAIR_DTYPE array."""
    triples = list(triples)
    arr = np.zeros(len(triples), PAIR_DTYPE)
    for i, (t, s, g) in enumerate(triples):
        arr[i] = (int(t), int(s), _lib.colmajor(np.eye(4) if g is None else g))
    return arr


def shard_by_target(target_ids, world_size):
    """Deal whole targets to ranks.  Targets are taken in order of first appearance and each goes
    to the rank with the fewest pairs so far (ties -> lowest rank): deterministic, balanced to
    within one target, and identical on every rank.  Returns a list of index arrays (one per rank,
    ascending pair indices)."""
    target_ids = np.asarray(target_ids)
    order, first = [], {}
    for i, t in enumerate(target_ids.tolist()):
        if t not in first:
            first[t] = len(order)
            order.append([])
        order[first[t]].append(i)
    load = [0] * world_size
    shards = [[] for _ in range(world_size)]
    for idx in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        shards[r].extend(idx)
        load[r] += len(idx)
    return [np.array(sorted(s), dtype=np.int64) for s in shards]


def gather_results(local_results, shards, rank, world_size, device=None, group=None):
    """All-gather of the per-rank result records; returns the full RESULT_DTYPE array in the
    original pair order on every rank.  `device` = torch device of the collective buffers
    (a CUDA device for the NCCL backend, None / "cpu" for gloo)."""
    n_total = int(sum(len(s) for s in shards))
    out = np.zeros(n_total, RESULT_DTYPE)
    if world_size == 1:
        out[shards[0]] = local_results
        return out
    import torch
    import torch.distributed as dist
    rec = RESULT_DTYPE.itemsize
    slot = max(len(s) for s in shards)
    send = np.zeros(slot * rec, np.uint8)
    mine = np.ascontiguousarray(local_results).view(np.uint8).reshape(-1)
    send[: mine.size] = mine
    dev = torch.device(device) if device is not None else torch.device("cpu")
    t_send = torch.from_numpy(send).to(dev)
    t_recv = torch.empty(world_size * slot * rec, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(t_recv, t_send, group=group)
    recv = t_recv.cpu().numpy().reshape(world_size, slot * rec)
    for r in range(world_size):
        k = len(shards[r])
        if k:
            out[shards[r]] = recv[r, : k * rec].view(RESULT_DTYPE)
    return out


def align_batch_sharded(registration, pairs, rank=0, world_size=1, with_fitness=True, fitness_max_range=float(np.finfo(np.float64).max), device=None, group=None):
    """Register this rank's share of `pairs` (PAIR_DTYPE array, identical on every rank; the clouds
    of the share must be cached on this rank's engine) and gather everyone's results."""
    pairs = np.asarray(pairs)
    shards = shard_by_target(pairs["target_id"], world_size)
    mine = pairs[shards[rank]]
    local = registration.alignBatch(mine, with_fitness=with_fitness, fitness_max_range=fitness_max_range)
    return gather_results(local, shards, rank, world_size, device=device, group=group), shards


def needed_clouds(pairs, shard):
    """ids of the clouds a rank has to hold for its share (targets and sources)."""
    sub = np.asarray(pairs)[shard]
    return sorted(set(sub["target_id"].tolist()) | set(sub["source_id"].tolist()))
