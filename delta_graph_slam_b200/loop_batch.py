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


class MultiGpuBatch:
    """b200reg_batch_* (include/b200reg.h): the loop-candidate batch over several GPUs from ONE process — what a C++
    LoopDetector links against (the multi-process form above is what bench.py runs under torchrun).  Same sharding rule
    as shard_by_target; one NCCL all-gather of the result records inside the library."""

    def __init__(self, devices, params=None):
        import ctypes as C
        # one NCCL per process: torch (device plumbing of this package) links its bundled libnccl.so.2; imported first,
        # the library's dlopen("libnccl.so.2") gets that same object.  The other order would put the system NCCL under
        # torch's soname and a later `import torch` would fail on its missing symbols.
        import torch  # noqa: F401
        p = dict(params or {})
        L = _lib.load()
        cfg = _lib.Config()
        L.b200reg_default_config(_lib.METHOD_NDT, C.byref(cfg))
        cfg.resolution = float(p.get("reg_resolution", 0.5))
        cfg.transformation_epsilon = float(p.get("reg_transformation_epsilon", 0.01))
        cfg.maximum_iterations = int(p.get("reg_maximum_iterations", 64))
        cfg.nn_search = {"KDTREE": _lib.KDTREE, "DIRECT1": _lib.DIRECT1}.get(p.get("reg_nn_search_method", "DIRECT7"), _lib.DIRECT7)
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        self._h = C.c_void_p()
        rc = L.b200reg_batch_create(C.byref(cfg), devs, len(devices), C.byref(self._h))
        if rc != _lib.OK:
            self._h = None
            raise _lib.B200RegError(rc, "b200reg_batch_create failed (a listed device is not a usable sm_100 GPU, or NCCL is not loadable for more than one device)")
        self.devices = list(devices)
        self._keep = {}

    def close(self):
        if getattr(self, "_h", None):
            _lib.load().b200reg_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != _lib.OK:
            raise _lib.B200RegError(rc, _lib.load().b200reg_batch_last_error(self._h).decode())

    def cloudPut(self, cloud_id, cloud):
        c = _lib.as_cloud(cloud)
        self._keep[int(cloud_id)] = c  # the library keeps the pointer, so the array must stay alive
        self._ck(_lib.load().b200reg_batch_cloud_put(self._h, int(cloud_id), c.ctypes.data if len(c) else None, len(c), 16))

    def cloudDrop(self, cloud_id):
        self._ck(_lib.load().b200reg_batch_cloud_drop(self._h, int(cloud_id)))
        self._keep.pop(int(cloud_id), None)

    def alignBatch(self, pairs, with_fitness=True, fitness_max_range=float(np.finfo(np.float64).max)):
        arr = np.ascontiguousarray(pairs) if isinstance(pairs, np.ndarray) and pairs.dtype == PAIR_DTYPE else make_pairs(pairs)
        out = np.zeros(len(arr), RESULT_DTYPE)
        if len(arr):
            self._ck(_lib.load().b200reg_batch_run(self._h, arr.ctypes.data, len(arr), int(with_fitness), float(fitness_max_range), out.ctypes.data))
        return out

    def info(self):
        import ctypes as C
        n, nc, ver, ms = C.c_int(), C.c_int(), C.c_int(), C.c_double()
        per = (C.c_int * len(self.devices))()
        self._ck(_lib.load().b200reg_batch_get_info(self._h, C.byref(n), C.byref(nc), C.byref(ver), per, C.byref(ms)))
        return dict(n_devices=n.value, uses_nccl=bool(nc.value), nccl_version=ver.value, pairs_per_device=list(per), gather_ms=ms.value)
