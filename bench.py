#!/usr/bin/env python
"""bench.py — throughput of the B200 scan-registration engine on the reference's two callers.

Headline workload (BASELINE.json configs[1], N = 1): 1000 consecutive synthetic HDL-64-shaped scans
(`street_v1` scene, `kitti_like` trajectory, ~1.3e5 points each) through
PrefilteringNodelet::downsample (VoxelGrid 0.1 m) and ScanMatchingOdometryNodelet::matching
(NDT_OMP-equivalent: DIRECT7, resolution 1.0, epsilon 0.01, 64 iterations; keyframe_delta
1.0 / 1.0 / 10000 as in launch/delta_graph_slam.launch).  One "step" is one pass over the sequence.

  value : registrations/s with the raw scans already resident in HBM (device pointers through the
          C ABI); `e2e`: the same pipeline through the host-buffer calls the reference's nodelets
          would make (H2D of every raw scan, D2H of every filtered cloud, H2D of the source,
          D2H of the result) — both timed with CUDA events on the engine's stream.
  roofline : k_ndt_align, algorithmic bytes (16 B per source point per pass + 48 B per
          (point, voxel) hit) / its CUDA-event duration, against the measured HBM peak.
  cpu_baseline : the oracle restatement of ndt_omp + pcl::VoxelGrid on the host cores, on the first
          frames of the same sequence.
  loop_batch : BASELINE.json configs[3] in the same run — 256 new keyframes x 16 candidates = 4096
          NDT + getFitnessScore pairs (LoopDetector::matching), whole targets sharded over the N
          ranks, one NCCL all-gather of the result records; pairs/s = 4096 / max-over-ranks time
          (strong scaling).  `--workload loop` makes this the headline line instead.
  --impl reference : the CPU path as its own arm (the reference's libraries cannot be built here,
          so it is the oracle restatement with OpenMP on all host cores).

With N > 1 (torchrun) every rank runs an independent odometry sequence on its own GPU ("replicas
only": frame k's guess is frame k-1's result); times are the max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ODOM_PARAMS = dict(  # launch/delta_graph_slam.launch:50-69 (NDT_OMP instead of the launch's FAST_GICP)
    keyframe_delta_trans=1.0, keyframe_delta_angle=1.0, keyframe_delta_time=10000.0, downsample_method="NONE",
    registration_method="NDT_OMP", reg_resolution=1.0, reg_nn_search_method="DIRECT7", reg_transformation_epsilon=0.01, reg_maximum_iterations=64,
)
PREFILTER_PARAMS = dict(downsample_method="VOXELGRID", downsample_resolution=0.1)
GICP_ODOM_PARAMS = dict(  # BASELINE configs[2]: FAST_GICP with the factory's code defaults [REF src/hdl_graph_slam/registrations.cpp:27-36]
    keyframe_delta_trans=1.0, keyframe_delta_angle=1.0, keyframe_delta_time=10000.0, downsample_method="NONE",
    registration_method="FAST_GICP", reg_transformation_epsilon=0.01, reg_maximum_iterations=64, reg_max_correspondence_distance=2.5, reg_correspondence_randomness=20,
)
LOOP_PARAMS = dict(registration_method="NDT_OMP", reg_resolution=1.0, reg_nn_search_method="DIRECT7", reg_transformation_epsilon=0.01, reg_maximum_iterations=64)
DEVNULL = open(os.devnull, "w")
DBL_MAX = float(np.finfo(np.float64).max)


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic(label):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture
    (tools/ncu_traffic.py -> profiles/*_ncu_traffic.json); None when no capture is on file."""
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_ncu_traffic.json")), reverse=True):
        try:
            with open(path) as f:
                t = json.load(f).get(label)
            if t:
                return float(t["dram_bytes_per_launch"])
        except Exception:
            pass
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    FIELDS = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.samples = []
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=float(self.samples[0][1]), reasons=reasons, samples=len(sm))


def run_sequence(pre, odo, clouds, out_buf=None):
    """Prefilter + matching over a list of clouds; returns the poses."""
    poses = []
    for k, cloud in enumerate(clouds):
        filtered = pre.downsample(cloud, out=out_buf) if out_buf is not None else pre.downsample(cloud)
        poses.append(odo.matching(0.1 * k, filtered))
    return poses


class OraclePrefilter:
    def __init__(self, oracle):
        self.oracle = oracle

    def downsample(self, cloud):
        return self.oracle.voxelgrid(cloud, PREFILTER_PARAMS["downsample_resolution"], is_dense=False)["out"]


def oracle_odometry(oracle, threads=0, params=None):
    from delta_graph_slam_b200.odometry import ScanMatchingOdometry
    params = params or ODOM_PARAMS
    if params["registration_method"] == "FAST_GICP":
        reg = oracle.Registration(oracle.GICP, trans_eps=params["reg_transformation_epsilon"], max_iter=params["reg_maximum_iterations"],
                                  max_corr_dist=params["reg_max_correspondence_distance"], k_corr=params["reg_correspondence_randomness"], num_threads=threads)
    else:
        reg = oracle.Registration(oracle.NDT, resolution=params["reg_resolution"], nn_search=oracle.DIRECT7, trans_eps=params["reg_transformation_epsilon"],
                                  max_iter=params["reg_maximum_iterations"], num_threads=threads)
    return OraclePrefilter(oracle), ScanMatchingOdometry(params, registration=reg, out=DEVNULL)


def time_oracle_odometry(host_clouds, frames, params=None):
    from oracle import oracle_py as oracle
    pre, odo = oracle_odometry(oracle, params=params)
    t0 = time.perf_counter()
    run_sequence(pre, odo, host_clouds[:frames])
    dt = time.perf_counter() - t0
    return (frames - 1) / dt, oracle.lib().orc_max_threads(), dt


def oracle_loop_pairs(oracle, clouds, pairs):
    """The reference's serial candidate loop on the CPU: setInputTarget once per target, then per
    candidate setInputSource + align + getFitnessScore.  Returns seconds."""
    reg = oracle.Registration(oracle.NDT, resolution=LOOP_PARAMS["reg_resolution"], nn_search=oracle.DIRECT7, trans_eps=LOOP_PARAMS["reg_transformation_epsilon"],
                              max_iter=LOOP_PARAMS["reg_maximum_iterations"])
    t0 = time.perf_counter()
    last = None
    for p in pairs:
        if int(p["target_id"]) != last:
            reg.setInputTarget(clouds[int(p["target_id"])])
            last = int(p["target_id"])
        reg.setInputSource(clouds[int(p["source_id"])])
        reg.align(np.array(p["guess"], np.float32).reshape(4, 4).T)
        reg.getFitnessScore(DBL_MAX)
    return time.perf_counter() - t0


def reference_arm(args, rank, emit=print):
    """--impl reference: the reference's CPU implementation of the path (oracle restatement, OpenMP on
    all host cores) on a bounded sample of the same workload.  Rank 0 only."""
    if rank != 0:
        return 0
    from oracle import oracle_py as oracle
    cores = oracle.lib().orc_max_threads()
    if args.workload == "loop":
        from delta_graph_slam_b200.loop_batch import make_pairs
        from delta_graph_slam_b200.synth.loop_scenario import loop_scenario
        sc = loop_scenario(oracle.synth_traj, n_targets=1, n_candidates=args.loop_cpu_pairs)
        clouds = {cid: oracle.voxelgrid(oracle.synth_scan(P, noise_seed=ns), 0.1)["out"] for cid, P, ns in sc["targets"] + sc["candidates"]}
        pairs = make_pairs([(t, c, g) for t, c, g, _ in sc["pairs"]])
        for _ in range(args.warmup):
            oracle_loop_pairs(oracle, clouds, pairs)
        dt = sum(oracle_loop_pairs(oracle, clouds, pairs) for _ in range(args.steps))
        value = args.steps * len(pairs) / dt
        metric, unit = "loop pairs/sec (NDT + fitness)", "pairs/s"
        sample = f"{len(pairs)} candidate pairs of one new keyframe per step (setInputTarget once, then align + getFitnessScore per candidate), oracle restatement of ndt_omp + pcl::Registration"
        cfg = {"workload": "LoopDetector batch: NDT DIRECT7 + getFitnessScore candidate pairs (BASELINE configs[3])", "pairs_per_step": len(pairs)}
    else:
        n = args.ref_frames
        clouds = [oracle.synth_scan(oracle.synth_traj(k), noise_seed=1000 + k) for k in range(n)]
        pre, _ = oracle_odometry(oracle)

        def one_step():
            _, odo = oracle_odometry(oracle)
            run_sequence(pre, odo, clouds)
        for _ in range(args.warmup):
            one_step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            one_step()
        dt = time.perf_counter() - t0
        value = args.steps * (n - 1) / dt
        metric, unit = "scan registrations/sec (NDT keyframe odometry)", "registrations/s"
        sample = f"first {n} frames of the sequence per step (VoxelGrid 0.1 + NDT DIRECT7 keyframe odometry), oracle restatement of pcl::VoxelGrid + ndt_omp"
        cfg = {"workload": "scan_matching_odometry: synthetic KITTI-like HDL-64 scans, VoxelGrid 0.1 m + NDT DIRECT7 keyframe odometry (BASELINE configs[1])", "frames_per_step": n,
               "registration": "NDT_OMP DIRECT7 res 1.0 eps 0.01 max_iter 64"}
    emit(json.dumps({
        "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak" if args.workload == "odometry" else "strong", "vs_baseline": None, "dtype": "f32 per hit, f64 sums", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))
    return 0


class Ctx:
    """Per-process bench context: device, rank plumbing, event timing on the engine's stream."""

    def __init__(self, args):
        import torch
        self.torch = torch
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.dev = self.local_rank
        torch.cuda.set_device(self.dev)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.dev))
            self.dist = dist

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        if self.dist is None:
            return vals
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=f"cuda:{self.dev}")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return tuple(float(x) for x in t)

    def timed(self, fn, stream_ptr, steps, warmup):
        """fn(i) -> stats.  (seconds by CUDA events on the engine stream, wall seconds, stats), max over ranks."""
        torch = self.torch
        for i in range(warmup):
            fn(i)
        self.barrier()
        stream = torch.cuda.ExternalStream(stream_ptr, device=f"cuda:{self.dev}")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stats = []
        w0 = time.perf_counter()
        e0.record(stream)
        for i in range(steps):
            stats.append(fn(i))
        e1.record(stream)
        e1.synchronize()
        wall = time.perf_counter() - w0
        self.barrier()
        sec, wall = self.max_over_ranks(e0.elapsed_time(e1) * 1e-3, wall)
        return sec, wall, stats


def bench_odometry(ctx, odom_params=None, frames=None, steps=None, warmup=None, label="NDT", cpu_frames=None):
    import delta_graph_slam_b200 as eng
    from delta_graph_slam_b200 import synth
    torch, args, dev, rank, world = ctx.torch, ctx.args, ctx.dev, ctx.rank, ctx.world
    odom_params = odom_params or ODOM_PARAMS
    steps = steps or args.steps
    warmup = warmup or args.warmup
    cpu_frames = args.cpu_frames if cpu_frames is None else cpu_frames
    is_ndt = odom_params["registration_method"] == "NDT_OMP"

    # ---- synthetic sequence, generated on the device; a pinned host copy feeds the e2e leg
    F = frames or args.frames
    rays = synth.num_rays(synth.HDL64)
    d_raw = torch.empty((F, rays, 4), dtype=torch.float32, device=f"cuda:{dev}")
    counts = []
    for k in range(F):
        P = synth.traj_kitti_like(k + 5000 * rank)
        counts.append(synth.scan_to_device(d_raw[k].data_ptr(), P, synth.HDL64, scene_seed=1, noise_seed=1000 + k + 5000 * rank, device=dev))
    h_raw = torch.empty((F, rays, 4), dtype=torch.float32, pin_memory=True)
    h_raw.copy_(d_raw)
    torch.cuda.synchronize()
    h_np = h_raw.numpy()
    host_clouds = [h_np[k, : counts[k]] for k in range(F)]
    dev_clouds = [eng.DeviceCloud(d_raw[k].data_ptr(), counts[k], d_raw) for k in range(F)]
    d_ds = torch.empty((3, rays, 4), dtype=torch.float32, device=f"cuda:{dev}")
    ds_bufs = [eng.DeviceCloud(d_ds[j].data_ptr(), rays, d_ds) for j in range(3)]
    ds_buf = ds_bufs[0]

    def new_pipeline():
        pre = eng.Prefilter(PREFILTER_PARAMS, device=dev, out=DEVNULL)
        odo = eng.ScanMatchingOdometry(odom_params, device=dev, out=DEVNULL)
        return pre, odo

    # The front end runs as the reference runs it: prefiltering and scan matching are two nodelets
    # joined by a topic, so the filter of scan k+1 is in flight while scan k is matched (eng.FrontEnd).
    # The filter handle's persistent kernel is given args.filter_sms SMs, the registration the rest.
    def new_front_end(bufs):
        pre, odo = new_pipeline()
        odo.prepare_promotion = bool(args.filter_sms) and args.prepare
        return eng.FrontEnd(pre, odo, bufs, filter_sms=args.filter_sms), pre, odo

    # ---- device-resident leg (value) + roofline of the align kernel
    fe_d, pre_d, odo_d = new_front_end(ds_bufs)
    odo_d.registration.setTiming(True)

    def step_device(i):
        odo_d.keyframe = None  # restart the sequence; engine buffers stay allocated
        acc = dict(alg_bytes=0, evals=0, hits=0)

        def on_frame(k, filtered):
            if k > 0:
                r = odo_d.registration.getResult()
                # NDT pass: 16 B per source point + 48 B per (point, voxel) hit.  GICP: a linearize pass reads the
                # point, its covariance, the correspondence's point and covariance and writes the Mahalanobis
                # matrix (16 + 48 + 16 + 48 + 48 B per correspondence); an error pass re-reads 16 + 16 + 48 B
                # (GICP outer iteration) — approximated with hits = linearize correspondences
                acc["alg_bytes"] += (16 * filtered.n * r["passes"] + 48 * r["hits"]) if is_ndt else (16 * filtered.n * r["passes"] + 160 * r["hits"])
                acc["evals"] += r["passes"]
                acc["hits"] += r["hits"]
        fe_d.run(dev_clouds, on_frame=on_frame)
        return dict(alg_bytes=acc["alg_bytes"], evals=acc["evals"], hits=acc["hits"], keyframes=odo_d.num_keyframes)

    sampler = ClockSampler(dev)
    with sampler:
        c0 = odo_d.registration.counters()
        sec_d, wall_d, st_d = ctx.timed(step_device, odo_d.registration.stream(), steps, warmup)
        c1 = odo_d.registration.counters()
    regs_per_step = F - 1
    value = world * steps * regs_per_step / sec_d
    n_al = c1["timed_aligns"]  # the counters also saw the warm-up steps: per-launch averages over everything the library timed
    align_ms = c1["align_kernel_ms"]
    alg_bytes_per_launch = sum(s["alg_bytes"] for s in st_d) / (steps * regs_per_step)
    avg_launch_ms = align_ms / max(n_al, 1)
    achieved = alg_bytes_per_launch / (avg_launch_ms * 1e-3) / 1e9
    peak, peak_kind = load_peaks()
    launches_timed = (c1["launches_total"] - c0["launches_total"]) * steps // (steps + warmup)

    # ---- host-buffer leg (e2e): the calls the reference's nodelets make, host clouds in and out
    # caller-owned output clouds of the filter (pcl::Filter::filter(output)), page-locked; three in
    # rotation because the odometry keeps the keyframe's cloud while the next scans are filtered
    h_out = torch.empty((3, rays, 4), dtype=torch.float32, pin_memory=True).numpy()
    fe_h, pre_h, odo_h = new_front_end([h_out[j] for j in range(3)])

    def step_host(i):
        odo_h.keyframe = None
        acc = dict(h2d=0, d2h=0)

        def on_frame(k, filtered):
            # raw scan up, filtered cloud down (the /filtered_points message), filtered cloud up again
            # (setInputSource of the odometry nodelet), result record down
            acc["h2d"] += host_clouds[k].nbytes + filtered.nbytes
            acc["d2h"] += filtered.nbytes + 128
        fe_h.run(host_clouds, on_frame=on_frame)
        return acc
    sec_h, wall_h, st_h = ctx.timed(step_host, odo_h.registration.stream(), steps, warmup)
    e2e_value = world * steps * regs_per_step / sec_h

    # ---- parity of the two legs (same inputs -> same poses) and odometry sanity vs ground truth
    nchk = min(50, F)
    fe_c, _, _ = new_front_end(ds_bufs)
    poses_dev = fe_c.run(dev_clouds[:nchk])
    fe_c2, _, _ = new_front_end([h_out[j] for j in range(3)])
    poses_host = fe_c2.run(host_clouds[:nchk])
    legs_equal = all(np.array_equal(a, b) for a, b in zip(poses_dev, poses_host))
    # the pipelined front end against the plain loop (filter, then match, one scan at a time) on the same SM budgets
    pre_s, odo_s = new_pipeline()
    if pre_s.filter is not None and args.filter_sms:
        pre_s.filter.setSmBudget(args.filter_sms)
        odo_s.registration.setSmBudget(148 - args.filter_sms)  # no prepared promotions here: they must not change a pose
    poses_seq = run_sequence(pre_s, odo_s, dev_clouds[:nchk], out_buf=ds_buf)
    pipeline_equal = all(np.array_equal(a, b) for a, b in zip(poses_dev, poses_seq))
    P0 = synth.traj_kitti_like(5000 * rank)
    gt = np.linalg.inv(P0) @ synth.traj_kitti_like(nchk - 1 + 5000 * rank)
    drift = float(np.linalg.norm(poses_dev[nchk - 1][:3, 3] - gt[:3, 3]))

    # ---- CPU baseline (rank 0, N = 1): the oracle on the first frames of the same sequence
    cpu = None
    if rank == 0 and world == 1 and cpu_frames > 1:
        v, cores, dt = time_oracle_odometry(host_clouds, min(cpu_frames, F), odom_params)
        cpu = {"value": v, "unit": "registrations/s", "cores": cores, "kind": "port",
               "sample": f"first {min(cpu_frames, F)} frames of the same sequence ({dt:.1f} s): oracle restatement of pcl::VoxelGrid 0.1 m + {'ndt_omp DIRECT7' if is_ndt else 'fast_gicp'} keyframe odometry, OpenMP on all host cores"}

    out = {
        "metric": f"scan registrations/sec ({label} keyframe odometry)", "value": value, "unit": "registrations/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": 1e3 * sec_d / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 per hit, f64 sums" if is_ndt else "f64 (f32 points and NN search)",
        "data": "synthetic",
        "config": {"workload": f"scan_matching_odometry: {F} consecutive synthetic KITTI-like HDL-64 scans, VoxelGrid 0.1 m + {label} keyframe odometry (BASELINE configs[{1 if is_ndt else 2}])",
                   "frames_per_step": F, "points_per_scan": int(np.mean(counts)),
                   "registration": "NDT_OMP-equivalent DIRECT7 res 1.0 eps 0.01 max_iter 64" if is_ndt else "FAST_GICP-equivalent k 20, max corr 2.5 m, eps 0.01, max_iter 64, LM, PLANE",
                   "l2": f"each step streams {F} distinct scans ({F * rays * 16 / 1e9:.1f} GB) through the engine: inputs larger than L2", "multi_gpu": "independent sequence per GPU (replicas only)",
                   "front_end": f"pipelined as the reference's two nodelets: filter of scan k+1 in flight while scan k is matched; filter handle {args.filter_sms - (16 if args.prepare else 0)} SMs, registration {148 - args.filter_sms} SMs{', 16 SMs for the NDT grid of a predicted next keyframe built during its own registration' if args.prepare else ''}" if args.filter_sms else "sequential: filter, then match",
                   "keyframes_per_step": st_d[-1]["keyframes"], "passes_per_registration": st_d[-1]["evals"] / regs_per_step},
        "e2e": {"value": e2e_value, "unit": "registrations/s", "h2d_bytes_per_step": st_h[-1]["h2d"], "d2h_bytes_per_step": st_h[-1]["d2h"], "ms_per_step": 1e3 * sec_h / steps,
                "wall_ms_per_step": 1e3 * wall_h / steps},
        "gpu_launches": int(launches_timed),
        "roofline": {"bound": "hbm", "kernel": "k_ndt_align<7>" if is_ndt else "k_gicp_align", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_kind": peak_kind,
                     "traffic": load_traffic("k_ndt_align_single" if is_ndt else "k_gicp_align"), "traffic_unit": "DRAM bytes per launch (ncu --set full capture of one registration of this workload)",
                     "algorithmic_bytes_per_launch": alg_bytes_per_launch, "avg_launch_ms": avg_launch_ms, "launches": int(n_al),
                     "share_of_step": align_ms / max(n_al, 1) * regs_per_step / (1e3 * sec_d / steps),
                     "note": "working set (source cloud + staged voxel grid) is L2/SMEM resident, so DRAM traffic is far below the algorithmic bytes; the kernel is latency / issue bound, see DESIGN.md"},
        "cpu_baseline": cpu,
        "clocks": sampler.summary(),
        "checks": {"device_and_host_legs_bit_identical_first_frames": bool(legs_equal), "pipelined_equals_sequential_first_frames": bool(pipeline_equal), "frames_checked": nchk, "position_error_m_after_frames_checked": drift,
                   "wall_ms_per_step": 1e3 * wall_d / steps},
    }
    del d_raw, h_raw
    torch.cuda.empty_cache()
    return out


def bench_loop(ctx, steps, warmup, dense=False):
    """BASELINE.json configs[3]: the loop-candidate batch, whole targets sharded over the ranks.
    dense=True: configs[4], the 128-beam 1M-point stress scans with NDT DIRECT1 on a smaller batch."""
    import delta_graph_slam_b200 as eng
    from delta_graph_slam_b200 import loop_batch, synth
    from delta_graph_slam_b200.synth.loop_scenario import loop_scenario
    torch, args, dev, rank, world = ctx.torch, ctx.args, ctx.dev, ctx.rank, ctx.world

    n_targets, n_candidates = (args.dense_targets, args.dense_candidates) if dense else (args.loop_targets, args.loop_candidates)
    sensor = synth.DENSE128 if dense else synth.HDL64
    loop_params = dict(LOOP_PARAMS, reg_nn_search_method="DIRECT1") if dense else LOOP_PARAMS
    sc = loop_scenario(synth.traj_kitti_like, n_targets=n_targets, n_candidates=n_candidates)
    pairs = loop_batch.make_pairs([(t, c, g) for t, c, g, _ in sc["pairs"]])
    n_pairs = len(pairs)
    shards = loop_batch.shard_by_target(pairs["target_id"], world)
    mine = pairs[shards[rank]]
    need = set(loop_batch.needed_clouds(pairs, shards[rank]))
    my_targets = sorted(set(mine["target_id"].tolist()))

    # ---- this rank's keyframe clouds: ray-cast and down-sampled (0.1 m) on the device, packed in one buffer
    rays = synth.num_rays(sensor)
    vg = eng.VoxelGrid(device=dev)
    vg.setLeafSize(0.1, 0.1, 0.1)
    d_raw = torch.empty((rays, 4), dtype=torch.float32, device=f"cuda:{dev}")
    d_tmp = torch.empty((rays, 4), dtype=torch.float32, device=f"cuda:{dev}")
    specs = [(cid, P, ns) for cid, P, ns in sc["targets"] + sc["candidates"] if cid in need]
    cap = 400000 if dense else 70000
    vg_ms, raw_n = [], []
    d_kf = torch.empty((len(specs), cap, 4), dtype=torch.float32, device=f"cuda:{dev}")
    kf_n, slot_of = [], {}
    for s, (cid, P, ns) in enumerate(specs):
        n = synth.scan_to_device(d_raw.data_ptr(), P, sensor, scene_seed=1, noise_seed=ns, device=dev)
        vg.setInputCloud(eng.DeviceCloud(d_raw.data_ptr(), n, d_raw), is_dense=False)
        f = vg.filter(out=eng.DeviceCloud(d_tmp.data_ptr(), rays, d_tmp))
        if dense and s >= 2:  # VoxelGrid of the 1 M-point scans, CUDA events on the filter's stream
            st = torch.cuda.ExternalStream(vg._reg.stream(), device=f"cuda:{dev}")
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            f = vg.filter(out=eng.DeviceCloud(d_tmp.data_ptr(), rays, d_tmp))
            e1.record(st)
            e1.synchronize()
            vg_ms.append(e0.elapsed_time(e1))
            raw_n.append(n)
        if f.n > cap:
            raise RuntimeError(f"down-sampled keyframe has {f.n} points, more than the bench buffer holds")
        d_kf[s, : f.n].copy_(d_tmp[: f.n])  # the filter call returned after its stream drained
        torch.cuda.synchronize()            # d_tmp / d_raw are reused by the next keyframe
        kf_n.append(f.n)
        slot_of[cid] = s
    torch.cuda.synchronize()
    dev_cloud = {cid: eng.DeviceCloud(d_kf[s].data_ptr(), kf_n[s], d_kf) for cid, s in slot_of.items()}
    h_kf = torch.empty((len(specs), cap, 4), dtype=torch.float32, pin_memory=True)
    h_kf.copy_(d_kf)
    torch.cuda.synchronize()
    h_np = h_kf.numpy()
    host_cloud = {cid: h_np[s, : kf_n[s]] for cid, s in slot_of.items()}

    reg = eng.select_registration_method(loop_params, device=dev, out=DEVNULL)
    reg.setTiming(True)
    gdev = f"cuda:{dev}"

    # ---- device-resident leg: keyframes already in HBM; a step re-registers the targets (their NDT
    # grid and NN structure are rebuilt: the setInputTarget of every new keyframe), aligns and scores
    # this rank's pairs, and gathers all result records
    for cid in need:
        reg.cloudPut(cid, dev_cloud[cid])
    last = {}

    def step_device(i):
        for t in my_targets:
            reg.cloudPut(t, dev_cloud[t])
        local = reg.alignBatch(mine, with_fitness=True, fitness_max_range=DBL_MAX)
        last["res"] = loop_batch.gather_results(local, shards, rank, world, device=gdev)
        last["local"] = local
        return reg.batchTiming()
    sampler = ClockSampler(dev)
    with sampler:
        c0 = reg.counters()
        sec_d, wall_d, st_d = ctx.timed(step_device, reg.stream(), steps, warmup)
        c1 = reg.counters()
    value = steps * n_pairs / sec_d
    res = last["res"]
    local = last["local"]
    alg_bytes = float(sum(16 * kf_n[slot_of[int(p["source_id"])]] * int(r["passes"]) + 48 * int(r["hits"]) for p, r in zip(mine, local)))
    align_ms = float(np.mean([s["align_kernel_ms"] for s in st_d]))
    fit_ms = float(np.mean([s["fitness_ms"] for s in st_d]))
    peak, peak_kind = load_peaks()
    achieved = alg_bytes / (align_ms * 1e-3) / 1e9 if align_ms > 0 else 0.0

    # ---- host-buffer leg (e2e): every keyframe cloud of the share uploaded from pinned host memory each step
    reg_h = eng.select_registration_method(loop_params, device=dev, out=DEVNULL)
    h2d = sum(host_cloud[cid].nbytes for cid in need) + mine.nbytes

    def step_host(i):
        for cid in need:
            reg_h.cloudPut(cid, host_cloud[cid])
        local_h = reg_h.alignBatch(mine, with_fitness=True, fitness_max_range=DBL_MAX)
        last["res_h"] = loop_batch.gather_results(local_h, shards, rank, world, device=gdev)
        return None
    sec_h, wall_h, _ = ctx.timed(step_host, reg_h.stream(), steps, warmup)
    e2e_value = steps * n_pairs / sec_h

    # ---- checks: both legs identical; recovered poses against the scenario's ground truth
    legs_equal = bool(np.array_equal(res.view(np.uint8), last["res_h"].view(np.uint8)))
    err_t = []
    for r, (_, _, _, rel) in zip(res, sc["pairs"]):
        T = np.array(r["transformation"], np.float32).reshape(4, 4).T
        err_t.append(float(np.max(np.abs(T[:3, 3] - rel[:3, 3]))))
    err_t = np.array(err_t)

    cpu = None
    if rank == 0 and world == 1 and args.loop_cpu_pairs > 0 and not dense:
        from oracle import oracle_py as oracle
        k = min(args.loop_cpu_pairs, args.loop_candidates)
        sub = pairs[:k]
        clouds = {int(c): np.array(host_cloud[int(c)]) for c in set(sub["target_id"].tolist()) | set(sub["source_id"].tolist())}
        dt = oracle_loop_pairs(oracle, clouds, sub)
        cpu = {"value": k / dt, "unit": "pairs/s", "cores": oracle.lib().orc_max_threads(), "kind": "port",
               "sample": f"first {k} pairs of the same batch ({dt:.1f} s): setInputTarget once, then align + getFitnessScore per candidate; oracle restatement of ndt_omp + pcl::Registration, OpenMP on all host cores"}

    out = {
        "metric": "loop pairs/sec (NDT + fitness)", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * sec_d / steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 per hit, f64 sums", "data": "synthetic",
        "config": {"workload": (f"dense-scan stress: 128-beam 1M-ray scans, VoxelGrid 0.1 m + {n_targets} x {n_candidates} = {n_pairs} NDT DIRECT1 + getFitnessScore pairs (BASELINE configs[4])" if dense else
                                f"LoopDetector batch: {n_targets} new keyframes x {n_candidates} candidates = {n_pairs} NDT DIRECT7 + getFitnessScore pairs (BASELINE configs[3])"),
                   "points_per_keyframe": int(np.mean(kf_n)), "registration": f"NDT_OMP-equivalent {'DIRECT1' if dense else 'DIRECT7'} res 1.0 eps 0.01 max_iter 64, fitness max_range DBL_MAX",
                   "sharding": "whole targets per rank, one all-gather of 104-byte result records", "pairs_this_rank": int(len(mine)),
                   "l2": f"{len(specs)} distinct keyframe clouds ({sum(kf_n) * 16 / 1e9:.2f} GB) per rank: inputs larger than L2",
                   "passes_per_registration": float(np.mean(res["passes"])), "reference_evaluations_per_registration": float(np.mean(res["evaluations"])), "converged_fraction": float(np.mean(res["converged"]))},
        "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(local.nbytes), "ms_per_step": 1e3 * sec_h / steps},
        "gpu_launches": int((c1["launches_total"] - c0["launches_total"]) * steps // (steps + warmup)),
        "roofline": {"bound": "hbm", "kernel": f"k_ndt_align<{1 if dense else 7}> ({'1, 2 or 4 CTAs per registration, whichever fills the last round of the batch best' if len(mine) >= 148 else str(148 // max(len(mine), 1)) + ' CTAs per registration'})", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_kind": peak_kind,
                     "traffic": None, "traffic_per_registration": load_traffic("k_ndt_align_batch_per_registration"), "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": align_ms, "share_of_step": align_ms / (1e3 * sec_d / steps), "fitness_ms_per_step": fit_ms},
        "cpu_baseline": cpu,
        "clocks": sampler.summary(),
        "checks": {"device_and_host_legs_bit_identical": legs_equal, "median_translation_error_m": float(np.median(err_t)), "pairs_within_5cm_of_ground_truth": float(np.mean(err_t < 0.05)),
                   "wall_ms_per_step": 1e3 * wall_d / steps},
    }
    if dense and vg_ms:
        # VoxelGrid(N -> M): 16 N + 16 M algorithmic bytes
        ms, nr = float(np.median(vg_ms)), float(np.mean(raw_n))
        out["voxelgrid"] = {"raw_points": int(nr), "filtered_points": int(np.mean(kf_n)), "ms_per_scan": ms, "scans_per_s": 1e3 / ms,
                            "achieved_gbs": (16 * nr + 16 * float(np.mean(kf_n))) / (ms * 1e-3) / 1e9, "peak_gbs": peak}
    del d_kf, h_kf
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="odometry", choices=["odometry", "loop"], help="headline line: odometry (configs[1]) or the loop-candidate batch (configs[3])")
    ap.add_argument("--frames", type=int, default=1000, help="scans per sequence (one odometry step = one pass over the sequence)")
    ap.add_argument("--cpu-frames", type=int, default=24, help="frames of the sequence the CPU baseline runs")
    ap.add_argument("--ref-frames", type=int, default=8, help="frames per step of the --impl reference arm")
    ap.add_argument("--loop-targets", type=int, default=256, help="new keyframes of the loop batch (x candidates = pairs)")
    ap.add_argument("--loop-candidates", type=int, default=16)
    ap.add_argument("--loop-cpu-pairs", type=int, default=8, help="pairs of the batch the CPU baseline registers")
    ap.add_argument("--no-loop", action="store_true", help="skip the loop-batch leg of the default (odometry) run")
    ap.add_argument("--no-gicp", action="store_true", help="skip the FAST_GICP odometry leg (BASELINE configs[2]) of the default run")
    ap.add_argument("--prepare", action="store_true", help="build a predicted next keyframe's target structures during its own registration (b200reg_prepare_promotion; measured: no net gain, off by default)")
    ap.add_argument("--no-dense", action="store_true", help="skip the dense-scan stress leg (BASELINE configs[4]) of the default run")
    ap.add_argument("--dense-targets", type=int, default=8)
    ap.add_argument("--dense-candidates", type=int, default=8)
    ap.add_argument("--filter-sms", type=int, default=40, help="SMs given to the prefilter handle's persistent kernel in the pipelined front end (the registration takes the rest)")
    ap.add_argument("--gicp-frames", type=int, default=300, help="frames of the sequence the FAST_GICP leg runs per step")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    # stdout carries exactly ONE JSON line: anything a library prints there (NCCL's version banner,
    # OpenMP notices) is sent to stderr instead, and the line itself goes to the saved descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        os.write(json_fd, (line + "\n").encode())
    if args.impl == "reference":
        return reference_arm(args, rank, emit)

    ctx = Ctx(args)
    if args.workload == "loop":
        out = bench_loop(ctx, args.steps, args.warmup)
    else:
        out = bench_odometry(ctx)
        keys = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "checks")
        if not args.no_loop:
            lb = bench_loop(ctx, max(1, min(args.steps, 2)), 3)
            out["loop_batch"] = {k: lb[k] for k in keys}
        if not args.no_dense:
            db = bench_loop(ctx, 1, 3, dense=True)
            out["dense_stress"] = {k: db[k] for k in keys + ("voxelgrid",) if k in db}
        if not args.no_gicp:
            gb = bench_odometry(ctx, GICP_ODOM_PARAMS, frames=min(args.frames, args.gicp_frames), steps=max(1, min(args.steps, 2)), warmup=3, label="FAST_GICP", cpu_frames=min(args.cpu_frames, 12))
            out["gicp_odometry"] = {k: gb[k] for k in keys}
    if ctx.rank == 0:
        emit(json.dumps(out))
    if ctx.dist is not None:
        ctx.dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
