"""Developer probe: poses of the native front end over the same sequence as device-resident scans (run_device), host scans in the
two-nodelet form and host scans fused — are the legs bit-identical, and is each leg repeatable run to run?  Under gpurun."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import delta_graph_slam_b200 as eng
from delta_graph_slam_b200 import synth

F = int(sys.argv[1]) if len(sys.argv) > 1 else 300
prep = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rays = synth.num_rays(synth.HDL64)
d_raw = torch.empty((F, rays, 4), dtype=torch.float32, device="cuda:0")
counts = [synth.scan_to_device(d_raw[k].data_ptr(), synth.traj_kitti_like(k), synth.HDL64, scene_seed=1, noise_seed=1000 + k, device=0) for k in range(F)]
dev = [eng.DeviceCloud(d_raw[k].data_ptr(), counts[k], d_raw) for k in range(F)]
h_raw = torch.empty((F, rays, 4), dtype=torch.float32, pin_memory=True)
h_raw.copy_(d_raw)
host = [h_raw[k].numpy()[: counts[k]] for k in range(F)]
params = {**bench.ODOM_PARAMS, **bench.PREFILTER_PARAMS}
h_out = torch.empty((3, rays, 4), dtype=torch.float32, pin_memory=True).numpy()
h_al = torch.empty((rays, 4), dtype=torch.float32, pin_memory=True).numpy()


def first_diff(a, b):
    d = np.abs(a.reshape(len(a), -1) - b.reshape(len(b), -1)).max(axis=1)
    nz = np.nonzero(d)[0]
    return (int(nz[0]), float(d.max())) if len(nz) else (None, 0.0)


runs = {}
for name in ("device", "device2", "host2n", "host2n_again", "fused", "host2n_noaligned"):
    fe = eng.NativeFrontEnd(params, filter_sms=52, prepare_promotion=prep)
    if name.startswith("device"):
        poses, res, nf = fe.run_device(dev)
    elif name == "fused":
        poses = fe.run_host(host, filtered_bufs=None, aligned_out=h_al)
    elif name == "host2n_noaligned":
        poses = fe.run_host(host, filtered_bufs=[h_out[j] for j in range(3)], aligned_out=None)
    else:
        poses = fe.run_host(host, filtered_bufs=[h_out[j] for j in range(3)], aligned_out=h_al)
    runs[name] = np.asarray(poses).copy()
    print(name, "keyframes", fe.num_keyframes() if hasattr(fe, "num_keyframes") else "?", flush=True)
    del fe
for a, b in (("device", "device2"), ("host2n", "host2n_again"), ("device", "host2n"), ("device", "fused"), ("host2n", "fused"), ("host2n", "host2n_noaligned")):
    print(a, "vs", b, "first differing frame / max |delta|:", first_diff(runs[a], runs[b]))

# the same instance reused: a second pass over the sequence must start from a clean state (b200reg_frontend_reset)
for prep_mode in (0, 1, 2):
    fe = eng.NativeFrontEnd(params, filter_sms=52, prepare_promotion=prep_mode)
    a = np.asarray(fe.run_device(dev)[0]).copy()
    b = np.asarray(fe.run_device(dev)[0]).copy()
    c = np.asarray(fe.run_device(dev)[0]).copy()
    fe.set_timing(True)
    d = np.asarray(fe.run_device(dev)[0]).copy()
    print("reused instance, prepare", prep_mode, ": run 1 vs 2", first_diff(a, b), "run 2 vs 3", first_diff(b, c), "run 3 vs timed run", first_diff(c, d), "fresh vs run 1", first_diff(runs["device"], a) if prep_mode == prep else "-")
    del fe
