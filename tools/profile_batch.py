"""Small deterministic driver for ncu: one loop-closure batch with more pairs than SMs (one CTA per
registration).  Not a bench: numbers printed here are not reported."""
import os, sys, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import delta_graph_slam_b200 as eng
from delta_graph_slam_b200 import synth, loop_batch
from delta_graph_slam_b200.synth.loop_scenario import loop_scenario

n_t, n_c = 10, 16
sc = loop_scenario(synth.traj_kitti_like, n_targets=n_t, n_candidates=n_c)
rays = synth.num_rays(synth.HDL64)
vg = eng.VoxelGrid(); vg.setLeafSize(0.1, 0.1, 0.1)
d_raw = torch.empty((rays, 4), dtype=torch.float32, device="cuda:0")
reg = eng.select_registration_method(dict(registration_method="NDT_OMP", reg_resolution=1.0), out=io.StringIO())
keep = []
for cid, P, ns in sc["targets"] + sc["candidates"]:
    n = synth.scan_to_device(d_raw.data_ptr(), P, synth.HDL64, 1, ns, 0)
    out = torch.empty((rays, 4), dtype=torch.float32, device="cuda:0")
    vg.setInputCloud(eng.DeviceCloud(d_raw.data_ptr(), n, d_raw), is_dense=False)
    f = vg.filter(out=eng.DeviceCloud(out.data_ptr(), rays, out))
    reg.cloudPut(cid, f)
    keep.append(out)
pairs = loop_batch.make_pairs([(t, c, g) for t, c, g, _ in sc["pairs"]])
for _ in range(2):
    res = reg.alignBatch(pairs)
print("pairs", len(pairs), "converged", res["converged"].mean(), "evals", res["evaluations"].mean())
