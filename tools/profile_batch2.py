"""ncu launch-list driver for a loop batch: clouds are made on the CPU-free path first, the profiled
region is one alignBatch (target builds + align + fitness).  Use with ncu --kernel-name filters or -s."""
import os, sys, io, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import delta_graph_slam_b200 as eng
from delta_graph_slam_b200 import synth, loop_batch
from delta_graph_slam_b200.synth.loop_scenario import loop_scenario

n_t, n_c = 16, 16
sc = loop_scenario(synth.traj_kitti_like, n_targets=n_t, n_candidates=n_c)
rays = synth.num_rays(synth.HDL64)
vg = eng.VoxelGrid(); vg.setLeafSize(0.1, 0.1, 0.1)
d_raw = torch.empty((rays, 4), dtype=torch.float32, device="cuda:0")
reg = eng.select_registration_method(dict(registration_method="NDT_OMP", reg_resolution=1.0), out=io.StringIO())
keep = {}
for cid, P, ns in sc["targets"] + sc["candidates"]:
    n = synth.scan_to_device(d_raw.data_ptr(), P, synth.HDL64, 1, ns, 0)
    out = torch.empty((rays, 4), dtype=torch.float32, device="cuda:0")
    vg.setInputCloud(eng.DeviceCloud(d_raw.data_ptr(), n, d_raw), is_dense=False)
    keep[cid] = vg.filter(out=eng.DeviceCloud(out.data_ptr(), rays, out))
    reg.cloudPut(cid, keep[cid])
pairs = loop_batch.make_pairs([(t, c, g) for t, c, g, _ in sc["pairs"]])
reg.alignBatch(pairs)
torch.cuda.synchronize()
reg.setTiming(True)
for rep in range(3):
    t0 = time.perf_counter()
    for t in range(n_t):
        reg.cloudPut(t, keep[t])
    t1 = time.perf_counter()
    res = reg.alignBatch(pairs)
    t2 = time.perf_counter()
    bt = reg.batchTiming()
    print(f"pairs {len(pairs)}: cloudPut {1e3 * (t1 - t0):.2f} ms  alignBatch {1e3 * (t2 - t1):.2f} ms  (align kernel {bt['align_kernel_ms']:.2f} ms, fitness {bt['fitness_ms']:.2f} ms, rest {1e3 * (t2 - t1) - bt['align_kernel_ms'] - bt['fitness_ms']:.2f} ms)")
