"""Developer tool: the More-Thuente line-search trace of one NDT registration, engine (profiled kernel instantiation,
b200reg_get_trace) next to the oracle (NDT::trace), to see where two runs leave a common optimisation path.
  python tools/ndt_trace_compare.py [tx ty yaw]      (frames 0 / 1 of the synthetic sequence, 0.1 m VoxelGrid)
Under gpurun."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import delta_graph_slam_b200 as eng  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

tx, ty, yaw = (float(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (0.3, -0.9, -0.1)
s0 = O.voxelgrid(O.synth_scan(O.synth_traj(0), noise_seed=1000), 0.1)["out"]
s1 = O.voxelgrid(O.synth_scan(O.synth_traj(1), noise_seed=1001), 0.1)["out"]
guess = np.eye(4, dtype=np.float32)
guess[:2, :2] = [[np.cos(yaw), -np.sin(yaw)], [np.sin(yaw), np.cos(yaw)]]
guess[:3, 3] = [tx, ty, 0.0]
ref = O.Registration(O.NDT, resolution=1.0, nn_search=O.DIRECT7, trans_eps=0.01, max_iter=64)
ndt = eng.select_registration_method(dict(registration_method="NDT_OMP", reg_resolution=1.0, reg_nn_search_method="DIRECT7"), out=open(os.devnull, "w"))
ndt.setProfile(True)
for r in (ref, ndt):
    r.setInputTarget(s0)
    r.setInputSource(s1)
    r.align(guess)
to, te = ref.ndt_trace(), ndt.trace()
res = ndt.getResult()
print(f"oracle: iterations {ref.getFinalNumIteration()} evaluations {int(ref.info()[1])} | engine: iterations {res['iterations']} evaluations {res['evaluations']} passes {res['passes']}")
# the engine's first record is the initial derivative pass (not a line-search evaluation)
te = te[1:]
cols = "it st a_t score phi_t d_phi_t psi_t d_psi_t open conv phi_0 d_phi_0".split()
print("columns:", cols)
np.set_printoptions(linewidth=250, precision=12, suppress=False)
for k in range(max(len(to), len(te))):
    a = to[k] if k < len(to) else None
    b = te[k] if k < len(te) else None
    flag = ""
    if a is not None and b is not None and (abs(a[4] - b[4]) > 1e-6 * abs(a[4]) or abs(a[5] - b[5]) > 1e-4 * max(abs(a[5]), 1e-9)):
        flag = "  <-- differs"
    print(f"eval {k:3d}{flag}")
    if a is not None:
        print("   oracle", " ".join(f"{v:.10g}" for v in a))
    if b is not None:
        print("   engine", " ".join(f"{v:.10g}" for v in b[:12]), "solve_path", int(b[12]))
