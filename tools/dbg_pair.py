import os, sys, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import delta_graph_slam_b200 as eng
from oracle import oracle_py as O
from helpers import *
from delta_graph_slam_b200.loop_detector import KeyFrame, isometry2d, candidate_guess
clouds, pairs, rels = small_loop_scenario(O)
p = [q for q in pairs if q["target_id"] == 1 and q["source_id"] == 9][0]
new_est = isometry2d(4.0, -1.0, 0.2)
g = np.array(p["guess"], np.float64).reshape(4, 4).T
g2 = np.array([[g[0, 0], g[0, 1], g[0, 3]], [g[1, 0], g[1, 1], g[1, 3]], [0, 0, 1.0]])
guess = candidate_guess(KeyFrame(1, None, new_est, 0), KeyFrame(9, None, new_est @ g2, 0))
print("guess diff vs scenario", np.abs(guess - np.array(p["guess"]).reshape(4, 4).T).max())
ref = O.Registration(O.NDT, resolution=1.0, nn_search=O.DIRECT7, trans_eps=0.01, max_iter=64)
ndt = eng.select_registration_method(dict(registration_method="NDT_OMP", reg_resolution=1.0), out=io.StringIO())
for r in (ref, ndt):
    r.setInputTarget(clouds[1]); r.setInputSource(clouds[9])
for it in range(1, 9):
    ref.set("max_iter", it); ndt.setMaximumIterations(it)
    ref.align(guess); ndt.align(guess)
    Ta, Tb = ref.getFinalTransformation(), ndt.getFinalTransformation()
    ra, rb = ref.info(), ndt.getResult()
    print("max_iter", it, "iters", ref.getFinalNumIteration(), rb["iterations"], "evals", int(ra[1]), rb["evaluations"], "dt %.3e" % np.abs(Ta[:3, 3] - Tb[:3, 3]).max(), "score %.10f %.10f" % (ra[0], rb["score"]))
e = O.euler_xyz(guess)
p0 = np.concatenate([guess[:3, 3], e]).astype(np.float64)
for pp in (p0, p0 + np.array([0.01, -0.02, 0.003, 1e-3, -2e-3, 1e-3])):
    s0, g0, H0 = ref.ndt_derivatives(pp); s1, g1, H1 = ndt.ndt_derivatives(pp)
    print("deriv: score rel %.2e  g rel %.2e  H rel %.2e" % (abs(s1 - s0) / abs(s0), np.abs(g1 - g0).max() / np.abs(g0).max(), np.abs(H1 - H0).max() / np.abs(H0).max()))
