import os, sys, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import delta_graph_slam_b200 as eng
from oracle import oracle_py as O
from helpers import *
from delta_graph_slam_b200.loop_detector import KeyFrame, LoopDetector, isometry2d, candidate_guess
from delta_graph_slam_b200.loop_batch import make_pairs
clouds, pairs, rels = small_loop_scenario(O)
ref = OracleBatchEngine(O)
ndt = eng.select_registration_method(dict(registration_method="NDT_OMP", reg_resolution=1.0, reg_nn_search_method="DIRECT7"), out=io.StringIO())
for k, v in clouds.items():
    ref.cloudPut(k, v); ndt.cloudPut(k, v)
for t in (0, 1):
    new_est = isometry2d(3.0 + t, -1.0, 0.2)
    new = KeyFrame(t, clouds[t], new_est, 100.0)
    trip = []
    for p in pairs[pairs["target_id"] == t]:
        g = np.array(p["guess"], np.float64).reshape(4, 4).T
        g2 = np.array([[g[0, 0], g[0, 1], g[0, 3]], [g[1, 0], g[1, 1], g[1, 3]], [0, 0, 1.0]])
        c = KeyFrame(int(p["source_id"]), clouds[int(p["source_id"])], new_est @ g2, 1.0)
        trip.append((t, c.id, candidate_guess(new, c)))
    pp = make_pairs(trip)
    a = ref.alignBatch(pp); b = ndt.alignBatch(pp)
    for x, y, p in zip(a, b, pp):
        Ta = np.array(x["transformation"]).reshape(4, 4).T; Tb = np.array(y["transformation"]).reshape(4, 4).T
        print(p["target_id"], p["source_id"], "it", x["iterations"], y["iterations"], "ev", x["evaluations"], y["evaluations"], "dt %.2e" % np.abs(Ta[:3, 3] - Tb[:3, 3]).max(),
              "dR %.2e" % rot_angle(Ta[:3, :3], Tb[:3, :3]), "fit %.6f %.6f" % (x["fitness"], y["fitness"]), "score %.9f %.9f" % (x["score"], y["score"]))
