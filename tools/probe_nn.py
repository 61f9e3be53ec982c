"""Developer probe: how many fitness queries leave the ring walk, and what the passes cost."""
import os, sys, time, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import delta_graph_slam_b200 as d
from oracle import oracle_py as O
from helpers import small_loop_scenario

def t(f, n=10, warm=2):
    for _ in range(warm): f()
    ts = []
    for _ in range(n):
        a = time.perf_counter(); f(); ts.append(time.perf_counter() - a)
    return np.median(ts) * 1e6

clouds, pairs, rels = small_loop_scenario(O, n_targets=2, n_candidates=4, leaf=0.1, stride=1)
ndt = d.select_registration_method(dict(registration_method="NDT_OMP", reg_resolution=1.0), out=io.StringIO())
for p, rel in zip(pairs, rels):
    ndt.setInputTarget(clouds[int(p["target_id"])]); ndt.setInputSource(clouds[int(p["source_id"])])
    ndt.align(np.array(p["guess"]).reshape(4, 4).T)
    f = ndt.getFitnessScore()
    st = ndt.nn_stats()
    us = t(lambda: ndt.getFitnessScore())
    us4 = t(lambda: ndt.getFitnessScore(4.0))
    print(f"pair {p['target_id']}-{p['source_id']}: n_src {st['queries']} far {st["far_pass"]} brute {st["brute_pass"]} fitness {f:.4f}  getFitnessScore {us:.0f} us  (max_range 4.0: {us4:.0f} us)  align {t(lambda: ndt.align(np.array(p['guess']).reshape(4,4).T)):.0f} us")
