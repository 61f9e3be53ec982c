import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_py
    oracle_py.lib()
    return oracle_py


@pytest.fixture(scope="session")
def scans(oracle):
    """Two consecutive synthetic HDL-64 scans (frames 0 and 1 of the kitti_like trajectory),
    raw and 0.1 m down-sampled by the oracle, plus the ground-truth relative pose."""
    import numpy as np
    P0, P1 = oracle.synth_traj(0), oracle.synth_traj(1)
    s0 = oracle.synth_scan(P0, noise_seed=1000)
    s1 = oracle.synth_scan(P1, noise_seed=1001)
    v0 = oracle.voxelgrid(s0, 0.1)["out"]
    v1 = oracle.voxelgrid(s1, 0.1)["out"]
    return dict(raw0=s0, raw1=s1, ds0=v0, ds1=v1, gt=np.linalg.inv(P0) @ P1)
