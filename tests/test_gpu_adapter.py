"""The C++ side of the drop-in boundary on a real GPU: include/b200reg_pcl.hpp (pcl::Registration / pcl::Filter adapters)
compiled against the mock of PCL's base classes (tests/cpp/mock_pcl — PCL itself is not in this image) and driven through
the call sequence of ScanMatchingOdometryNodelet::matching [REF apps/scan_matching_odometry_nodelet.cpp:180-228].
tests/cpp/adapter_check exits 0 only if: NDT and FAST_GICP recover a known shift through the virtual surface; align()
fills `output` with the transformed source; NO CPU kd-tree is built on the align path (the mock's initCompute models
pcl::Registration's, the adapters install their search object with setSearchMethodTarget(tree, true)); and the base
class's non-virtual getFitnessScore still works (tree built lazily, once) and agrees with the GPU fitness."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_pcl_adapters_run_through_the_registration_surface():
    d = os.path.join(ROOT, "tests", "cpp")
    subprocess.check_call(["make", "-C", d, "-s"])
    r = subprocess.run([os.path.join(d, "adapter_check")], capture_output=True, text=True, timeout=300)
    print(r.stdout)
    print(r.stderr)
    assert r.returncode == 0, f"adapter_check failed (rc {r.returncode}):\n{r.stdout}\n{r.stderr}"
    assert "FAILED" not in r.stdout
