"""The helper-thread copy pool behind pageable caller clouds (csrc/host_copy.hpp): a CPU-only build-and-run check."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def test_host_copy_pool_copies_exactly_and_survives_concurrent_callers(tmp_path):
    exe = str(tmp_path / "host_copy_check")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-O2", "-std=c++17", "-pthread", os.path.join(HERE, "cpp", "host_copy_check.cpp"), "-o", exe], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True, timeout=300).stdout
    assert out.strip().endswith("ok") and "MISMATCH" not in out
