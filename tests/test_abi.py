"""CPU checks of the drop-in boundary: libb200reg.so loads and exports every symbol
include/b200reg.h declares, struct layouts match the ctypes mirror, defaults follow
select_registration_method [REF src/hdl_graph_slam/registrations.cpp:26-119], and the product path
fails loudly (no CPU fallback) when there is no CUDA device.  No compute calls here."""
import ctypes as C
import io
import os
import re
import subprocess

import numpy as np
import pytest

import delta_graph_slam_b200 as pkg
from delta_graph_slam_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "b200reg.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200reg_[a-z0-9_]+)\s*\(", src)))


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    declared = header_functions()
    assert len(declared) >= 40
    for name in declared:
        assert hasattr(L, name), f"{name} is declared in include/b200reg.h but not exported by libb200reg.so"
    assert sorted(_lib.EXPORTS) == declared, "the Python binding's export list mirrors the header"
    assert L.b200reg_version().decode().startswith("b200reg")


def test_struct_layouts_match_the_header():
    # the C structs as the header spells them, compiled here with the host compiler
    prog = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "b200reg.h"
    int main() {
      printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(b200reg_config), sizeof(b200reg_result), sizeof(b200reg_pair), offsetof(b200reg_result, fitness), offsetof(b200reg_result, hits),
             offsetof(b200reg_pair, guess), offsetof(b200reg_config, rotation_epsilon));
      return 0;
    }'''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(prog)
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        subprocess.check_call([cc, "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")])
        got = [int(x) for x in subprocess.check_output([os.path.join(d, "t")]).split()]
    want = [C.sizeof(_lib.Config), C.sizeof(_lib.Result), C.sizeof(_lib.Pair), _lib.Result.fitness.offset, _lib.Result.hits.offset, _lib.Pair.guess.offset, _lib.Config.rotation_epsilon.offset]
    assert got == want
    assert _lib.RESULT_DTYPE.itemsize == C.sizeof(_lib.Result) == 104
    assert _lib.PAIR_DTYPE.itemsize == C.sizeof(_lib.Pair) == 80


def test_default_config_follows_the_reference_factory():
    L = _lib.load()
    for method in (_lib.METHOD_NDT, _lib.METHOD_GICP):
        cfg = _lib.Config()
        L.b200reg_default_config(method, C.byref(cfg))
        assert cfg.method == method
        assert cfg.resolution == 0.5                    # reg_resolution default of the NDT branch (:93)
        assert cfg.nn_search == _lib.DIRECT7            # reg_nn_search_method (:103)
        assert cfg.transformation_epsilon == 0.01       # reg_transformation_epsilon
        assert cfg.maximum_iterations == 64             # reg_maximum_iterations
        assert cfg.max_correspondence_distance == 2.5   # reg_max_correspondence_distance (:33)
        assert cfg.correspondence_randomness == 20      # reg_correspondence_randomness (:34)
        assert cfg.step_size == 0.1 and cfg.outlier_ratio == 0.55 and cfg.rotation_epsilon == 2e-3
        assert cfg.regularization == _lib.REG_PLANE and cfg.lsq_optimizer == _lib.LSQ_LM


def test_argument_validation_without_a_device():
    L = _lib.load()
    h = C.c_void_p()
    assert L.b200reg_create(None, C.byref(h)) == _lib.E_INVALID
    cfg = _lib.Config()
    L.b200reg_default_config(_lib.METHOD_NDT, C.byref(cfg))
    cfg.method = 77
    assert L.b200reg_create(C.byref(cfg), C.byref(h)) == _lib.E_INVALID
    assert L.b200reg_destroy(None) == _lib.E_INVALID
    assert L.b200reg_last_error(None) == b"null handle"


@pytest.mark.skipif(has_gpu(), reason="checks the no-device behaviour")
def test_no_cpu_fallback_without_a_gpu():
    """The product path must fail loudly: create -> B200REG_E_CUDA, the wrappers raise."""
    L = _lib.load()
    cfg = _lib.Config()
    L.b200reg_default_config(_lib.METHOD_NDT, C.byref(cfg))
    h = C.c_void_p()
    assert L.b200reg_create(C.byref(cfg), C.byref(h)) == _lib.E_CUDA
    assert not h.value
    with pytest.raises(pkg.B200RegError):
        pkg.NormalDistributionsTransform()
    with pytest.raises(pkg.B200RegError):
        pkg.select_registration_method(dict(registration_method="NDT_OMP"), out=io.StringIO())
    with pytest.raises(pkg.B200RegError):
        pkg.VoxelGrid()
    # nothing under delta_graph_slam_b200/ may import the oracle
    for dirpath, _, files in os.walk(os.path.join(ROOT, "delta_graph_slam_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, f"{f} reaches into oracle/"


@pytest.mark.skipif(has_gpu(), reason="checks the no-device behaviour")
def test_multi_gpu_batch_entry_validates_and_fails_loudly_without_devices():
    L = _lib.load()
    cfg = _lib.Config()
    L.b200reg_default_config(_lib.METHOD_NDT, C.byref(cfg))
    m = C.c_void_p()
    devs = (C.c_int * 2)(0, 1)
    assert L.b200reg_batch_create(C.byref(cfg), devs, 0, C.byref(m)) == _lib.E_INVALID
    assert L.b200reg_batch_create(C.byref(cfg), (C.c_int * 2)(3, 3), 2, C.byref(m)) == _lib.E_INVALID  # one handle per GPU
    assert L.b200reg_batch_create(C.byref(cfg), devs, 2, C.byref(m)) == _lib.E_CUDA and not m.value
    assert L.b200reg_batch_destroy(None) == _lib.E_INVALID
    assert L.b200reg_batch_last_error(None) == b"null batch object"
    from delta_graph_slam_b200.loop_batch import MultiGpuBatch
    with pytest.raises(pkg.B200RegError):
        MultiGpuBatch([0])


def test_factory_parameter_handling_mirrors_the_reference(monkeypatch):
    """select_registration_method: names, defaults, banners and the unknown-method warning of
    [REF src/hdl_graph_slam/registrations.cpp:22-124], checked against a recording stand-in handle."""
    from delta_graph_slam_b200 import registration as R
    calls = []

    class Fake:
        def __init__(self, device=0, **kw):
            calls.append(("ctor", type(self).__name__))

        def __getattr__(self, name):
            if name.startswith("set"):
                return lambda *a: calls.append((name,) + a)
            raise AttributeError(name)

    monkeypatch.setattr(R, "NormalDistributionsTransform", type("NormalDistributionsTransform", (Fake,), {}))
    monkeypatch.setattr(R, "FastGICP", type("FastGICP", (Fake,), {}))
    out = io.StringIO()
    R.select_registration_method({}, out=out)
    assert out.getvalue() == "registration: NDT_OMP DIRECT7 0.5 (0 threads)\n"
    assert calls == [("ctor", "NormalDistributionsTransform"), ("setTransformationEpsilon", 0.01), ("setMaximumIterations", 64), ("setResolution", 0.5),
                     ("setNeighborhoodSearchMethod", _lib.DIRECT7)]
    calls.clear()
    out = io.StringIO()
    R.select_registration_method(dict(registration_method="NDT_OMP", reg_num_threads=4, reg_nn_search_method="KDTREE", reg_resolution=2.0, reg_transformation_epsilon=0.1,
                                      reg_maximum_iterations=32), out=out)
    assert out.getvalue() == "registration: NDT_OMP KDTREE 2 (4 threads)\n"
    assert ("setNumThreads", 4) in calls and ("setNeighborhoodSearchMethod", _lib.KDTREE) in calls and ("setResolution", 2.0) in calls
    calls.clear()
    out = io.StringIO()
    R.select_registration_method(dict(registration_method="FAST_GICP", reg_max_correspondence_distance=2.0), out=out)
    assert out.getvalue() == "registration: FAST_GICP\n"
    assert calls == [("ctor", "FastGICP"), ("setNumThreads", 0), ("setTransformationEpsilon", 0.01), ("setMaximumIterations", 64), ("setMaxCorrespondenceDistance", 2.0),
                     ("setCorrespondenceRandomness", 20)]
    # unknown strings: the reference warns "use NDT" (:88-91) and, the string holding no "OMP", hands out the
    # single-thread pcl NDT (:94-99) — not a method this engine replaces: same warning, then the same refusal as "NDT".
    # An unknown string that does hold "OMP" reaches the NDT_OMP branch (:100-119) and gets the engine.
    calls.clear()
    R.select_registration_method(dict(registration_method="MY_OMP_THING"), out=io.StringIO())
    assert calls[0] == ("ctor", "NormalDistributionsTransform")
    for m in ("FAST_VGICP", "ICP", "GICP", "GICP_OMP", "NDT", "SOMETHING"):
        with pytest.raises(NotImplementedError):
            R.select_registration_method(dict(registration_method=m), out=io.StringIO())


def test_cpp_adapter_compiles_against_mock_pcl_and_refuses_to_run_without_a_gpu():
    d = os.path.join(ROOT, "tests", "cpp")
    subprocess.check_call(["make", "-C", d, "-s"])
    rc = subprocess.call([os.path.join(d, "adapter_check")], stdout=subprocess.DEVNULL)
    assert rc == (0 if has_gpu() else 3)
