"""CPU-side checks of the boundary: the CUDA library loads, exports every symbol include/lvio_b200.h declares, and
refuses loudly to compute without a device (no CPU fallback on the product path)."""
import ctypes
import os
import re
import subprocess

import pytest

from lvio_fusion_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "lvio_b200.h")).read()
    declared = set(re.findall(r"LVB_API [a-z_ \*]+?(lvb_[a-z0-9_]+)\(", hdr))
    assert len(declared) >= 30
    lib = ctypes.CDLL(_capi.LIB_PATH)
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert declared == set(_capi.EXPORTED_SYMBOLS), declared ^ set(_capi.EXPORTED_SYMBOLS)


def test_oracle_mirrors_the_same_abi():
    from oracle import binding
    orc = binding.load()
    for name in _capi._SIGS:
        assert hasattr(orc, name)


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a machine without a GPU")
def test_no_cpu_fallback_without_a_device():
    lvb = _capi.load()
    h = ctypes.c_void_p()
    rc = lvb.ctx_create(0, None, ctypes.byref(h))
    assert rc == -2                                   # LVB_ERR_CUDA
    assert b"no CPU fallback" in lvb.last_error()


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a machine without a GPU")
def test_shim_compiles_and_reports_failure_without_a_device(tmp_path):
    import numpy as np
    import shim_util
    from lvio_fusion_b200 import synth
    exe = shim_util.build_shim_binary()
    d = synth.make_ba_problem(3, 30, with_imu=False, seed=2)
    shim_util.dump_ba(tmp_path / "in.bin", d, 3)
    p = subprocess.run([exe, "ba", str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert p.returncode == 3 and "no device" in p.stderr      # ceres::Solve reports FAILURE through the Summary


def test_lidar_and_imu_host_header_builds_and_fails_loudly(tmp_path):
    """include/lvio_b200/lidar_features.h against a pcl::PointXYZI-shaped record; without a device the calls return
    false (no CPU fallback), with one they run (empty inputs)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib_dir = os.path.join(root, "lvio_fusion_b200", "csrc")
    exe = str(tmp_path / "lidar_header")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I" + os.path.join(root, "include"), "-o", exe,
                           os.path.join(root, "tests", "cpp", "compile_lidar_header.cpp"), "-L" + lib_dir, "-llvio_b200", "-Wl,-rpath," + lib_dir])
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == (0 if _has_gpu() else 3)


def test_host_headers_are_cxx14():
    """Ceres 2.x, which the reference builds against, needs C++14 and nothing newer: the header-only host side must not either."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for src in ("test_shim.cpp", "compile_lidar_header.cpp", "test_host_solver.cpp"):
        subprocess.check_call(["g++", "-std=c++14", "-fsyntax-only", "-Wall", "-I" + os.path.join(root, "include"), os.path.join(root, "tests", "cpp", src)])


def test_shim_solves_from_several_threads_at_once(tmp_path):
    """The reference solves from the backend, the global and the relocator thread concurrently: the shim's context is per host
    thread (ceres_shim.h, lvb::Runtime).  Four threads build and solve the same window at once (C ABI served by the oracle here):
    four distinct runtimes, four results identical to the single-threaded one."""
    import numpy as np
    import shim_util
    from lvio_fusion_b200 import synth
    exe = shim_util.build_shim_binary_on_oracle()
    d = synth.make_ba_problem(4, 150, with_imu=True, seed=9)
    shim_util.dump_ba(tmp_path / "in.bin", d, 6)
    subprocess.check_call([exe, "ba", str(tmp_path / "in.bin"), str(tmp_path / "single.bin")])
    subprocess.check_call([exe, "ba_threads", str(tmp_path / "in.bin"), str(tmp_path / "multi.bin"), "4"])
    single = np.fromfile(tmp_path / "single.bin")
    assert single[-3] < 0.5 * single[-4]                      # final < initial cost
    for k in range(4):
        assert np.array_equal(np.fromfile(str(tmp_path / "multi.bin") + ".%d" % k), single)
