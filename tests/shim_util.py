"""Helpers for the C++ shim tests: dump a synthetic problem in the flat binary form tests/cpp/test_shim.cpp reads."""
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "test_shim")


def build_shim_binary():
    src = os.path.join(ROOT, "tests", "cpp", "test_shim.cpp")
    lib_dir = os.path.join(ROOT, "lvio_fusion_b200", "csrc")
    if not os.path.exists(os.path.join(lib_dir, "liblvio_b200.so")):
        raise RuntimeError("liblvio_b200.so not built")
    if not os.path.exists(BIN) or os.path.getmtime(BIN) < os.path.getmtime(src):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-I" + os.path.join(ROOT, "include"), "-o", BIN, src,
                               "-L" + lib_dir, "-llvio_b200", "-Wl,-rpath," + lib_dir])
    return BIN


def build_shim_binary_on_oracle():
    """The same program with the C ABI served by the CPU oracle (lvb_* renamed to orc_*): host-side logic checks without a device."""
    src = os.path.join(ROOT, "tests", "cpp", "test_shim.cpp")
    odir = os.path.join(ROOT, "oracle")
    subprocess.check_call(["make", "-s", "-C", odir])
    syms = subprocess.run(["nm", "-D", os.path.join(odir, "liboracle.so")], capture_output=True, text=True, check=True).stdout.split("\n")
    defs = ["-Dlvb_%s=orc_%s" % (t.split()[2][4:], t.split()[2][4:]) for t in syms if len(t.split()) == 3 and t.split()[1] == "T" and t.split()[2].startswith("orc_")]
    out = BIN + "_orc"
    defs.append("-DLVB_NO_RESIDENT_MAP")          # the oracle mirrors the per-call entry points only
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread"] + defs + ["-I" + os.path.join(ROOT, "include"), "-o", out, src, "-L" + odir, "-loracle", "-Wl,-rpath," + odir])
    return out


def dump_ba(path, d, max_iter):
    strides = [5, 6, 5, 469, 8, 9]
    with open(path, "wb") as f:
        n = [len(d["factors"].get(k, (np.zeros((0, strides[k])),))[0]) for k in range(6)]
        np.array([len(d["poses"]), len(d["vec3"]), len(d["rho"])] + n + [max_iter], dtype=np.int32).tofile(f)
        np.asarray(d["cameras"], dtype=np.float64).tofile(f)
        for a in (d["poses"], d["vec3"], d["rho"]):
            np.ascontiguousarray(a, dtype=np.float64).tofile(f)
        for k in range(6):
            if n[k]:
                np.ascontiguousarray(d["factors"][k][0], dtype=np.float64).tofile(f)
                np.ascontiguousarray(d["factors"][k][1], dtype=np.int32).tofile(f)


def load_ba_result(path, d):
    a = np.fromfile(path, dtype=np.float64)
    np_, nv, nr = len(d["poses"]), len(d["vec3"]), len(d["rho"])
    P = a[:7 * np_].reshape(np_, 7); a = a[7 * np_:]
    V = a[:3 * nv].reshape(nv, 3); a = a[3 * nv:]
    R = a[:nr]; s = a[nr:]
    return P, V, R, dict(initial_cost=s[0], final_cost=s[1], num_successful_steps=int(s[2]), num_residual_blocks=int(s[3]))


def dump_icp(path, sc, rpyxyz):
    with open(path, "wb") as f:
        np.array([len(sc["scan"]), len(sc["map"]), sc["mode"], sc["n_features_left"]], dtype=np.int32).tofile(f)
        for a in (sc["frame_pose"], sc["map_pose"], rpyxyz):
            np.ascontiguousarray(a, dtype=np.float64).tofile(f)
        np.ascontiguousarray(sc["scan"][:, :4], dtype=np.float32).tofile(f)
        np.ascontiguousarray(sc["map"][:, :4], dtype=np.float32).tofile(f)
