"""ctypes binding of the CPU oracle (oracle/liboracle.so) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  It mirrors the product ABI with the prefix ``orc_`` plus a few producers
(IMU preintegration, SE3 helpers) used to build synthetic inputs.
"""
import ctypes as C
import os
import subprocess

from lvio_fusion_b200 import _capi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")

_dp = _capi.c_double_p
_EXTRA = {
    "ba_set_threads": (C.c_int, [C.c_void_p, C.c_int]),
    "ba_iterate": (C.c_int, [C.c_void_p, C.c_int, C.c_double, _dp]),
    "icp_set_threads": (C.c_int, [C.c_void_p, C.c_int]),
    "icp_set_brute": (C.c_int, [C.c_void_p, C.c_int]),
    "preintegrate": (C.c_int, [C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _dp]),
    "sqrt_information": (C.c_int, [_dp, _dp]),
    "pose_plus": (C.c_int, [_dp, _dp, _dp]),
    "pose_tangent": (C.c_int, [_dp, C.c_int, C.c_int, _dp, _dp]),
    "se3_to_rpyxyz": (C.c_int, [_dp, _dp]),
    "rpyxyz_to_se3": (C.c_int, [_dp, _dp]),
    "se3_compose": (C.c_int, [_dp, _dp, _dp]),
    "se3_inverse": (C.c_int, [_dp, _dp]),
    "transform_f32": (C.c_int, [_dp, C.c_int, _capi.c_float_p, _capi.c_float_p]),
}

_api = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def load():
    global _api
    if _api is None:
        if not os.path.exists(LIB_PATH):
            build()
        _api = _capi.Api(C.CDLL(LIB_PATH), "orc", _EXTRA)
    return _api
