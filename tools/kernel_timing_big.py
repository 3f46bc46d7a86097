import ctypes, numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lvio_fusion_b200 import _capi, backend, synth
lvb=_capi.load(); ctx=backend.Context(lvb)
for (nk, nl, imu) in [(5000, 500000, False), (5000, 500000, True), (20, 8000, True)]:
    d=synth.make_ba_problem(nk, nl, with_imu=imu)
    p=backend.Problem.from_dict(ctx,d)
    p.solve(max_num_iterations=3)
    p.update_params(d['poses'],d['vec3'],d['rho'])
    print("=====", nk, nl, imu, p.dims(), flush=True)
    lvb.lib.lvb_debug_timing(1)
    p.solve(max_num_iterations=2, function_tolerance=0.0, gradient_tolerance=0.0, parameter_tolerance=0.0)
    lvb.lib.lvb_debug_timing(0)
    sys.stdout.flush()
