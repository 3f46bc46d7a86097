import ctypes, numpy as np, sys
sys.path.insert(0,'/root/repo')
from lvio_fusion_b200 import _capi, backend, synth
lvb=_capi.load(); ctx=backend.Context(lvb)
for (nk, nl) in ((10,4000),(20,8000)):
    d=synth.make_ba_problem(nk, nl, with_imu=True)
    p=backend.Problem.from_dict(ctx,d)
    out=(ctypes.c_longlong*8)()
    p.solve(max_num_iterations=5)
    lvb.lib.lvb_debug_cholesky_clocks(out,1)
    p.update_params(d['poses'],d['vec3'],d['rho'])
    s=p.solve(max_num_iterations=10, function_tolerance=0.0, gradient_tolerance=0.0, parameter_tolerance=0.0)
    lvb.lib.lvb_debug_cholesky_clocks(out,1)
    n=out[5]
    print("dimc", p.dims()[0], "calls", n, "per-call clocks: diag %.0f panel %.0f trail %.0f backward %.0f total %.0f"%tuple(out[i]/n for i in range(5)))
