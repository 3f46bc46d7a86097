import time, numpy as np, sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lvio_fusion_b200 import _capi, backend, synth
import torch
lvb=_capi.load(); ctx=backend.Context(lvb)
for (nk,nl) in ((10,4000),(20,8000)):
    d=synth.make_ba_problem(nk,nl,with_imu=True)
    for mode in (0,1):
        p=backend.Problem.from_dict(ctx,d)
        o=backend.default_options(lvb, max_num_iterations=10, function_tolerance=0.0, gradient_tolerance=0.0, parameter_tolerance=0.0, schur_mode=mode)
        for _ in range(3):
            p.update_params(d['poses'],d['vec3'],d['rho']); s=p.solve(o)
        torch.cuda.synchronize(); t=time.perf_counter(); n=0
        for _ in range(5):
            p.update_params(d['poses'],d['vec3'],d['rho']); s=p.solve(o); n+=s.num_iterations
        torch.cuda.synchronize(); dt=time.perf_counter()-t
        print("kf %d lm %d schur_mode %d: %.1f us / LM iteration (final cost %.6e, iters %d)"%(nk,nl,mode,dt/n*1e6,s.final_cost,s.num_iterations))
if len(sys.argv)>1:
    lvb.lib.lvb_debug_timing(1)
    p.update_params(d['poses'],d['vec3'],d['rho']); p.solve(backend.default_options(lvb, max_num_iterations=1, schur_mode=1))
    lvb.lib.lvb_debug_timing(0)
