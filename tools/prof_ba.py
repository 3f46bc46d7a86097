import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lvio_fusion_b200 import _capi, backend, synth
lvb = _capi.load(); ctx = backend.Context(lvb)
d = synth.make_ba_problem(10, 4000, with_imu=True)
for mode in (0, 1):
    p = backend.Problem.from_dict(ctx, d)
    o = backend.default_options(lvb, max_num_iterations=3, function_tolerance=0.0, gradient_tolerance=0.0, parameter_tolerance=0.0, schur_mode=mode)
    p.solve(o)
