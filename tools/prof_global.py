"""ncu target: two LM iterations of the configs[4]-scale map BA (5000 keyframes, 500k landmarks) with direct launches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["LVB_NO_GRAPH"] = "1"
from lvio_fusion_b200 import _capi, backend, synth
lvb = _capi.load(); ctx = backend.Context(lvb)
d = synth.make_ba_problem(5000, 500000, with_imu=True, seed=synth.SEED + 1)
p = backend.Problem.from_dict(ctx, d)
p.solve(max_num_iterations=2, function_tolerance=0.0, gradient_tolerance=0.0, parameter_tolerance=0.0)
