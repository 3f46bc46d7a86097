import sys, time, numpy as np
from lvio_fusion_b200 import synth, backend, _capi
ctx = backend.Context(_capi.load())
for (n_kf, n_lm, imu) in [(200, 20000, True), (1000, 100000, True), (5000, 500000, False), (5000, 500000, True)]:
    t = time.time(); d = synth.make_ba_problem(n_kf, n_lm, with_imu=imu, seed=3); tg = time.time() - t
    t = time.time(); p = backend.Problem.from_dict(ctx, d); tf = time.time() - t
    try:
        t = time.time(); s = p.solve(max_num_iterations=10, function_tolerance=0, gradient_tolerance=0, parameter_tolerance=0); ts = time.time() - t
        print(n_kf, n_lm, imu, "dims", p.dims(), "gen %.2f finalize %.2f solve %.3f s" % (tg, tf, ts), "iters", s.num_iterations, "cost %.4g -> %.4g" % (s.initial_cost, s.final_cost), "term", s.termination_type, flush=True)
        err = np.abs(p.poses()[:, 4:] - d["poses_true"][:, 4:]).max(); err0 = np.abs(d["poses"][:, 4:] - d["poses_true"][:, 4:]).max()
        print("   pos err %.4f -> %.4f" % (err0, err), flush=True)
    except RuntimeError as e:
        print(n_kf, n_lm, imu, "ERR", e, flush=True)
