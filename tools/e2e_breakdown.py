"""Host-side breakdown of the e2e window solve (Problem.from_dict + solve + download + close), wall clock per stage."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lvio_fusion_b200 import _capi, backend, synth
import bench

lvb = _capi.load(); ctx = backend.Context(lvb)
for nk, nl in ((10, 4000), (20, 8000)):
    d = synth.make_ba_problem(nk, nl, with_imu=True, seed=synth.SEED)
    opts = bench.bench_options(lvb, 10) if hasattr(bench, "bench_options") else None
    acc = {}
    def tick(name, t0):
        t = time.perf_counter(); acc.setdefault(name, []).append((t - t0) * 1e3); return t
    for rep in range(int(os.environ.get("REPS", "60"))):
        t = time.perf_counter(); t_all = t
        p = backend.Problem(ctx)
        p.set_cameras(d["cameras"]); p.set_poses(d["poses"], d.get("pose_const")); p.set_vec3(d["vec3"], d.get("vec3_const")); p.set_inv_depths(d["rho"], d.get("rho_const"))
        t = tick("create+set params", t)
        for kind in range(6):
            f = d["factors"].get(kind)
            if f is not None and len(f[0]): p.add_factors(kind, f[0], f[1])
        for kind, a in d.get("loss", {}).items(): p.set_loss(kind, a)
        t = tick("add_factors", t)
        p.finalize()
        t = tick("finalize", t)
        s = p.solve(opts) if opts is not None else p.solve(max_num_iterations=10)
        t = tick("solve", t)
        p.poses(); p.vec3(); p.inv_depths()
        t = tick("download", t)
        p.close()
        t = tick("close", t)
        acc.setdefault("total", []).append((t - t_all) * 1e3)
    print("window %d keyframes / %d landmarks, %d iterations:" % (nk, nl, s.num_iterations))
    for k, v in acc.items(): print("  %-20s median %.3f ms   min %.3f" % (k, float(np.median(v[min(10, len(v) - 1):])), float(np.min(v[min(10, len(v) - 1):]))))
