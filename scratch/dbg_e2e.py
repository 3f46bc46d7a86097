import time, numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lvio_fusion_b200 import _capi, backend, synth
lvb=_capi.load(); ctx=backend.Context(lvb)
d=synth.make_ba_problem(10, 4000, with_imu=True)
opt = backend.default_options(lvb, max_num_iterations=10, function_tolerance=0.0, gradient_tolerance=0.0, parameter_tolerance=0.0)
for rep in range(4):
    t=[time.perf_counter()]
    p=backend.Problem(ctx); p.set_cameras(d["cameras"]); p.set_poses(d["poses"]); p.set_vec3(d["vec3"]); p.set_inv_depths(d["rho"]); t.append(time.perf_counter())
    for k in range(6):
        f=d["factors"].get(k)
        if f is not None and len(f[0]): p.add_factors(k,f[0],f[1])
    for k,a in d["loss"].items(): p.set_loss(k,a)
    t.append(time.perf_counter())
    p.finalize(); t.append(time.perf_counter())
    s=p.solve(opt); t.append(time.perf_counter())
    P=p.poses(); V=p.vec3(); R=p.inv_depths(); t.append(time.perf_counter())
    p.close(); t.append(time.perf_counter())
    print("rep",rep," ".join("%s %.0fus"%(n,(b-a)*1e6) for n,a,b in zip(["params","add","finalize","solve","download","close"],t[:-1],t[1:])), "iters", s.num_iterations)
# ICP breakdown
sc = synth.make_icp_problem(120000, 1000000, kind="surf")
fa = backend.FeatureAssociation(ctx)
e0 = synth.relative_rpyxyz(sc["map_pose"], sc["frame_pose"])
for rep in range(3):
    t0=time.perf_counter(); fa.set_map(sc["map"], sc["cell_size"]); t1=time.perf_counter()
    e,s = fa.scan_to_map(sc["mode"], sc["scan"], sc["frame_pose"], sc["map_pose"], e0, sc["weight"], -1.0, sc["huber_a"], sc["thr"]); t2=time.perf_counter()
    idx,d2 = fa.knn3(sc["scan"], sc["frame_pose"], sc["cell_size"]**2); t3=time.perf_counter()
    print("icp set_map %.0fus scan_to_map %.0fus (iters %d, blocks %d) knn3(with D2H) %.0fus"%((t1-t0)*1e6,(t2-t1)*1e6,s.num_iterations,s.num_residual_blocks,(t3-t2)*1e6))
