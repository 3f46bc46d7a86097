"""Generate tests/golden/*.npz from the CPU oracle.

The reference ships no golden vectors and cannot be built or imported in this environment, so
these fixtures pin the *oracle*: they are regenerated only when its arithmetic is deliberately changed, and both
the oracle (tests/test_golden.py, CPU) and the CUDA path (tests/test_gpu_golden.py) must keep reproducing them.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from lvio_fusion_b200 import backend, synth  # noqa: E402
from oracle import binding  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def ba_case():
    d = synth.make_ba_problem(4, 40, with_imu=True, seed=101)
    d["poses"][:, :4] *= 1.002
    d["factors"] = dict(d["factors"])
    d["factors"][4] = (np.array([[0.01, -0.02, 0.005, 1.0, 0.02, -0.01, 100.0, 0.3]]), np.array([[0, 1]], dtype=np.int32))
    d["factors"][5] = (np.concatenate([d["poses"][2] * 1.0, [100.0, 0.5]])[None], np.array([[2]], dtype=np.int32))
    return d


def icp_case(kind):
    return synth.make_icp_problem(300, 4000, seed=102, kind=kind)


def lidar_case():
    return synth.make_lidar_scan(seed=103, horizon_scan=450, n_boxes=8)


def imu_case():
    rng = np.random.default_rng(104)
    counts = np.array([0, 7, 12, 1, 9])
    first = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    m = int(first[-1])
    samples = np.concatenate([rng.uniform(0.004, 0.012, (m, 1)), rng.normal(0, 1.5, (m, 3)) + [0, 0, 9.81], rng.normal(0, 0.3, (m, 3))], axis=1)
    n = len(counts)
    return (first, samples, rng.normal(0, 1.5, (n, 3)) + [0, 0, 9.81], rng.normal(0, 0.3, (n, 3)), rng.normal(0, 0.05, (n, 3)), rng.normal(0, 0.01, (n, 3)),
            np.array(synth.IMU_NOISE, dtype=np.float64))


def main():
    orc = binding.load()
    ctx = backend.Context(orc)
    d = ba_case()
    p = backend.Problem.from_dict(ctx, d)
    out = {}
    for k in range(6):
        r, J = p.evaluate(k)
        out["r%d" % k], out["J%d" % k] = r, J
    S, b, cost = p.reduced_system(1e4)
    out.update(S=S, b=b, cost=np.array(cost))
    s = p.solve(max_num_iterations=12)
    out.update(poses=p.poses(), vec3=p.vec3(), rho=p.inv_depths(), final_cost=np.array(s.final_cost), iterations=np.array(s.num_iterations))
    np.savez_compressed(os.path.join(HERE, "ba_small.npz"), **out)
    for kind in ("ground", "surf"):
        sc = icp_case(kind)
        fa = backend.FeatureAssociation(ctx)
        orc.icp_set_brute(fa.h, 1)
        fa.set_map(sc["map"], sc["cell_size"])
        idx, d2 = fa.knn3(sc["scan"], sc["frame_pose"], sc["cell_size"] ** 2)
        e0 = synth.relative_rpyxyz(sc["map_pose"], sc["frame_pose"])
        acc, r, J = fa.evaluate(sc["mode"], sc["scan"], sc["frame_pose"], sc["map_pose"], e0, sc["weight"], sc["thr"])
        e, s = fa.scan_to_map(sc["mode"], sc["scan"], sc["frame_pose"], sc["map_pose"], e0, sc["weight"], -1.0, sc["huber_a"], sc["thr"])
        np.savez_compressed(os.path.join(HERE, "icp_%s.npz" % kind), idx=idx, d2=d2, acc=acc, r=r, J=J, e0=e0, e=e,
                            final_cost=np.array(s.final_cost), blocks=np.array(s.num_residual_blocks))
    lf = backend.LidarFeatures(ctx, horizon_scan=450, extrinsic=[0.0, 0.0, 0.0, 1.0, 0.27, 0.0, 0.08])
    scan = lidar_case()
    seg = lf.segment(scan)
    ground, surf = lf.extract(scan)
    np.savez_compressed(os.path.join(HERE, "lidar_small.npz"), seg_points=seg["points"], seg_range=seg["range"], seg_ground=seg["ground"], seg_col=seg["col"],
                        seg_curvature=seg["curvature"], start_ring=seg["start_ring"], end_ring=seg["end_ring"], orientation=seg["orientation"],
                        ground=ground, surf=surf)
    np.savez_compressed(os.path.join(HERE, "imu_small.npz"), consts=backend.preintegrate(ctx, *imu_case()))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
