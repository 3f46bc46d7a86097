"""Committed golden fixtures (tests/golden/*.npz, produced by tests/golden/make_golden.py from the oracle):
the oracle must keep reproducing them on CPU, the CUDA path on the GPU."""
import os
import sys

import numpy as np
import pytest

from lvio_fusion_b200 import backend, synth

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden  # noqa: E402


def _check_ba(ctx, jtol, stol):
    g = np.load(os.path.join(HERE, "golden", "ba_small.npz"))
    p = backend.Problem.from_dict(ctx, make_golden.ba_case())
    for k in range(6):
        r, J = p.evaluate(k)
        scale = max(1.0, np.abs(g["r%d" % k]).max())
        assert np.max(np.abs(r - g["r%d" % k])) < jtol * scale, k
        assert np.max(np.abs(J - g["J%d" % k])) < jtol * 10 * max(1.0, np.abs(g["J%d" % k]).max()), k
    S, b, cost = p.reduced_system(1e4)
    assert abs(cost - g["cost"]) < 1e-10 * g["cost"]
    assert np.max(np.abs(S - g["S"])) < 1e-9 * np.abs(g["S"]).max()
    assert np.max(np.abs(b - g["b"])) < 1e-9 * np.abs(g["b"]).max()
    s = p.solve(max_num_iterations=12)
    assert s.num_iterations == int(g["iterations"])
    assert abs(s.final_cost - g["final_cost"]) < stol * g["final_cost"]
    assert np.max(np.abs(p.poses() - g["poses"])) < 1e-6
    assert np.max(np.abs(p.inv_depths() - g["rho"])) < 1e-6


def _check_icp(ctx, kind, brute):
    g = np.load(os.path.join(HERE, "golden", "icp_%s.npz" % kind))
    sc = make_golden.icp_case(kind)
    fa = backend.FeatureAssociation(ctx)
    if brute:
        ctx.api.icp_set_brute(fa.h, 1)
    fa.set_map(sc["map"], sc["cell_size"])
    idx, d2 = fa.knn3(sc["scan"], sc["frame_pose"], sc["cell_size"] ** 2)
    assert np.array_equal(idx, g["idx"]) and np.array_equal(d2.view(np.uint32), g["d2"].view(np.uint32))
    acc, r, J = fa.evaluate(sc["mode"], sc["scan"], sc["frame_pose"], sc["map_pose"], g["e0"], sc["weight"], sc["thr"])
    assert np.array_equal(acc, g["acc"])
    assert np.max(np.abs(r - g["r"])) < 1e-9 * max(1.0, np.abs(g["r"]).max())
    assert np.max(np.abs(J - g["J"])) < 1e-9 * max(1.0, np.abs(g["J"]).max())
    e, s = fa.scan_to_map(sc["mode"], sc["scan"], sc["frame_pose"], sc["map_pose"], g["e0"], sc["weight"], -1.0, sc["huber_a"], sc["thr"])
    assert s.num_residual_blocks == int(g["blocks"])
    assert np.max(np.abs(e - g["e"])) < 1e-7


def test_oracle_reproduces_golden_ba(orc_ctx):
    _check_ba(orc_ctx, 1e-12, 1e-9)


@pytest.mark.parametrize("kind", ["ground", "surf"])
def test_oracle_reproduces_golden_icp(orc_ctx, kind):
    _check_icp(orc_ctx, kind, brute=False)       # the kd-tree path must agree with the brute-force fixtures


@pytest.mark.gpu
def test_cuda_reproduces_golden_ba(lvb_ctx):
    _check_ba(lvb_ctx, 1e-9, 1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["ground", "surf"])
def test_cuda_reproduces_golden_icp(lvb_ctx, kind):
    _check_icp(lvb_ctx, kind, brute=False)


def _check_lidar(ctx):
    g = np.load(os.path.join(HERE, "golden", "lidar_small.npz"))
    lf = backend.LidarFeatures(ctx, horizon_scan=450, extrinsic=[0.0, 0.0, 0.0, 1.0, 0.27, 0.0, 0.08])
    scan = make_golden.lidar_case()
    seg = lf.segment(scan)
    for k, gk in (("points", "seg_points"), ("range", "seg_range"), ("ground", "seg_ground"), ("col", "seg_col"), ("curvature", "seg_curvature"),
                  ("start_ring", "start_ring"), ("end_ring", "end_ring"), ("orientation", "orientation")):
        assert np.array_equal(seg[k], g[gk]), k                       # float32 / integer path: bit-exact
    ground, surf = lf.extract(scan)
    assert len(g["ground"]) > 100 and len(g["surf"]) > 100
    assert np.array_equal(ground, g["ground"]) and np.array_equal(surf, g["surf"])


def _check_imu(ctx, tol):
    g = np.load(os.path.join(HERE, "golden", "imu_small.npz"))["consts"]
    c = backend.preintegrate(ctx, *make_golden.imu_case())
    assert c.shape == g.shape
    for lo, hi in ((0, 17), (17, 242), (242, 467), (467, 469)):
        assert np.max(np.abs(c[:, lo:hi] - g[:, lo:hi])) <= tol * max(1e-300, np.abs(g[:, lo:hi]).max())


def test_oracle_reproduces_golden_lidar_and_imu(orc_ctx):
    _check_lidar(orc_ctx)
    _check_imu(orc_ctx, 1e-15)


@pytest.mark.gpu
def test_cuda_reproduces_golden_lidar_and_imu(lvb_ctx):
    _check_lidar(lvb_ctx)
    _check_imu(lvb_ctx, 1e-12)
