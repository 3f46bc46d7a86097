"""GPU parity tests of the bundle-adjustment path: CUDA kernels (through the C ABI) vs the CPU oracle
on the same seeded inputs.  Tolerances (FP64): residual 1e-10 relative to the largest residual of the
kind, Jacobian 1e-9 relative (whitened IMU: 1e-8, the 15x15 covariance inverse amplifies rounding),
reduced system 1e-9, post-solve pose 1e-6 m / 1e-7 rad-equivalent."""
import os

import numpy as np
import pytest

from lvio_fusion_b200 import backend, synth
from lvio_fusion_b200.backend import IMU, POSE_GRAPH, POSE_ONLY, POSE_PRIOR, TWO_CAMERA, TWO_FRAME

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float(np.max(np.abs(a - b)) / max(1.0, np.max(np.abs(b)))) if a.size else 0.0


def _with_priors(d):
    P = d["poses"]
    rng = np.random.default_rng(0)
    d = dict(d)
    d["factors"] = dict(d["factors"])
    d["factors"][POSE_GRAPH] = (np.concatenate([rng.normal(0, 0.05, 6) + [0, 0, 0, 1, 0, 0], [100.0, 0.3]])[None], np.array([[0, 1]], dtype=np.int32))
    d["factors"][POSE_PRIOR] = (np.concatenate([P[2] + rng.normal(0, 0.01, 7), [100.0, 0.5]])[None], np.array([[2]], dtype=np.int32))
    return d


@pytest.fixture(scope="module")
def c1():
    d = synth.make_ba_problem(5, 2000, with_imu=True, seed=synth.SEED)
    d["poses"][:, :4] *= (1.0 + 0.003 * np.arange(5))[:, None]   # stored quaternions need not be unit
    return _with_priors(d)


@pytest.mark.parametrize("kind,rtol,jtol", [(TWO_FRAME, 1e-10, 1e-9), (POSE_ONLY, 1e-10, 1e-9), (TWO_CAMERA, 1e-10, 1e-9),
                                            (IMU, 1e-8, 1e-8), (POSE_GRAPH, 1e-10, 1e-9), (POSE_PRIOR, 1e-10, 1e-9)])
def test_eval_matches_oracle(lvb_ctx, orc_ctx, c1, kind, rtol, jtol):
    pg, po = backend.Problem.from_dict(lvb_ctx, c1), backend.Problem.from_dict(orc_ctx, c1)
    rg, Jg = pg.evaluate(kind)
    ro, Jo = po.evaluate(kind)
    assert rg.shape == ro.shape and len(rg) > 0
    assert _rel(rg, ro) < rtol
    assert _rel(Jg, Jo) < jtol


def test_eval_without_tma_staging(lvb, orc_ctx, c1):
    os.environ["LVB_NO_TMA"] = "1"
    try:
        ctx = backend.Context(lvb)
    finally:
        del os.environ["LVB_NO_TMA"]
    pg, po = backend.Problem.from_dict(ctx, c1), backend.Problem.from_dict(orc_ctx, c1)
    rg, Jg = pg.evaluate(TWO_FRAME)
    ro, Jo = po.evaluate(TWO_FRAME)
    assert _rel(rg, ro) < 1e-10 and _rel(Jg, Jo) < 1e-9


@pytest.mark.parametrize("with_const", [False, True])
def test_reduced_system_matches_oracle(lvb_ctx, orc_ctx, c1, with_const):
    d = dict(c1)
    if with_const:
        d["pose_const"] = np.array([1, 0, 0, 0, 0], dtype=np.uint8)
        rc = np.zeros(len(d["rho"]), dtype=np.uint8); rc[::7] = 1
        d["rho_const"] = rc
        vc = np.zeros(len(d["vec3"]), dtype=np.uint8); vc[1] = 1
        d["vec3_const"] = vc
    pg, po = backend.Problem.from_dict(lvb_ctx, d), backend.Problem.from_dict(orc_ctx, d)
    assert pg.dims() == po.dims()
    Sg, bg, cg = pg.reduced_system(1e4)
    So, bo, co = po.reduced_system(1e4)
    assert abs(cg - co) < 1e-10 * co
    assert _rel(Sg, So) < 1e-9
    assert _rel(bg, bo) < 1e-9
    # the solution of the reduced system agrees too (exercises conditioning)
    xg, xo = np.linalg.solve(Sg, bg), np.linalg.solve(So, bo)
    assert np.max(np.abs(xg - xo)) < 1e-7 * max(1.0, np.abs(xo).max())


@pytest.mark.parametrize("n_kf,n_lm,imu", [(5, 2000, False), (10, 4000, True)])
def test_solve_matches_oracle(lvb_ctx, orc_ctx, n_kf, n_lm, imu):
    d = synth.make_ba_problem(n_kf, n_lm, with_imu=imu, seed=7)
    if not imu:
        d = dict(d); d["factors"] = dict(d["factors"])
        d["factors"][POSE_PRIOR] = (np.concatenate([d["poses"][0], [100.0, 0.0]])[None], np.zeros((1, 1), dtype=np.int32))
    pg, po = backend.Problem.from_dict(lvb_ctx, d), backend.Problem.from_dict(orc_ctx, d)
    sg = pg.solve(max_num_iterations=50)
    so = po.solve(max_num_iterations=50, num_threads=4)
    assert sg.termination_type == so.termination_type == 0
    assert abs(sg.initial_cost - so.initial_cost) < 1e-9 * so.initial_cost
    assert abs(sg.final_cost - so.final_cost) < 1e-6 * so.final_cost
    assert sg.final_cost < 0.2 * sg.initial_cost
    Pg, Po = pg.poses(), po.poses()
    assert np.max(np.abs(Pg[:, 4:] - Po[:, 4:])) < 1e-6
    assert np.max(np.abs(Pg[:, :4] - Po[:, :4])) < 1e-7
    assert np.max(np.abs(pg.inv_depths() - po.inv_depths())) < 1e-6
    if imu:
        assert np.max(np.abs(pg.vec3() - po.vec3())) < 1e-5
    # the solve actually moved towards the truth
    assert np.abs(Pg[:, 4:] - d["poses_true"][:, 4:]).max() < np.abs(d["poses"][:, 4:] - d["poses_true"][:, 4:]).max()


def test_single_iteration_and_time_cap(lvb_ctx, orc_ctx):
    """UpdateFrontend runs max_num_iterations = 1 (backend.cpp:264); Backend::Optimize caps wall time (:208)."""
    d = synth.make_ba_problem(5, 500, with_imu=True, seed=9)
    pg, po = backend.Problem.from_dict(lvb_ctx, d), backend.Problem.from_dict(orc_ctx, d)
    sg, so = pg.solve(max_num_iterations=1), po.solve(max_num_iterations=1)
    assert sg.num_iterations == so.num_iterations == 1
    assert abs(sg.final_cost - so.final_cost) < 1e-8 * so.final_cost
    assert np.max(np.abs(pg.poses() - po.poses())) < 1e-8
    p2 = backend.Problem.from_dict(lvb_ctx, d)
    s2 = p2.solve(max_num_iterations=50, max_solver_time_in_seconds=0.0)
    assert s2.num_iterations <= 1


def test_reprojection_errors(lvb_ctx, orc_ctx, c1):
    pg, po = backend.Problem.from_dict(lvb_ctx, c1), backend.Problem.from_dict(orc_ctx, c1)
    c, ix = c1["factors"][POSE_ONLY]
    ob_pw = c[:, :5]
    eg, eo = pg.reprojection_errors(ob_pw, ix[:, 0]), po.reprojection_errors(ob_pw, ix[:, 0])
    assert np.max(np.abs(eg - eo)) < 1e-9 * max(1.0, eo.max())


def test_errors_are_loud(lvb_ctx):
    p = backend.Problem(lvb_ctx)
    with pytest.raises(RuntimeError):
        p.solve()                      # not finalized
    d = synth.make_ba_problem(3, 20, with_imu=False, seed=1)
    d["factors"][TWO_FRAME][1][0, 1] = 99   # pose index out of range
    with pytest.raises(RuntimeError):
        backend.Problem.from_dict(lvb_ctx, d)


def test_full_ba_with_imu_init_error(lvb_ctx, orc_ctx):
    """imu::FullBA (tools.cpp:92-171): ImuInitError factors (imu_error.hpp:124-229) sharing one ba and one bg block."""
    d = synth.make_fullba_problem(6, 600, seed=23)
    pg, po = backend.Problem.from_dict(lvb_ctx, d), backend.Problem.from_dict(orc_ctx, d)
    rg, Jg = pg.evaluate(IMU)
    ro, Jo = po.evaluate(IMU)
    assert _rel(rg, ro) < 1e-8 and _rel(Jg, Jo) < 1e-8
    assert np.all(Jo[:, :, 26:] == 0) and np.all(Jg[:, :, 26:] == 0)        # no ba_j / bg_j blocks
    Sg, bg, cg = pg.reduced_system(1e4)
    So, bo, co = po.reduced_system(1e4)
    assert abs(cg - co) < 1e-10 * co and _rel(Sg, So) < 1e-9 and _rel(bg, bo) < 1e-9
    sg, so = pg.solve(max_num_iterations=30), po.solve(max_num_iterations=30, num_threads=2)
    assert abs(sg.final_cost - so.final_cost) < 1e-6 * so.final_cost
    assert np.max(np.abs(pg.poses() - po.poses())) < 1e-6
    assert np.max(np.abs(pg.vec3() - po.vec3())) < 1e-5


@pytest.mark.parametrize("n_kf,n_lm,imu", [(80, 3000, True), (150, 3000, False)])
def test_banded_storage_matches_oracle(lvb_ctx, orc_ctx, n_kf, n_lm, imu):
    """Camera systems above the dense cap (736) use banded storage of H_pp / S and the envelope Cholesky; the oracle
    stays dense.  Covers the map-scale BA of backend.cpp:343-370 (imu::FullBA / global BA over all keyframes)."""
    d = synth.make_ba_problem(n_kf, n_lm, with_imu=imu, seed=11)
    if not imu:
        d = dict(d); d["factors"] = dict(d["factors"])
        d["factors"][POSE_PRIOR] = (np.concatenate([d["poses"][0], [100.0, 0.0]])[None], np.zeros((1, 1), dtype=np.int32))
    pg, po = backend.Problem.from_dict(lvb_ctx, d), backend.Problem.from_dict(orc_ctx, d)
    assert pg.dims() == po.dims() and pg.dims()[0] > 736
    Sg, bg, cg = pg.reduced_system(1e4)
    So, bo, co = po.reduced_system(1e4)
    assert abs(cg - co) < 1e-10 * co
    assert _rel(Sg, So) < 1e-9 and _rel(bg, bo) < 1e-9
    sg = pg.solve(max_num_iterations=12)
    so = po.solve(max_num_iterations=12, num_threads=4)
    assert sg.termination_type == so.termination_type
    assert sg.num_iterations == so.num_iterations
    assert abs(sg.final_cost - so.final_cost) < 1e-6 * so.final_cost
    assert sg.final_cost < 0.2 * sg.initial_cost
    Pg, Po = pg.poses(), po.poses()
    assert np.max(np.abs(Pg[:, 4:] - Po[:, 4:])) < 1e-5
    assert np.max(np.abs(Pg[:, :4] - Po[:, :4])) < 1e-6
    assert np.max(np.abs(pg.inv_depths() - po.inv_depths())) < 1e-5
