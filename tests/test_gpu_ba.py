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


def test_staged_uploads_match_per_array_uploads(lvb, c1):
    """finalize assembles its uploads in the context's pinned staging area (one copy, one allocation); LVB_NO_STAGING=1 keeps the
    per-array uploads.  Both must build the same device problem: identical evaluation, reduced system and solve."""
    ctx_a = backend.Context(lvb)
    os.environ["LVB_NO_STAGING"] = "1"
    try:
        # the switch is read once per process at the first finalize: run the unstaged arm in a child interpreter
        import subprocess, sys, json, tempfile
        code = ("import json,sys,numpy as np\n"
                "from lvio_fusion_b200 import _capi, backend, synth\n"
                "from lvio_fusion_b200.backend import TWO_FRAME, IMU\n"
                "sys.path.insert(0, %r)\n"
                "import test_gpu_ba as T\n"
                "d = synth.make_ba_problem(5, 2000, with_imu=True, seed=synth.SEED)\n"
                "d['poses'][:, :4] *= (1.0 + 0.003 * np.arange(5))[:, None]\n"
                "d = T._with_priors(d)\n"
                "ctx = backend.Context(_capi.load())\n"
                "p = backend.Problem.from_dict(ctx, d)\n"
                "r, J = p.evaluate(TWO_FRAME); ri, Ji = p.evaluate(IMU)\n"
                "s = p.solve(max_num_iterations=8)\n"
                "np.savez(sys.argv[1], r=r, J=J, ri=ri, Ji=Ji, P=p.poses(), V=p.vec3(), R=p.inv_depths(), cost=np.array([s.initial_cost, s.final_cost]))\n") % os.path.dirname(os.path.abspath(__file__))
        with tempfile.TemporaryDirectory() as td:
            out = os.path.join(td, "u.npz")
            subprocess.run([sys.executable, "-c", code, out], check=True, env=dict(os.environ), timeout=600)
            u = dict(np.load(out))
    finally:
        del os.environ["LVB_NO_STAGING"]
    p = backend.Problem.from_dict(ctx_a, c1)
    r, J = p.evaluate(TWO_FRAME)
    ri, Ji = p.evaluate(IMU)
    s = p.solve(max_num_iterations=8)
    assert np.array_equal(r, u["r"]) and np.array_equal(J, u["J"]) and np.array_equal(ri, u["ri"]) and np.array_equal(Ji, u["Ji"])
    assert abs(s.initial_cost - u["cost"][0]) <= 1e-12 * abs(u["cost"][0]) and abs(s.final_cost - u["cost"][1]) <= 1e-9 * abs(u["cost"][1])
    assert np.max(np.abs(p.poses() - u["P"])) < 1e-9 and np.max(np.abs(p.vec3() - u["V"])) < 1e-8 and np.max(np.abs(p.inv_depths() - u["R"])) < 1e-9


def test_getters_share_one_snapshot_and_follow_updates(lvb_ctx):
    """The three getters are served from one read-back; a solve or update_params must invalidate it."""
    d = synth.make_ba_problem(5, 500, with_imu=True, seed=11)
    p = backend.Problem.from_dict(lvb_ctx, d)
    assert np.array_equal(p.poses(), d["poses"]) and np.array_equal(p.vec3(), d["vec3"]) and np.array_equal(p.inv_depths(), d["rho"])
    p.solve(max_num_iterations=3)
    P1, V1, R1 = p.poses(), p.vec3(), p.inv_depths()
    assert np.max(np.abs(P1 - d["poses"])) > 0 and np.max(np.abs(R1 - d["rho"])) > 0
    assert np.array_equal(p.poses(), P1) and np.array_equal(p.vec3(), V1)
    p.update_params(d["poses"], d["vec3"], d["rho"])
    assert np.array_equal(p.poses(), d["poses"]) and np.array_equal(p.inv_depths(), d["rho"])
    # several problems on one context take turns in the staging area; the earlier one must stay intact
    q = backend.Problem.from_dict(lvb_ctx, synth.make_ba_problem(4, 300, with_imu=False, seed=12))
    q.solve(max_num_iterations=2)
    p.solve(max_num_iterations=3)
    assert np.max(np.abs(p.poses() - P1)) < 1e-9 and np.max(np.abs(p.inv_depths() - R1)) < 1e-9


def test_loop_closure_envelope_is_solved(lvb_ctx, orc_ctx):
    """A relative-pose constraint between the first and the last keyframe of a map-sized problem (Relocator / PoseGraph output fed
    back into a global BA) makes the envelope of the reduced camera system as wide as the system: no separator tree, no shared-memory
    panel.  The solve goes through the multi-grid fallback (ba_wide.cuh) and matches the dense oracle."""
    n_kf = 80
    d = synth.make_ba_problem(n_kf, 3000, with_imu=True, seed=13)
    d = dict(d); d["factors"] = dict(d["factors"])
    P = d["poses_true"]
    r = synth.relative_rpyxyz(P[0], P[n_kf - 1])      # PoseGraphError's measurement: rpyxyz of T_0^-1 T_last (pose_error.hpp:25-39)
    d["factors"][POSE_GRAPH] = (np.concatenate([r, [50.0, 0.5]])[None], np.array([[0, n_kf - 1]], dtype=np.int32))
    pg, po = backend.Problem.from_dict(lvb_ctx, d), backend.Problem.from_dict(orc_ctx, d)
    assert pg.dims() == po.dims() and pg.dims()[0] > 736
    Sg, bg, cg = pg.reduced_system(1e4)
    So, bo, co = po.reduced_system(1e4)
    assert abs(cg - co) < 1e-10 * co and _rel(Sg, So) < 1e-9 and _rel(bg, bo) < 1e-9
    sg = pg.solve(max_num_iterations=8)
    so = po.solve(max_num_iterations=8, num_threads=4)
    assert sg.termination_type == so.termination_type and sg.num_iterations == so.num_iterations
    assert abs(sg.final_cost - so.final_cost) < 1e-6 * so.final_cost and sg.final_cost < 0.5 * sg.initial_cost
    assert np.max(np.abs(pg.poses() - po.poses())) < 1e-5
