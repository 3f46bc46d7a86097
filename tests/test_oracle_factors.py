"""CPU tests of the oracle itself: the reference ships no golden vectors, so the
oracle's fidelity is pinned by (1) literal restatement, (2) an independent second derivation of
every Jacobian (central differences here), (3) structural identities."""
import ctypes as C

import numpy as np
import pytest

from lvio_fusion_b200 import backend, synth
from lvio_fusion_b200.backend import IMU, POSE_GRAPH, POSE_ONLY, POSE_PRIOR, TWO_CAMERA, TWO_FRAME


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


@pytest.fixture(scope="module")
def small(orc_ctx):
    d = synth.make_ba_problem(4, 60, with_imu=True, seed=11)
    # add the two prior kinds so every kind is exercised
    P = d["poses"]
    tgt = np.zeros(6)
    from oracle import binding
    api = binding.load()
    rel = np.zeros(7); inv = np.zeros(7)
    api.se3_inverse(_dp(P[0].copy()), _dp(inv)); api.se3_compose(_dp(inv), _dp(P[1].copy()), _dp(rel)); api.se3_to_rpyxyz(_dp(rel), _dp(tgt))
    tgt += 0.01
    d["factors"][POSE_GRAPH] = (np.concatenate([tgt, [100.0, 0.7]])[None, :], np.array([[0, 1]], dtype=np.int32))
    pr = P[2].copy(); pr[4:] += 0.02
    d["factors"][POSE_PRIOR] = (np.concatenate([pr, [100.0, 0.5]])[None, :], np.array([[2]], dtype=np.int32))
    return d


def _eval_all(ctx, d, poses, vec3, rho, kind):
    p = backend.Problem.from_dict(ctx, d)
    p.update_params(poses, vec3, rho)
    return p.evaluate(kind)


# column -> (array name, block index position in idx, offset inside block)
_LAYOUT = {
    TWO_FRAME: [("rho", 0, 1), ("poses", 1, 7), ("poses", 2, 7)],
    POSE_ONLY: [("poses", 0, 7)],
    TWO_CAMERA: [("rho", 0, 1)],
    POSE_GRAPH: [("poses", 0, 7), ("poses", 1, 7)],
    POSE_PRIOR: [("poses", 0, 7)],
}


@pytest.mark.parametrize("kind", [TWO_FRAME, POSE_ONLY, TWO_CAMERA, POSE_GRAPH, POSE_PRIOR])
def test_autodiff_matches_central_differences(orc_ctx, small, kind):
    d = small
    consts, idx = d["factors"][kind]
    n = min(len(consts), 25)
    sub = dict(d); sub["factors"] = {kind: (consts[:n], idx[:n])}
    base = dict(poses=d["poses"].copy(), vec3=d["vec3"].copy(), rho=d["rho"].copy())
    r0, J = _eval_all(orc_ctx, sub, base["poses"], base["vec3"], base["rho"], kind)
    col = 0
    for name, pos, width in _LAYOUT[kind]:
        for k in range(width):
            h = 1e-6
            num = np.zeros_like(r0)
            for f in range(n):
                for sgn in (+1, -1):
                    arrs = {a: b.copy() for a, b in base.items()}
                    if name == "rho":
                        arrs["rho"][idx[f, pos]] += sgn * h * 1e-2
                    else:
                        arrs["poses"][idx[f, pos], k] += sgn * h
                    r, _ = _eval_all(orc_ctx, sub, arrs["poses"], arrs["vec3"], arrs["rho"], kind)
                    num[f] += sgn * r[f]
                num[f] /= 2 * h * (1e-2 if name == "rho" else 1.0)
            ana = J[:, :, col]
            scale = np.maximum(np.abs(ana).max(axis=1, keepdims=True), 1.0)  # FD noise scales with the block, not the row
            assert np.max(np.abs(num - ana) / scale) < 2e-5, (kind, name, k)
            col += 1


def test_quaternion_jacobian_is_tangential(orc_ctx, small):
    """ceres::QuaternionRotatePoint normalises q, so J_q . q == 0 for every pose block."""
    d = small
    r, J = _eval_all(orc_ctx, d, d["poses"], d["vec3"], d["rho"], TWO_FRAME)
    idx = d["factors"][TWO_FRAME][1]
    q1 = d["poses"][idx[:, 1], :4]; q2 = d["poses"][idx[:, 2], :4]
    rad1 = np.einsum("frk,fk->fr", J[:, :, 1:5], q1); rad2 = np.einsum("frk,fk->fr", J[:, :, 8:12], q2)
    assert np.max(np.abs(rad1)) < 1e-7 * np.abs(J).max() and np.max(np.abs(rad2)) < 1e-7 * np.abs(J).max()


def _right_perturb(q, dth):
    dq = np.concatenate([0.5 * dth, [1.0]])
    out = synth.quat_mul(q[None, :], dq[None, :])[0]
    return out


def test_imu_residual_and_jacobian_convention(orc_ctx, small):
    """ImuError's analytic blocks are derivatives under VINS' right-perturbation
    q (x) [1, dtheta/2], stuffed into columns 0..2 of the 15x7 block with column 3 == 0
    (imu_error.hpp:45-52).  Checked against central differences of the whitened residual."""
    d = small
    consts, idx = d["factors"][IMU]
    sub = dict(d); sub["factors"] = {IMU: (consts[:2], idx[:2])}
    P, V, R = d["poses"].copy(), d["vec3"].copy(), d["rho"].copy()
    r0, J = _eval_all(orc_ctx, sub, P, V, R, IMU)
    assert np.all(J[:, :, 3] == 0) and np.all(J[:, :, 16 + 3] == 0)
    h = 1e-6
    for f in range(2):
        cols = [(0, idx[f, 0], "pose"), (7, idx[f, 1], "v"), (10, idx[f, 2], "v"), (13, idx[f, 3], "v"),
                (16, idx[f, 4], "pose"), (23, idx[f, 5], "v"), (26, idx[f, 6], "v"), (29, idx[f, 7], "v")]
        for c0, blk, typ in cols:
            width = 6 if typ == "pose" else 3
            for k in range(width):
                num = np.zeros(15)
                for sgn in (+1, -1):
                    Pp, Vp = P.copy(), V.copy()
                    if typ == "v":
                        Vp[blk, k] += sgn * h
                    elif k < 3:
                        e = np.zeros(3); e[k] = sgn * h
                        Pp[blk, :4] = _right_perturb(P[blk, :4], e)
                    else:
                        Pp[blk, 4 + k - 3] += sgn * h
                    r, _ = _eval_all(orc_ctx, sub, Pp, Vp, R, IMU)
                    num += sgn * r[f]
                num /= 2 * h
                col = c0 + (k if (typ == "v" or k < 3) else k + 1)
                ana = J[f, :, col]
                tol = 2e-4 * max(1.0, np.abs(ana).max())
                # bias Jacobians are first-order in the reference (imu_error.hpp:62-80): loose tolerance
                assert np.max(np.abs(num - ana)) < (5e-2 * max(1.0, np.abs(ana).max()) if c0 in (10, 13) else tol), (f, c0, k)


def test_sqrt_information_factorises_inverse_covariance(orc, small):
    consts = small["factors"][IMU][0]
    cov = consts[0, 242:467].reshape(15, 15).copy()
    U = np.zeros((15, 15))
    assert orc.sqrt_information(_dp(cov), _dp(U)) == 0
    assert np.allclose(np.triu(U), U)
    lhs = U.T @ U @ cov
    assert np.max(np.abs(lhs - np.eye(15))) < 1e-6


def test_preintegration_producers_agree(orc, small):
    """synth.preintegrate_batch (numpy) vs the oracle's restatement of preintegration.cpp:30-127."""
    rng = np.random.default_rng(3)
    S = 10
    acc = rng.normal(0, 1, (1, S + 1, 3)) + np.array([0, 0, 9.8]); gyr = rng.normal(0, 0.1, (1, S + 1, 3))
    ba = rng.normal(0, 0.01, (1, 3)); bg = rng.normal(0, 0.001, (1, 3))
    a = synth.preintegrate_batch(0.01, acc, gyr, ba, bg)[0]
    samples = np.concatenate([np.full((S, 1), 0.01), acc[0, 1:], gyr[0, 1:]], axis=1).copy()
    out = np.zeros(469)
    orc.preintegrate(S, _dp(samples), _dp(acc[0, 0].copy()), _dp(gyr[0, 0].copy()), _dp(ba[0].copy()), _dp(bg[0].copy()),
                     _dp(synth.IMU_NOISE.copy()), _dp(out))
    assert np.allclose(a[:17], out[:17], rtol=0, atol=1e-7)
    assert np.allclose(a[17:242], out[17:242], rtol=1e-6, atol=1e-8)
    assert np.allclose(a[242:467], out[242:467], rtol=1e-5, atol=1e-14)


def test_pose_plus_is_left_multiplication(orc):
    x = np.array([0.1, -0.2, 0.3, 0.9, 1.0, 2.0, 3.0]); x[:4] /= np.linalg.norm(x[:4])
    d = np.array([0.01, -0.02, 0.03, 0.1, 0.2, 0.3])
    out = np.zeros(7)
    orc.pose_plus(_dp(x), _dp(d), _dp(out))
    n = np.linalg.norm(d[:3])
    dq = np.concatenate([np.sin(n) / n * d[:3], [np.cos(n)]])
    assert np.allclose(out[:4], synth.quat_mul(dq[None], x[None, :4])[0], atol=1e-15)
    assert np.allclose(out[4:], x[4:] + d[3:])


def test_imu_init_error_variant(orc_ctx):
    """ImuInitError (imu_error.hpp:124-229): Baj = Bgj = 0, priors replace the bias blocks of cov^-1."""
    d = synth.make_fullba_problem(4, 40, seed=3)
    p = backend.Problem.from_dict(orc_ctx, d)
    r, J = p.evaluate(IMU)
    assert np.all(J[:, :, 26:] == 0)
    # residual rows 9..14 are sqrt_info-mixed (-Bai, -Bgi); check through the un-whitened relation with U
    c = d["factors"][IMU][0][0]
    U = np.zeros((15, 15))
    cov_inv = np.linalg.inv(c[242:467].reshape(15, 15))
    cov_inv[9:12, 9:12] = c[467] * np.eye(3); cov_inv[12:15, 12:15] = c[468] * np.eye(3)
    L = np.linalg.cholesky(cov_inv)
    raw = np.linalg.solve(L.T, r[0])
    ba, bg = d["vec3"][4], d["vec3"][5]
    assert np.allclose(raw[9:12], -ba, atol=1e-9) and np.allclose(raw[12:15], -bg, atol=1e-9)
    s = p.solve(max_num_iterations=20)
    assert s.final_cost < s.initial_cost


def test_batched_preintegration_agrees_with_numpy_producer(orc_ctx):
    """Two independent producers of the LVB_IMU records: the oracle's C++ restatement (batch ABI form) and
    synth.preintegrate_batch (numpy).  They differ only by the O(|w dt|^2) un-normalised-quaternion detail noted there."""
    from lvio_fusion_b200 import backend, synth
    rng = np.random.default_rng(3)
    F, S = 12, 10
    acc = rng.normal(0, 1.0, (F, S + 1, 3)) + [0, 0, 9.81]
    gyr = rng.normal(0, 0.2, (F, S + 1, 3))
    ba = rng.normal(0, 0.02, (F, 3)); bg = rng.normal(0, 0.005, (F, 3))
    ref = synth.preintegrate_batch(0.01, acc, gyr, ba, bg)
    first = (np.arange(F + 1) * S).astype(np.int32)
    samples = np.concatenate([np.full((F, S, 1), 0.01), acc[:, 1:], gyr[:, 1:]], axis=2).reshape(F * S, 7)
    out = backend.preintegrate(orc_ctx, first, samples, acc[:, 0], gyr[:, 0], ba, bg, np.array(synth.IMU_NOISE))
    assert out.shape == ref.shape
    assert np.max(np.abs(out[:, :17] - ref[:, :17])) < 1e-6
    assert np.max(np.abs(out[:, 17:242] - ref[:, 17:242])) < 1e-6
    assert np.max(np.abs(out[:, 242:467] - ref[:, 242:467])) < 1e-6 * np.abs(ref[:, 242:467]).max()
