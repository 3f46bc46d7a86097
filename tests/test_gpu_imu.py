"""GPU parity of the batched IMU preintegration (lvb_imu_preintegrate) against the CPU oracle's restatement of
Preintegration::Propagate / Repropagate (preintegration.cpp:30-142).  Tolerance: 1e-12 relative to the largest entry of
each record section (FP64, same summation order; the device contracts multiply-adds into FMAs, the oracle does not)."""
import numpy as np
import pytest

from lvio_fusion_b200 import backend, synth

pytestmark = pytest.mark.gpu


def _ragged_samples(n, seed, lo=0, hi=25):
    rng = np.random.default_rng(seed)
    counts = rng.integers(lo, hi, size=n)
    first = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    m = int(first[-1])
    samples = np.empty((m, 7))
    samples[:, 0] = rng.uniform(0.004, 0.012, m)
    samples[:, 1:4] = rng.normal(0, 1.5, (m, 3)) + [0, 0, 9.81]
    samples[:, 4:7] = rng.normal(0, 0.3, (m, 3))
    acc0 = rng.normal(0, 1.5, (n, 3)) + [0, 0, 9.81]
    gyr0 = rng.normal(0, 0.3, (n, 3))
    ba = rng.normal(0, 0.05, (n, 3)); bg = rng.normal(0, 0.01, (n, 3))
    return first, samples, acc0, gyr0, ba, bg


def _close(a, b, tol):
    for lo, hi in ((0, 3), (3, 7), (7, 10), (10, 17), (17, 242), (242, 467), (467, 469)):
        sa, sb = a[:, lo:hi], b[:, lo:hi]
        assert np.max(np.abs(sa - sb)) <= tol * max(1e-300, np.max(np.abs(sb))), (lo, hi)


def test_preintegrate_matches_oracle(lvb_ctx, orc_ctx):
    args = _ragged_samples(300, 5)          # includes empty intervals (identity Jacobian, zero covariance)
    noise = np.array(synth.IMU_NOISE, dtype=np.float64)
    g = backend.preintegrate(lvb_ctx, *args, noise)
    o = backend.preintegrate(orc_ctx, *args, noise)
    assert g.shape == o.shape == (300, 469)
    _close(g, o, 1e-12)
    empty = np.flatnonzero(np.diff(args[0]) == 0)
    assert len(empty) > 0
    assert np.array_equal(g[empty, 17:242].reshape(-1, 15, 15), np.tile(np.eye(15), (len(empty), 1, 1)))


def test_repropagate_and_factor_use(lvb_ctx, orc_ctx):
    """Repropagate = the same call with other biases (preintegration.cpp:129-142); the records feed LVB_IMU factors."""
    first, samples, acc0, gyr0, ba, bg = _ragged_samples(40, 6, lo=5, hi=15)
    noise = np.array(synth.IMU_NOISE, dtype=np.float64)
    g0 = backend.preintegrate(lvb_ctx, first, samples, acc0, gyr0, ba, bg, noise)
    g1 = backend.preintegrate(lvb_ctx, first, samples, acc0, gyr0, ba + 0.01, bg - 0.002, noise)
    o1 = backend.preintegrate(orc_ctx, first, samples, acc0, gyr0, ba + 0.01, bg - 0.002, noise)
    _close(g1, o1, 1e-12)
    assert np.max(np.abs(g1[:, :3] - g0[:, :3])) > 1e-6
    # first-order bias correction of the factor (preintegration.cpp:151-157) predicts the repropagated delta_p
    J = g0[:, 17:242].reshape(-1, 15, 15)
    pred = g0[:, 0:3] + np.einsum("fij,j->fi", J[:, 0:3, 9:12], np.full(3, 0.01)) + np.einsum("fij,j->fi", J[:, 0:3, 12:15], np.full(3, -0.002))
    assert np.max(np.abs(pred - g1[:, 0:3])) < 1e-5
    # the device records drive a BA problem exactly like host-produced ones
    d = synth.make_ba_problem(6, 300, with_imu=True, seed=8)
    c, ix = d["factors"][backend.IMU]
    d["factors"][backend.IMU] = (np.ascontiguousarray(g0[:len(c)]), ix)
    pg, po = backend.Problem.from_dict(lvb_ctx, d), backend.Problem.from_dict(orc_ctx, d)
    rg, Jg = pg.evaluate(backend.IMU)
    ro, Jo = po.evaluate(backend.IMU)
    assert np.max(np.abs(rg - ro)) < 1e-8 * max(1.0, np.abs(ro).max())
    assert np.max(np.abs(Jg - Jo)) < 1e-8 * max(1.0, np.abs(Jo).max())
