"""The reduced-system solver on its own (SuiteSparse's role behind SPARSE_SCHUR, backend.cpp:207): random SPD band systems through
the multifrontal separator tree (ba_tree.cuh) and through the single-CTA envelope kernel, against numpy's dense solve."""
import ctypes as C

import numpy as np
import pytest

from lvio_fusion_b200 import _capi

pytestmark = pytest.mark.gpu


def _band_spd(n, true_band, rng):
    a = np.zeros((n, n))
    for i in range(n):
        lo = max(0, i - true_band)
        a[i, lo:i + 1] = rng.normal(size=i + 1 - lo)
    s = a @ a.T
    s[np.abs(np.subtract.outer(np.arange(n), np.arange(n))) > true_band] = 0.0
    return s + 1e-3 * np.diag(np.diag(s)) + 1e-6 * np.eye(n)


def _pack(s, band):
    n = len(s)
    out = np.zeros((n, band + 1))
    for i in range(n):
        lo = max(0, i - band)
        out[i, band - (i - lo):] = s[i, lo:i + 1]
    return out


@pytest.mark.parametrize("n,true_band,expect_tree", [(150, 40, False), (700, 45, True), (1200, 59, True), (2500, 33, True), (4000, 75, True), (9000, 50, True),
                                                     (900, 23, True), (1500, 10, True), (600, 31, True), (2000, 5, True)])
def test_tree_and_single_cta_solve_match_numpy(lvb_ctx, n, true_band, expect_tree):
    rng = np.random.default_rng(n)
    s = _band_spd(n, true_band, rng)
    b = rng.normal(size=n)
    want = np.linalg.solve(s, b)
    band = true_band + 31
    packed = np.ascontiguousarray(_pack(s, band))
    api = lvb_ctx.api
    for use_tree in (1, 0):
        x = np.zeros(n)
        lv = C.c_int(-1)
        api.check(api.debug_band_solve(lvb_ctx.h, n, band, packed.ctypes.data_as(_capi.c_double_p), b.ctypes.data_as(_capi.c_double_p),
                                       x.ctypes.data_as(_capi.c_double_p), use_tree, C.byref(lv)), "debug_band_solve")
        if use_tree:
            assert (lv.value > 0) == expect_tree, lv.value
        else:
            assert lv.value == 0
        err = np.max(np.abs(x - want)) / np.max(np.abs(want))
        resid = np.max(np.abs(s @ x - b)) / np.max(np.abs(b))
        assert err < 1e-8 and resid < 1e-9, (use_tree, lv.value, err, resid)


def test_indefinite_system_is_reported(lvb_ctx):
    n, true_band = 900, 40
    rng = np.random.default_rng(5)
    s = _band_spd(n, true_band, rng)
    s[600, 600] = -1.0
    band = true_band + 31
    packed = np.ascontiguousarray(_pack(s, band)); b = np.ones(n); x = np.zeros(n)
    api = lvb_ctx.api
    rc = api.debug_band_solve(lvb_ctx.h, n, band, packed.ctypes.data_as(_capi.c_double_p), b.ctypes.data_as(_capi.c_double_p),
                              x.ctypes.data_as(_capi.c_double_p), 1, None)
    assert rc != 0 and b"pivot" in api.last_error()


@pytest.mark.parametrize("n,true_band", [(1, 0), (20, 19), (32, 31), (33, 32), (47, 46), (64, 5), (95, 20), (129, 128), (150, 149), (160, 64), (300, 299), (300, 91), (736, 100)])
def test_single_cta_kernel_sizes(lvb_ctx, n, true_band):
    """The one-CTA kernel alone over the shapes its schedule distinguishes: fewer rows than a block, a partial last block, a full
    window (150 / 300 unknowns, dense), head rows cut by the envelope (band narrower than a block), the widest dense layout."""
    rng = np.random.default_rng(n * 7 + true_band)
    s = _band_spd(n, true_band, rng)
    b = rng.normal(size=n)
    want = np.linalg.solve(s, b)
    band = max(31, true_band + 31)
    packed = np.ascontiguousarray(_pack(s, band)); x = np.zeros(n); lv = C.c_int(-1)
    api = lvb_ctx.api
    api.check(api.debug_band_solve(lvb_ctx.h, n, band, packed.ctypes.data_as(_capi.c_double_p), b.ctypes.data_as(_capi.c_double_p),
                                   x.ctypes.data_as(_capi.c_double_p), 0, C.byref(lv)), "debug_band_solve")
    assert lv.value == 0
    err = np.max(np.abs(x - want)) / np.max(np.abs(want))
    resid = np.max(np.abs(s @ x - b)) / np.max(np.abs(b))
    assert err < 1e-9 and resid < 1e-10, (err, resid)


def test_single_cta_kernel_reports_a_non_positive_pivot(lvb_ctx):
    n = 150
    rng = np.random.default_rng(3)
    s = _band_spd(n, n - 1, rng)
    s[101, 101] = -5.0          # inside the fourth diagonal block: found by the two-column pivot step
    band = n - 1 + 31
    packed = np.ascontiguousarray(_pack(s, band)); b = np.ones(n); x = np.zeros(n)
    api = lvb_ctx.api
    rc = api.debug_band_solve(lvb_ctx.h, n, band, packed.ctypes.data_as(_capi.c_double_p), b.ctypes.data_as(_capi.c_double_p),
                              x.ctypes.data_as(_capi.c_double_p), 0, None)
    assert rc != 0 and b"pivot" in api.last_error()


@pytest.mark.parametrize("n,true_band", [(1500, 1200), (2100, 2099), (1000, 930)])
def test_wide_envelope_uses_the_multi_grid_solver(lvb_ctx, n, true_band):
    """Envelopes too wide for the one-CTA panel and with no room for a separator tree (loop closures) go through the per-phase grids
    of ba_wide.cuh instead of LVB_ERR_UNSUPPORTED."""
    rng = np.random.default_rng(n + true_band)
    s = _band_spd(n, true_band, rng)
    b = rng.normal(size=n)
    want = np.linalg.solve(s, b)
    band = true_band + 31
    packed = np.ascontiguousarray(_pack(s, band)); api = lvb_ctx.api
    for use_tree in (1, 0):
        x = np.zeros(n); lv = C.c_int(-1)
        api.check(api.debug_band_solve(lvb_ctx.h, n, band, packed.ctypes.data_as(_capi.c_double_p), b.ctypes.data_as(_capi.c_double_p),
                                       x.ctypes.data_as(_capi.c_double_p), use_tree, C.byref(lv)), "debug_band_solve")
        assert lv.value == 0
        err = np.max(np.abs(x - want)) / np.max(np.abs(want))
        resid = np.max(np.abs(s @ x - b)) / np.max(np.abs(b))
        assert err < 1e-8 and resid < 1e-9, (use_tree, err, resid)
