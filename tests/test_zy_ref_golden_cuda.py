"""The CUDA half of tests/test_ref_golden.py: the same fixtures (outputs of the reference's own sources compiled in place),
the same checks, on liblvio_b200.so.  Kept in a file that sorts after every test_gpu_*.py."""
import numpy as np
import pytest

import test_ref_golden as R
from test_ref_golden import ref      # noqa: F401  (fixture)
from lvio_fusion_b200 import backend


@pytest.mark.gpu
def test_cuda_reproduces_the_reference_functors(lvb_ctx, orc, ref):
    R._check(lvb_ctx, ref, orc, 1e-9)


@pytest.mark.gpu
def test_cuda_imu_path_reproduces_the_reference(lvb_ctx, ref):
    out = backend.preintegrate(lvb_ctx, ref["imu_first"], ref["imu_samples"], ref["imu_acc0"], ref["imu_gyr0"], ref["imu_ba"], ref["imu_bg"], ref["imu_noise"])
    g = ref["imu_record"]
    for lo, hi in ((0, 17), (17, 242), (242, 467)):
        assert np.max(np.abs(out[:, lo:hi] - g[:, lo:hi])) <= 1e-11 * np.abs(g[:, lo:hi]).max(), (lo, hi)
    p = R._imu_problem(lvb_ctx, ref, R._records_from_reference(ref))
    r, J = p.evaluate(backend.IMU)
    assert R._rel(r, ref["imu_r"]) < 1e-7
    assert R._rel(J.reshape(ref["imu_J"].shape), ref["imu_J"]) < 1e-7
    rec = R._records_from_reference(ref); rec[:, 467:] = 1e8
    ri, Ji = R._imu_problem(lvb_ctx, ref, rec, init=True).evaluate(backend.IMU)
    assert R._rel(ri, ref["imuinit_r"]) < 1e-7
    assert R._rel(Ji.reshape(len(ri), 15, 32)[:, :, :26], ref["imuinit_J"]) < 1e-7
    R._check_real_priors(lvb_ctx, ref, 1e-6)


@pytest.mark.gpu
def test_cuda_reproduces_the_reference_range_image_code(lvb_ctx):
    R._check_lidar_projection(lvb_ctx)


@pytest.mark.gpu
def test_cuda_reproduces_the_reference_feature_extraction_and_association(lvb_ctx):
    R._check_extract(lvb_ctx)
    R._check_scan2map(lvb_ctx, brute=False)


@pytest.mark.gpu
def test_cuda_knn_matches_flann(lvb_ctx):
    """tests/test_flann_pin.py on the device: the voxel-hash 3-NN against FLANN's KDTreeSingleIndex (OpenCV's bundled copy)."""
    F = pytest.importorskip("test_flann_pin")
    F._check(lvb_ctx, None, 20000, 300000, "surf", 23, F.FRAC)
    F._check(lvb_ctx, None, 20000, 300000, "ground", 24, F.FRAC)
