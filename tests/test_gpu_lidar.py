"""GPU parity of the lidar feature pipeline (lvb_lidar_*) against the CPU oracle (oracle/lidar.h) through the C ABI.
Everything on this path is float32 / integer work: the bar is **bit-exact** (array_equal), including the order of the
output clouds."""
import numpy as np
import pytest

from lvio_fusion_b200 import backend, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pipes(lvb_ctx, orc_ctx):
    return backend.LidarFeatures(lvb_ctx), backend.LidarFeatures(orc_ctx)


@pytest.mark.parametrize("seed", [11, 12])
def test_segment_bit_exact(pipes, seed):
    g, o = pipes
    scan = synth.make_lidar_scan(seed=seed)
    sg, so = g.segment(scan), o.segment(scan)
    assert len(so["points"]) > 40000
    for k in ("points", "range", "ground", "col", "curvature", "start_ring", "end_ring", "orientation"):
        assert np.array_equal(sg[k], so[k]), k


def test_segment_edge_cases(pipes):
    g, o = pipes
    empty = np.zeros((0, 4), np.float32)
    nan = np.full((100, 4), np.nan, np.float32)
    near = np.tile(np.array([[1.0, 0.5, 0.1, 0.0]], np.float32), (50, 1))          # inside min_range: all filtered
    rng = np.random.default_rng(0)
    sparse = (rng.normal(0, 8, (300, 4))).astype(np.float32)                       # few scattered returns: tiny segments
    for cloud in (empty, nan, near, sparse):
        sg, so = g.segment(cloud), o.segment(cloud)
        for k in ("points", "range", "ground", "col", "curvature", "start_ring", "end_ring", "orientation"):
            assert np.array_equal(sg[k], so[k]), k
    # a 32-byte record stride (pcl::PointXYZI with padding) gives the same result as the packed one
    scan = synth.make_lidar_scan(seed=13, horizon_scan=900)
    wide = np.zeros((len(scan), 8), np.float32); wide[:, :3] = scan[:, :3]; wide[:, 4] = 7.0
    g9 = backend.LidarFeatures(g.ctx, horizon_scan=900); o9 = backend.LidarFeatures(o.ctx, horizon_scan=900)
    a, b, c = g9.segment(scan), g9.segment(wide), o9.segment(scan)
    assert np.array_equal(a["points"], b["points"]) and np.array_equal(a["points"], c["points"]) and np.array_equal(a["curvature"], c["curvature"])


def test_filters_bit_exact(pipes):
    g, o = pipes
    rng = np.random.default_rng(3)
    cloud = np.concatenate([rng.uniform(-25, 25, (20000, 2)), rng.uniform(-2, 3, (20000, 1)), rng.uniform(0, 64, (20000, 1))], axis=1).astype(np.float32)
    vg, vo = g.voxel_grid(cloud, 0.4), o.voxel_grid(cloud, 0.4)
    assert len(vo) < len(cloud) and np.array_equal(vg, vo)
    rg, ro = g.radius_outlier_removal(vo, 0.8, 4), o.radius_outlier_removal(vo, 0.8, 4)
    assert 0 < len(ro) < len(vo) and np.array_equal(rg, ro)
    dense = cloud.copy(); dense[:, :3] *= np.float32(0.05)                          # many points per voxel: in-voxel order matters
    assert np.array_equal(g.voxel_grid(dense, 0.4), o.voxel_grid(dense, 0.4))
    for tiny in (np.zeros((0, 4), np.float32), cloud[:1], cloud[:2]):
        assert np.array_equal(g.voxel_grid(tiny, 0.4), o.voxel_grid(tiny, 0.4))
        assert np.array_equal(g.radius_outlier_removal(tiny, 0.8, 1), o.radius_outlier_removal(tiny, 0.8, 1))
        assert np.array_equal(g.segment_ground(tiny, 0.02), o.segment_ground(tiny, 0.02))


def test_segment_ground_bit_exact(pipes):
    g, o = pipes
    rng = np.random.default_rng(4)
    plane = np.stack([rng.uniform(-20, 20, 3000), rng.uniform(-20, 20, 3000), -1.73 + rng.normal(0, 0.006, 3000)], 1)
    junk = np.stack([rng.uniform(-20, 20, 1500), rng.uniform(-20, 20, 1500), rng.uniform(-1.6, 1.0, 1500)], 1)
    cloud = np.concatenate([plane, junk])[rng.permutation(4500)]
    cloud = np.concatenate([cloud, rng.uniform(0, 64, (4500, 1))], 1).astype(np.float32)
    a, b = g.segment_ground(cloud, 0.02), o.segment_ground(cloud, 0.02)
    assert len(b) > 2000 and np.array_equal(a, b)
    # degenerate: all points collinear -> no sample is ever "good" -> empty
    line = np.zeros((50, 4), np.float32); line[:, 0] = np.arange(50)
    assert len(o.segment_ground(line, 0.02)) == 0 and len(g.segment_ground(line, 0.02)) == 0


@pytest.mark.parametrize("seed", [11, 14])
def test_extract_features_bit_exact(lvb_ctx, orc_ctx, seed):
    ext = [0.01, -0.02, 0.7, 0.7, 0.8, -0.3, 0.9]                                   # un-normalised on purpose: ceres normalises
    g, o = backend.LidarFeatures(lvb_ctx, extrinsic=ext), backend.LidarFeatures(orc_ctx, extrinsic=ext)
    scan = synth.make_lidar_scan(seed=seed)
    gg, gs = g.extract(scan)
    og, os_ = o.extract(scan)
    assert len(og) > 500 and len(os_) > 500
    assert np.array_equal(gg, og)
    assert np.array_equal(gs, os_)
    # the features feed the scan-to-map entry point unchanged (x y z intensity, 16-byte records)
    fa = backend.FeatureAssociation(lvb_ctx)
    fa.set_map(gs, 2.0)
    idx, d2 = fa.knn3(gs[:100], np.array([0, 0, 0, 1, 0, 0, 0.0]), 4.0)
    assert np.array_equal(idx[:, 0], np.arange(100)) and np.all(d2[:, 0] == 0)
