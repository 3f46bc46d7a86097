"""GPU parity tests of the scan-to-map path: voxel-hash kNN bit-exact against the oracle (indices and
float32 squared distances), association gate, residual/Jacobian within 1e-9, post-solve rpyxyz within
1e-7."""
import numpy as np
import pytest

from lvio_fusion_b200 import backend, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["surf", "ground"])
def scene(request):
    return synth.make_icp_problem(3000, 40000, seed=21, kind=request.param)


def _pair(lvb_ctx, orc_ctx, scene, brute=False):
    fg, fo = backend.FeatureAssociation(lvb_ctx), backend.FeatureAssociation(orc_ctx)
    if brute:
        orc_ctx.api.icp_set_brute(fo.h, 1)
    fg.set_map(scene["map"], scene["cell_size"]); fo.set_map(scene["map"], scene["cell_size"])
    return fg, fo


@pytest.mark.parametrize("stride_floats", [4, 8])
def test_knn3_bit_exact(lvb_ctx, orc_ctx, stride_floats):
    sc = synth.make_icp_problem(1500, 20000, seed=5, kind="surf", stride_floats=stride_floats)
    fg, fo = _pair(lvb_ctx, orc_ctx, sc, brute=True)
    m2 = sc["cell_size"] ** 2
    ig, dg = fg.knn3(sc["scan"], sc["frame_pose"], m2)
    io, do = fo.knn3(sc["scan"], sc["frame_pose"], m2)
    assert (ig >= 0).mean() > 0.5
    assert np.array_equal(ig, io)
    assert np.array_equal(dg.view(np.uint32), do.view(np.uint32))


def test_knn3_larger_against_kdtree(lvb_ctx, orc_ctx):
    sc = synth.make_icp_problem(20000, 300000, seed=6, kind="ground")
    fg, fo = _pair(lvb_ctx, orc_ctx, sc)
    orc_ctx.api.icp_set_threads(fo.h, 4)
    m2 = sc["cell_size"] ** 2
    ig, dg = fg.knn3(sc["scan"], sc["frame_pose"], m2)
    io, do = fo.knn3(sc["scan"], sc["frame_pose"], m2)
    assert np.array_equal(ig, io) and np.array_equal(dg.view(np.uint32), do.view(np.uint32))


def test_association_and_eval(lvb_ctx, orc_ctx, scene):
    fg, fo = _pair(lvb_ctx, orc_ctx, scene)
    e0 = synth.relative_rpyxyz(scene["map_pose"], scene["frame_pose"])
    ag, rg, Jg = fg.evaluate(scene["mode"], scene["scan"], scene["frame_pose"], scene["map_pose"], e0, scene["weight"], scene["thr"])
    ao, ro, Jo = fo.evaluate(scene["mode"], scene["scan"], scene["frame_pose"], scene["map_pose"], e0, scene["weight"], scene["thr"])
    assert np.array_equal(ag, ao) and ag.sum() > 100
    ok = np.isfinite(ro)
    assert np.array_equal(np.isfinite(rg), ok)
    assert np.max(np.abs(rg[ok] - ro[ok])) < 1e-9 * max(1.0, np.abs(ro[ok]).max())
    okJ = np.isfinite(Jo).all(axis=1)
    assert np.max(np.abs(Jg[okJ] - Jo[okJ])) < 1e-9 * max(1.0, np.abs(Jo[okJ]).max())


@pytest.mark.parametrize("prior", [True, False])
def test_scan_to_map_matches_oracle(lvb_ctx, orc_ctx, scene, prior):
    fg, fo = _pair(lvb_ctx, orc_ctx, scene)
    e0 = synth.relative_rpyxyz(scene["map_pose"], scene["frame_pose"])
    pw = scene["n_features_left"] * synth.W_VISUAL if prior else -1.0
    args = (scene["mode"], scene["scan"], scene["frame_pose"], scene["map_pose"], e0, scene["weight"], pw, scene["huber_a"], scene["thr"])
    eg, sg = fg.scan_to_map(*args)
    eo, so = fo.scan_to_map(*args)
    assert sg.num_residual_blocks == so.num_residual_blocks
    assert sg.num_iterations == so.num_iterations
    assert abs(sg.initial_cost - so.initial_cost) < 1e-9 * max(1e-30, so.initial_cost)
    assert abs(sg.final_cost - so.final_cost) < 1e-7 * max(1e-30, so.final_cost)
    assert np.max(np.abs(eg - eo)) < 1e-7
    if not prior:
        assert not np.array_equal(eg, e0)    # with the reference's prior weight (n_features * fx/10) the prior pins the pose


def test_empty_and_unmatched_scans(lvb_ctx, scene):
    fg = backend.FeatureAssociation(lvb_ctx)
    fg.set_map(scene["map"], scene["cell_size"])
    e0 = synth.relative_rpyxyz(scene["map_pose"], scene["frame_pose"])
    far = scene["scan"].copy(); far[:, :3] += 1000.0
    e, s = fg.scan_to_map(scene["mode"], far, scene["frame_pose"], scene["map_pose"], e0, scene["weight"], 100.0, scene["huber_a"], scene["thr"])
    assert s.num_residual_blocks == 1 and np.allclose(e, e0)     # only the prior: stays put
    idx, d2 = fg.knn3(far, scene["frame_pose"], scene["cell_size"] ** 2)
    assert (idx == -1).all() and np.isinf(d2).all()
    idx, d2 = fg.knn3(np.zeros((0, 4), dtype=np.float32), scene["frame_pose"], 1.0)
    assert idx.shape == (0, 3)


def test_merge_scan_transform_bit_exact(lvb_ctx, orc_ctx):
    """Mapping::MergeScan / ToWorld (mapping.cpp:193-220): float32 transform of 32-byte PointXYZI records."""
    sc = synth.make_icp_problem(5000, 2000, seed=9, kind="surf", stride_floats=8)
    sc["scan"][:, 4] = np.arange(len(sc["scan"]), dtype=np.float32)          # intensity must survive
    fg, fo = backend.FeatureAssociation(lvb_ctx), backend.FeatureAssociation(orc_ctx)
    wg, wo = fg.transform_cloud(sc["scan"], sc["true_pose"]), fo.transform_cloud(sc["scan"], sc["true_pose"])
    assert np.array_equal(wg.view(np.uint32), wo.view(np.uint32))
    assert np.array_equal(wg[:, 4], sc["scan"][:, 4])
    ref = synth.se3_apply(np.broadcast_to(sc["true_pose"], (len(wg), 7)), sc["scan"][:, :3].astype(np.float64))
    assert np.max(np.abs(wg[:, :3] - ref)) < 1e-4
