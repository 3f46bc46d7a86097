"""Device-resident map (Mapping::ToWorld / BuildMapFrame, mapping.cpp:114-137,193-220): the per-keyframe world clouds
stay on the device and the map frame is merged, ground-filtered and hashed there.  It must give exactly what the per-call path
gives -- MergeScan on the host side of the API, concatenation, SegmentGround, set_map: same merged cloud (bit for bit), same 3-NN
indices and distances, same registration."""
import numpy as np
import pytest

from lvio_fusion_b200 import backend, synth

pytestmark = pytest.mark.gpu


def _keyframe_clouds(sc, n_kf, rng):
    """Split the synthetic world map into n_kf robot-frame clouds with their keyframe poses."""
    mp = sc["map"]
    parts = np.array_split(np.arange(len(mp)), n_kf)
    out = []
    for k, idx in enumerate(parts):
        pose = sc["map_pose"].copy()
        pose[4:] += rng.normal(0, 0.5, 3)
        q = rng.normal(0, 0.02, 3)
        pose[:4] = synth.quat_mul(synth.quat_from_rotvec(q[None])[0][None], pose[:4][None])[0]
        w = mp[idx].astype(np.float64)
        robot = mp[idx].copy()
        robot[:, :3] = synth.se3_inv_apply(np.broadcast_to(pose, (len(idx), 7)), w[:, :3]).astype(np.float32)
        out.append((pose, np.ascontiguousarray(robot)))
    return out


@pytest.mark.parametrize("kind,ground", [("surf", False), ("ground", True)])
def test_resident_map_equals_per_call_path(lvb_ctx, kind, ground):
    rng = np.random.default_rng(8)
    sc = synth.make_icp_problem(6000, 90000, seed=12, kind=kind)
    kfs = _keyframe_clouds(sc, 3, rng)
    fa, fb = backend.FeatureAssociation(lvb_ctx), backend.FeatureAssociation(lvb_ctx)
    # per-call path: MergeScan per keyframe (device transform, host concatenation), optional SegmentGround, set_map
    world = [fa.transform_cloud(c, p) for p, c in kfs]
    merged = np.ascontiguousarray(np.concatenate(world))
    thr = 0.02
    if ground:
        merged = np.ascontiguousarray(backend.LidarFeatures(lvb_ctx).segment_ground(merged[:, :4], thr))      # SegmentGround(points_ground_merged), mapping.cpp:126
        assert 1000 < len(merged) < sum(len(w) for w in world)
    fa.set_map(merged, sc["cell_size"])
    # resident path
    for k, (p, c) in enumerate(kfs):
        fb.map_append(100 + k, c, p)
    n = fb.map_build([100, 101, 102], sc["cell_size"], thr if ground else -1.0)
    assert n == len(merged)
    got = fb.map_download()
    assert np.array_equal(got[:, :3].view(np.uint32), np.ascontiguousarray(merged[:, :3]).view(np.uint32))
    m2 = sc["cell_size"] ** 2
    ia, da = fa.knn3(sc["scan"], sc["frame_pose"], m2)
    ib, db = fb.knn3(sc["scan"], sc["frame_pose"], m2)
    assert (ia >= 0).mean() > 0.3
    assert np.array_equal(ia, ib) and np.array_equal(da.view(np.uint32), db.view(np.uint32))
    e0 = synth.relative_rpyxyz(sc["map_pose"], sc["frame_pose"])
    args = (sc["mode"], sc["scan"], sc["frame_pose"], sc["map_pose"], e0, sc["weight"], -1.0, sc["huber_a"], sc["thr"])
    ea, sa = fa.scan_to_map(*args)
    eb, sb = fb.scan_to_map(*args)
    assert sa.num_residual_blocks == sb.num_residual_blocks and np.max(np.abs(ea - eb)) < 1e-11        # (the 3 x 3 sums are reduced with atomics)
    # sliding the window: the scan just registered becomes a keyframe cloud without a second upload; the oldest one leaves
    fb.map_append(103, None, sc["frame_pose"])
    fb.map_evict(100)
    n2 = fb.map_build([101, 102, 103], sc["cell_size"])
    assert n2 == len(world[1]) + len(world[2]) + len(sc["scan"])
    with pytest.raises(RuntimeError):
        fb.map_build([100, 101], sc["cell_size"])          # evicted
