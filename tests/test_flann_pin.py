"""Row a7's nearest-neighbour search pinned against FLANN itself.

pcl::KdTreeFLANN (association.cpp:278-300) is a thin wrapper over flann::Index<L2_Simple<float>> with KDTreeSingleIndexParams(15) and
SearchParams(checks = -1, eps = 0, sorted) (pcl/kdtree/impl/kdtree_flann.hpp) -- an exact search.  PCL and its FLANN are not in this
image, but OpenCV's Python module is, and it bundles the same library (cv2.flann_Index, algorithm 4 = FLANN_INDEX_KDTREE_SINGLE): the
single kd-tree, exact, float32 squared distances accumulated x, y, z.  The oracle's two searches (brute force, its own kd-tree)
and the CUDA voxel-hash search must return the same three neighbour indices and bit-identical squared distances as FLANN for every
scan point whose neighbours lie inside the gate; outside the gate they report -1 (the reference drops those points, :297-300)."""
import numpy as np
import pytest

from lvio_fusion_b200 import backend, synth

cv2 = pytest.importorskip("cv2")
FRAC = 0.8


def _flann(map_xyz, query_xyz):
    index = cv2.flann_Index(np.ascontiguousarray(map_xyz, dtype=np.float32), dict(algorithm=4, leaf_max_size=15, reorder=True, dim=3))
    idx, d2 = index.knnSearch(np.ascontiguousarray(query_xyz, dtype=np.float32), 3, params=dict(checks=-1, eps=0.0, sorted=True))
    return idx.astype(np.int32), d2.astype(np.float32)


def _check(ctx, brute, n_scan, n_map, kind, seed, frac):
    sc = synth.make_icp_problem(n_scan, n_map, seed=seed, kind=kind)
    fa = backend.FeatureAssociation(ctx)
    if brute is not None:
        ctx.api.icp_set_brute(fa.h, 1 if brute else 0)
    fa.set_map(sc["map"], sc["cell_size"])
    world = fa.transform_cloud(sc["scan"], sc["frame_pose"])                       # the float32 transform of association.cpp:289-291, bit-exact elsewhere
    fidx, fd2 = _flann(np.asarray(sc["map"])[:, :3], np.asarray(world)[:, :3])
    gate = np.float32(np.percentile(fd2[:, 2], 100 * frac))                         # a gate that cuts off the farthest third neighbours (< cell^2)
    assert gate < np.float32(sc["cell_size"]) ** 2
    idx, d2 = fa.knn3(sc["scan"], sc["frame_pose"], float(gate))
    inside = fd2 < gate
    assert 0.5 < inside[:, 2].mean() < 1.0                                        # most points have all three neighbours inside the gate, some do not
    assert np.array_equal(idx[inside], fidx[inside])
    assert np.array_equal(d2[inside].view(np.uint32), fd2[inside].view(np.uint32))
    assert np.all(idx[~inside] == -1)
    # no exact distance ties among the compared neighbours (their order would be FLANN's traversal order: not pinned)
    assert np.all(np.diff(fd2[inside[:, 2]], axis=1) > 0)


@pytest.mark.parametrize("brute", [True, False])
def test_oracle_knn_matches_flann(orc_ctx, brute):
    _check(orc_ctx, brute, 3000, 60000, "surf", 21, FRAC)
    _check(orc_ctx, brute, 3000, 60000, "ground", 22, FRAC)

