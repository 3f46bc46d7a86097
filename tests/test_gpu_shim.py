"""The C++ host side (include/lvio_b200/*.h) driven like Backend::BuildProblem / Mapping::Optimize drive Ceres and
PCL, compared with the CPU oracle on the same problem."""
import subprocess

import numpy as np
import pytest

from lvio_fusion_b200 import backend, synth
from lvio_fusion_b200.backend import POSE_PRIOR

import shim_util

pytestmark = pytest.mark.gpu


def test_backend_shaped_solve_through_the_shim(tmp_path, orc_ctx):
    d = synth.make_ba_problem(6, 800, with_imu=True, seed=13)
    d["factors"] = dict(d["factors"])
    d["factors"][POSE_PRIOR] = (np.concatenate([d["poses"][0], [100.0, 0.0]])[None], np.zeros((1, 1), dtype=np.int32))
    exe = shim_util.build_shim_binary()
    shim_util.dump_ba(tmp_path / "in.bin", d, 20)
    subprocess.check_call([exe, "ba", str(tmp_path / "in.bin"), str(tmp_path / "out.bin")])
    P, V, R, s = shim_util.load_ba_result(tmp_path / "out.bin", d)
    po = backend.Problem.from_dict(orc_ctx, d)
    so = po.solve(max_num_iterations=20, num_threads=2)
    assert s["num_residual_blocks"] == so.num_residual_blocks
    assert abs(s["initial_cost"] - so.initial_cost) < 1e-9 * so.initial_cost
    assert abs(s["final_cost"] - so.final_cost) < 1e-6 * so.final_cost
    assert np.max(np.abs(P - po.poses())) < 1e-6
    assert np.max(np.abs(R - po.inv_depths())) < 1e-6
    assert np.max(np.abs(V - po.vec3())) < 1e-5
    assert not np.allclose(P, d["poses"])          # the caller's memory was updated in place


@pytest.mark.parametrize("kind", ["ground", "surf"])
def test_mapping_shaped_scan_to_map_through_the_shim(tmp_path, orc_ctx, kind):
    sc = synth.make_icp_problem(2500, 30000, seed=17, kind=kind)
    sc["n_features_left"] = 0 if kind == "surf" else 2      # weak prior so that the pose actually moves
    e0 = synth.relative_rpyxyz(sc["map_pose"], sc["frame_pose"])
    exe = shim_util.build_shim_binary()
    shim_util.dump_icp(tmp_path / "in.bin", sc, e0)
    subprocess.check_call([exe, "icp", str(tmp_path / "in.bin"), str(tmp_path / "out.bin")])
    out = np.fromfile(tmp_path / "out.bin", dtype=np.float64)
    fo = backend.FeatureAssociation(orc_ctx)
    cell = float(np.nextafter(np.float32(np.sqrt(sc["thr"])), np.float32(1e30))) * 1.0001
    fo.set_map(sc["map"], np.float32(cell))
    eo, so = fo.scan_to_map(sc["mode"], sc["scan"], sc["frame_pose"], sc["map_pose"], e0, sc["weight"],
                            sc["n_features_left"] * synth.W_VISUAL, sc["huber_a"], sc["thr"])
    assert int(out[8]) == so.num_residual_blocks
    assert abs(out[7] - so.final_cost) < 1e-7 * max(so.final_cost, 1e-30)
    assert np.max(np.abs(out[:6] - eo)) < 1e-7
    # the same registration against a device-resident map frame (AppendKeyframe x3 + BuildMapFrame)
    subprocess.check_call([exe, "icp_resident", str(tmp_path / "in.bin"), str(tmp_path / "out_r.bin")])
    out_r = np.fromfile(tmp_path / "out_r.bin", dtype=np.float64)
    assert int(out_r[8]) == int(out[8]) and np.max(np.abs(out_r[:6] - out[:6])) < 1e-10
