"""Wire formats (result.csv, TUM ground truth) and the ATE harness (lvio_fusion_node.cpp:295-350)."""
import numpy as np

from lvio_fusion_b200 import evaluation as ev
from lvio_fusion_b200 import synth


def test_result_csv_format_and_round_trip(tmp_path):
    P = synth.Trajectory(12).poses()
    P[:, :4] *= 1.7                                   # the writer emits the unit quaternion
    t = 1317617735.5 + 0.1 * np.arange(12)
    p = tmp_path / "result.csv"
    ev.write_result(str(p), t, P, init_time=t[0])
    lines = p.read_text().splitlines()
    assert len(lines) == 12
    first = lines[0].split(",")
    assert len(first) == 8 and first[0] == "0.00000"
    assert all(len(v.split(".")[1]) == 5 for v in first)          # ios::fixed, precision 5
    tt, PP = ev.read_result(str(p))
    assert np.allclose(tt, t - t[0], atol=5e-6)
    assert np.allclose(PP[:, 4:], P[:, 4:], atol=5e-6)
    assert np.allclose(PP[:, :4], P[:, :4] / 1.7, atol=5e-6)
    # negative zero keeps its sign like the C++ stream does
    ev.write_result(str(p), [0.0], [[0, 0, 0, 1, -1e-7, 0, 0]])
    assert p.read_text().startswith("0.00000,-0.00000,")


def test_ground_truth_frame_change(tmp_path):
    # a KITTI-camera pose (z forward, x right, y down) maps to the body convention (x forward, y left, z up)
    p = tmp_path / "gt.txt"
    p.write_text("# comment\n0.0 0 0 0 0 0 0 1\n0.5 1.0 2.0 3.0 0 0 0 1\n")
    t, P = ev.read_ground_truth(str(p), first_keyframe_time=100.0)
    assert np.allclose(t, [100.0, 100.5])
    assert np.allclose(P[1, 4:], [3.0, -1.0, -2.0])                # tf * t
    assert np.allclose(np.abs(P[:, 3]), 1.0)                        # R_tf R R_tf^T = I for R = I
    # a rotation about the camera's y axis (yaw, down) becomes a rotation about body -z... i.e. about z with flipped sign
    a = 0.3
    p.write_text("0.0 0 0 0 0 %.17g 0 %.17g\n" % (np.sin(a / 2), np.cos(a / 2)))
    _, P = ev.read_ground_truth(str(p))
    R = ev.quat_to_matrix(P[0, :4])
    assert np.allclose(R, [[np.cos(a), np.sin(a), 0], [-np.sin(a), np.cos(a), 0], [0, 0, 1]], atol=1e-12)
    q = ev.matrix_to_quat(ev.quat_to_matrix(np.array([[0.1, -0.7, 0.2, 0.3], [0.9, 0.1, 0.1, -0.05]])))
    assert np.allclose(ev.quat_to_matrix(q), ev.quat_to_matrix(np.array([[0.1, -0.7, 0.2, 0.3], [0.9, 0.1, 0.1, -0.05]])))


def test_ape_recovers_injected_noise():
    rng = np.random.default_rng(1)
    P = synth.Trajectory(200).poses()
    t = 0.1 * np.arange(200)
    # reference = estimate moved by a rigid transform, plus noise
    ang = 0.4
    Rz = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]])
    noise = rng.normal(0, 0.02, (200, 3))
    G = P.copy()
    G[:, 4:] = P[:, 4:] @ Rz.T + [5.0, -3.0, 1.0] + noise
    st = ev.ape(t, P, t + 0.002, G, max_dt=0.01)
    assert st["n"] == 200
    assert abs(st["rmse"] - np.sqrt((noise ** 2).sum(1).mean())) < 2e-3
    raw = ev.ape(t, P, t, G, align=False)
    assert raw["rmse"] > 1.0
    s, R, tt = ev.umeyama(P[:, 4:], 2.0 * P[:, 4:] @ Rz.T + 1.0, with_scale=True)
    assert abs(s - 2.0) < 1e-9 and np.allclose(R, Rz, atol=1e-9)
    ie, ir = ev.associate([0.0, 1.0, 2.0], [0.004, 1.5, 2.2], 0.01)
    assert ie.tolist() == [0] and ir.tolist() == [0]
