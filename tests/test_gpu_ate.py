"""ATE of the CUDA solver vs the oracle's on a synthetic KITTI-shaped trajectory (north_star: "ATE within 1 cm of the
reference").  The real sequences are not available offline, so the trajectory, the factors and the ground truth are
synthetic; both solvers get the same window problem and are scored with the same evo-style APE (evaluation.ape)."""
import numpy as np
import pytest

from lvio_fusion_b200 import backend, evaluation, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_kf,imu", [(20, True), (40, False)])
def test_ate_matches_oracle(lvb_ctx, orc_ctx, n_kf, imu, tmp_path):
    d = synth.make_ba_problem(n_kf, 200 * n_kf, with_imu=imu, seed=21)
    if not imu:
        d["factors"][backend.POSE_PRIOR] = (np.concatenate([d["poses"][0], [100.0, 0.0]])[None], np.zeros((1, 1), dtype=np.int32))
    t = 0.1 * np.arange(n_kf)
    stats = {}
    for name, ctx in (("cuda", lvb_ctx), ("oracle", orc_ctx)):
        p = backend.Problem.from_dict(ctx, d)
        p.solve(max_num_iterations=30, num_threads=4)
        path = tmp_path / ("result_%s.csv" % name)
        evaluation.write_result(str(path), t, p.poses())          # through the reference's own result format
        te, Pe = evaluation.read_result(str(path))
        stats[name] = evaluation.ape(te, Pe, t, d["poses_true"])
    init = evaluation.ape(t, d["poses"], t, d["poses_true"])
    assert stats["cuda"]["rmse"] < 0.5 * init["rmse"]
    assert abs(stats["cuda"]["rmse"] - stats["oracle"]["rmse"]) < 1e-4          # 0.1 mm; the target is 1 cm
    assert abs(stats["cuda"]["max"] - stats["oracle"]["max"]) < 1e-3
