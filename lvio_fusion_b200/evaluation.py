"""Result / ground-truth wire formats of the reference node and an ATE harness.

* ``write_result``      -- ``write_result`` of lvio_fusion_node.cpp:295-317: one row per keyframe,
  ``t - init_time,x,y,z,qx,qy,qz,qw`` in fixed notation with 5 decimals.
* ``read_ground_truth`` -- ``read_ground_truth`` of lvio_fusion_node.cpp:319-350: TUM rows
  ``time x y z qx qy qz qw`` mapped from the KITTI camera convention into the body frame of lvio_fusion
  (``a.so3() = a.so3() * q_tf^-1 ; pose = tf * a`` with ``R_tf = [[0,0,1],[-1,0,0],[0,-1,0]]``).
* ``ape``               -- absolute pose error of the translation part after a rigid (Umeyama, no scale) alignment,
  the statistic behind the reference's evo plots.  Host-side numpy: evaluation tooling, not the hot path.

Poses are ``[qx, qy, qz, qw, tx, ty, tz]`` (Sophus::SE3d::data()) as everywhere in this package.
"""
import numpy as np

R_TF = np.array([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])


def quat_to_matrix(q):
    q = np.asarray(q, dtype=np.float64)
    q = q / np.linalg.norm(q, axis=-1, keepdims=True)
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = 1 - 2 * (y * y + z * z); R[..., 0, 1] = 2 * (x * y - z * w); R[..., 0, 2] = 2 * (x * z + y * w)
    R[..., 1, 0] = 2 * (x * y + z * w); R[..., 1, 1] = 1 - 2 * (x * x + z * z); R[..., 1, 2] = 2 * (y * z - x * w)
    R[..., 2, 0] = 2 * (x * z - y * w); R[..., 2, 1] = 2 * (y * z + x * w); R[..., 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def matrix_to_quat(R):
    """Rotation matrix -> unit quaternion xyzw with w >= 0 (branch on the largest diagonal term)."""
    R = np.asarray(R, dtype=np.float64)
    out = np.empty(R.shape[:-2] + (4,))
    flat_R, flat_o = R.reshape(-1, 3, 3), out.reshape(-1, 4)
    for i, m in enumerate(flat_R):
        t = m[0, 0] + m[1, 1] + m[2, 2]
        if t > 0:
            s = np.sqrt(t + 1.0) * 2
            q = [(m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s, 0.25 * s]
        else:
            k = int(np.argmax([m[0, 0], m[1, 1], m[2, 2]]))
            a, b, c = k, (k + 1) % 3, (k + 2) % 3
            s = np.sqrt(1.0 + m[a, a] - m[b, b] - m[c, c]) * 2
            q = [0.0, 0.0, 0.0, (m[c, b] - m[b, c]) / s]
            q[a] = 0.25 * s; q[b] = (m[b, a] + m[a, b]) / s; q[c] = (m[c, a] + m[a, c]) / s
        q = np.asarray(q)
        flat_o[i] = q if q[3] >= 0 else -q
    return out


def write_result(path, times, poses, init_time=0.0):
    """lvio_fusion_node.cpp:295-317 (``out.setf(ios::fixed); out.precision(5)``; unit quaternion)."""
    times = np.asarray(times, dtype=np.float64)
    poses = np.asarray(poses, dtype=np.float64).reshape(-1, 7)
    assert len(times) == len(poses)
    q = poses[:, :4] / np.linalg.norm(poses[:, :4], axis=1, keepdims=True)
    with open(path, "w") as f:
        for t, qq, p in zip(times, q, poses[:, 4:]):
            f.write(",".join("%.5f" % v for v in (t - init_time, p[0], p[1], p[2], qq[0], qq[1], qq[2], qq[3])) + "\n")


def read_result(path):
    """Rows of ``write_result`` -> (times[n], poses[n,7])."""
    a = np.loadtxt(path, delimiter=",", ndmin=2)
    if a.size == 0:
        return np.zeros(0), np.zeros((0, 7))
    return a[:, 0].copy(), np.concatenate([a[:, 4:8], a[:, 1:4]], axis=1)


def read_ground_truth(path, first_keyframe_time=0.0):
    """lvio_fusion_node.cpp:319-350: returns (times[n] = first_keyframe_time + t, poses[n,7]) in the body convention."""
    rows = []
    with open(path) as f:
        for line in f:
            parts = line.split()
            if len(parts) < 8 or parts[0].startswith("#"):
                continue
            rows.append([float(v) for v in parts[:8]])
    a = np.asarray(rows, dtype=np.float64).reshape(-1, 8)
    R = quat_to_matrix(a[:, 4:8]) @ R_TF.T          # a.so3() * q_tf^-1
    R = R_TF @ R                                     # tf * a
    t = a[:, 1:4] @ R_TF.T
    return first_keyframe_time + a[:, 0], np.concatenate([matrix_to_quat(R), t], axis=1)


def associate(t_est, t_ref, max_dt=0.01):
    """Nearest-timestamp matching (each reference stamp used at most once); returns index arrays (i_est, i_ref)."""
    t_est, t_ref = np.asarray(t_est, dtype=np.float64), np.asarray(t_ref, dtype=np.float64)
    if len(t_est) == 0 or len(t_ref) == 0:
        return np.zeros(0, dtype=int), np.zeros(0, dtype=int)
    order = np.argsort(t_ref)
    pos = np.searchsorted(t_ref[order], t_est)
    ie, ir, used = [], [], set()
    for i, p in enumerate(pos):
        cands = [c for c in (p - 1, p) if 0 <= c < len(order)]
        j = min(cands, key=lambda c: abs(t_ref[order[c]] - t_est[i]))
        if abs(t_ref[order[j]] - t_est[i]) <= max_dt and order[j] not in used:
            used.add(order[j]); ie.append(i); ir.append(order[j])
    return np.asarray(ie, dtype=int), np.asarray(ir, dtype=int)


def umeyama(src, dst, with_scale=False):
    """Least-squares similarity / rigid transform dst ~ s R src + t (Umeyama 1991)."""
    src, dst = np.asarray(src, dtype=np.float64), np.asarray(dst, dtype=np.float64)
    mu_s, mu_d = src.mean(0), dst.mean(0)
    xs, xd = src - mu_s, dst - mu_d
    cov = xd.T @ xs / len(src)
    U, D, Vt = np.linalg.svd(cov)
    S = np.eye(3)
    if np.linalg.det(U) * np.linalg.det(Vt) < 0:
        S[2, 2] = -1
    R = U @ S @ Vt
    s = float(np.trace(np.diag(D) @ S) / xs.var(0).sum()) if with_scale else 1.0
    t = mu_d - s * R @ mu_s
    return s, R, t


def ape(t_est, poses_est, t_ref, poses_ref, max_dt=0.01, align=True, with_scale=False):
    """Absolute translation error statistics after alignment: dict(rmse, mean, median, max, min, std, n)."""
    ie, ir = associate(t_est, t_ref, max_dt)
    if len(ie) < 3:
        raise ValueError("ape: fewer than 3 associated poses")
    pe = np.asarray(poses_est, dtype=np.float64)[ie, 4:]
    pr = np.asarray(poses_ref, dtype=np.float64)[ir, 4:]
    if align:
        s, R, t = umeyama(pe, pr, with_scale)
        pe = s * pe @ R.T + t
    e = np.linalg.norm(pe - pr, axis=1)
    return {"rmse": float(np.sqrt(np.mean(e * e))), "mean": float(e.mean()), "median": float(np.median(e)), "max": float(e.max()),
            "min": float(e.min()), "std": float(e.std()), "n": int(len(e))}
